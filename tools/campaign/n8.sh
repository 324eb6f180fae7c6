# 8-GPU box: torchrun bench at N = 8, 4, 2 (weak scaling + the strong / sharded extras, every line verified against the oracle),
# the group ctx in one process (1 / 2 / 4 / 8), the sharded-opening parity check over NCCL
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02ze_bench_n$n.jsonl 2> gpurun_out/r02ze_bench_n$n.err
  echo "N=$n rc=$?"; grep '^{' gpurun_out/r02ze_bench_n$n.jsonl | cut -c 1-220
done
timeout 600 python tools/group_bench.py --gpus 1,2,4,8 > gpurun_out/r02ze_group_bench.jsonl 2> gpurun_out/r02ze_group_bench.err; echo "group rc=$?"; cut -c 1-400 gpurun_out/r02ze_group_bench.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 tools/sharded_open_check.py > gpurun_out/r02ze_sharded_open_check_n8.txt 2>&1; echo "open check rc=$?"; tail -5 gpurun_out/r02ze_sharded_open_check_n8.txt
timeout 300 env ACCMSM_HAVE_2_GPUS=1 python -m pytest tests/test_gpu_sharded_open.py tests/test_gpu_multi.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -3

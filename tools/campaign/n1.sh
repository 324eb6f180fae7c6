set -x
python bench.py --steps 50 --warmup 5 > gpurun_out/r02y_bench.jsonl 2> gpurun_out/r02y_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_reference.jsonl 2>> gpurun_out/r02y_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-open --no-verify > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_accumulate|k_sums|k_sort|k_fixup|k_leaf|k_finish' --launch-skip 60 --launch-count 30 -o gpurun_out/r02y_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-open --no-verify > /dev/null 2>&1
ncu -i gpurun_out/r02y_full.ncu-rep --page raw --csv > gpurun_out/r02y_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02y_full.ncu-rep
rm -f gpurun_out/r02y_full.ncu-rep
python tools/msm_sweep.py --min 12 --max 24 --curves 0,1 > gpurun_out/r02y_msm_sweep.jsonl 2> gpurun_out/r02y_msm_sweep.err
python tools/no_cliff.py 20 > gpurun_out/r02y_no_cliff.txt 2>&1
python tools/vec_roofline.py > gpurun_out/r02y_vec_roofline.jsonl 2>&1
for w in pc hp nark; do python examples/scaling.py $w 10 18 --cpu-max 14 > gpurun_out/r02y_scaling_$w.jsonl 2>&1; done
python examples/scaling.py pc 20 20 --cpu-max 0 >> gpurun_out/r02y_scaling_pc.jsonl 2>&1
tail -c 600 gpurun_out/r02y_bench.err
tail -3 gpurun_out/r02y_msm_sweep.jsonl | cut -c 1-300

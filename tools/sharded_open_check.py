"""torchrun --nproc-per-node N tools/sharded_open_check.py K : ShardedIpaOpen over NCCL against the oracle, element by element."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import accumulation_b200 as ab
from accumulation_b200.sharded import ShardedIpaOpen, cyclic_shard
from oracle import cref
from tests.test_gpu_ipa_open import oracle_open, sponge_stand_in
from tests.test_gpu_sharded_open import _case
from tests.util import same_point

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{torch.cuda.current_device()}"))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 9
ctx = ab.Context(torch.cuda.current_device())
for curve in (0, 1):
    for indexed in (True, False):
        sf, key, h, xi0, hp, coeffs, z = _case(curve, k, 500 + k)
        shard = cyclic_shard(key, rank, world)
        bases = ctx.register_bases(curve, np.concatenate([shard, h.reshape(1, 8)]))
        bases.precompute()
        so = ShardedIpaOpen(ctx, curve, bases, k, rank=rank, world=world, hiding_index=shard.shape[0])
        squeeze = sponge_stand_in(sf)
        res = so.open(cyclic_shard(coeffs, rank, world), z, squeeze, h_prime_xy=None if indexed else hp, xi0_mont=xi0 if indexed else None)
        if rank == 0:
            el, er, efk, ec, echs = oracle_open(curve, key, coeffs, z, hp, squeeze)
            print(f"curve {curve} indexed {indexed}: l", [same_point(x, y) for x, y in zip(res[0], el)], "r", [same_point(x, y) for x, y in zip(res[1], er)],
                  "fk", np.array_equal(res[2], efk), "c", np.array_equal(res[3], ec), flush=True)
        bases.release()
ctx.close()
dist.barrier()
dist.destroy_process_group()

"""CPU suite (gloo, world_size 2 and 3): host logic of the point-range sharding -- shard_range tiling, the
rank-ordered gather of 128-byte partials, and that summing per-shard commitments gives the whole commitment
(SURVEY.md 8e).  The per-shard points come from the ORACLE here (no GPU in this suite); the GPU suite runs
the same flow with the CUDA path (tests/test_gpu_msm.py::test_sharded_*)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from accumulation_b200.sharded import PARTIAL_WORDS, gather_partials, shard_range
from oracle import cref


def test_shard_range_tiles_exactly():
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                lo, cnt = shard_range(n, r, world)
                assert lo == pos and cnt >= 0
                pos += cnt
            assert pos == n
            sizes = [shard_range(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        curve = 0
        pts = cref.gen_points(curve, 11, n)
        sc = cref.gen_scalars(cref.FQ, 12, n, True)
        lo, cnt = shard_range(n, rank, world)
        xy, inf = cref.commit(curve, pts[lo:lo + cnt], sc[lo:lo + cnt])
        one = cref.to_mont(cref.FP, cref.from_int(1).reshape(1, 4)).reshape(4)
        zz = np.zeros(4, np.uint64) if inf else one          # XYZZ image of an affine point: ZZ = ZZZ = 1
        part = np.concatenate([xy, zz, zz]).view(np.int64)
        allp = gather_partials(torch.from_numpy(part.copy()), world)
        assert tuple(allp.shape) == (world, PARTIAL_WORDS)
        if rank == 0:
            acc, acc_inf = None, 1
            for r in range(world):
                p = allp[r].numpy().view(np.uint64)
                p_inf = int(not p[8:12].any())
                if acc is None:
                    acc, acc_inf = p[:8].copy(), p_inf
                else:
                    acc, acc_inf = cref.point_add(curve, acc, acc_inf, p[:8], p_inf)
            exp = cref.commit(curve, pts, sc)
            q.put(bool(acc_inf == exp[1] and np.array_equal(acc, exp[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_and_combine_over_gloo(world):
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, 1001, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

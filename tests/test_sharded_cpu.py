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


# ---- cyclic sharding of IpaPC::open (accumulation_b200/sharded.py::ShardedIpaOpen): host logic against the oracle

def test_cyclic_shard_keeps_fold_partners_together():
    from accumulation_b200.sharded import cyclic_shard
    n = 64
    idx = np.arange(n)
    for world in (1, 2, 4, 8):
        shards = [cyclic_shard(idx, r, world) for r in range(world)]
        assert sorted(np.concatenate(shards).tolist()) == list(range(n))
        h = n // 2
        while h >= world:                       # partners i, i + h of every round with h >= world share a rank
            for r, s in enumerate(shards):
                assert all(((i + h) % world) == r for i in s if i < h)
            h //= 2


@pytest.mark.parametrize("curve", [0, 1])
def test_z_fold_factor_and_shard_identity_vs_oracle(curve):
    """What ShardedIpaOpen relies on, checked on the oracle's round-by-round folding: after r rounds (i) the z-vector
    is F_r (1, z, z^2, ..); (ii) element g of the length-G vectors equals what shard g (indices g mod G) computes alone;
    (iii) l and r of every round are the sums of the shards' l and r."""
    from accumulation_b200.mirror import _fe_to_int, _int_to_fe, _MODULI
    from accumulation_b200.sharded import cyclic_shard, z_fold_factor
    sf = cref.scalar_field(curve)
    k, G = 5, 4
    n = 1 << k
    pts = cref.gen_points(curve, 300, n + 1)
    key, hp = pts[:n], pts[n]
    a = cref.gen_scalars(sf, 301, n, True)
    z = cref.gen_scalars(sf, 302, 1, True).reshape(4)
    xis = cref.gen_scalars(sf, 303, k, True)
    zv = cref.powers(sf, z, n)
    zi = _fe_to_int(sf, z)
    # per-shard state: z-vector of shard g is z^(g + G i)
    sh = []
    for g in range(G):
        zg = np.stack([_int_to_fe(sf, pow(zi, g + G * i, _MODULI[sf])) for i in range(n // G)])
        sh.append([cyclic_shard(key, g, G), cyclic_shard(a, g, G), zg])
    gk, ga, gz = key, a, zv
    for r in range(k - 2):                       # k - log2(G) rounds
        xi = xis[r]
        xinv = cref.fe_inv(sf, xi.reshape(1, 4)).reshape(4)
        l, rr = cref.ipa_open_round_lr(curve, gk, ga, gz, hp)
        acc_l, acc_r = None, None
        for g in range(G):
            pl, pr = cref.ipa_open_round_lr(curve, sh[g][0], sh[g][1], sh[g][2], hp)
            acc_l = pl if acc_l is None else cref.point_add(curve, acc_l[0], acc_l[1], pl[0], pl[1])
            acc_r = pr if acc_r is None else cref.point_add(curve, acc_r[0], acc_r[1], pr[0], pr[1])
            sh[g] = list(cref.ipa_open_fold(curve, sh[g][0], sh[g][1], sh[g][2], xi, xinv))
        assert acc_l[1] == l[1] and np.array_equal(acc_l[0], l[0]) and acc_r[1] == rr[1] and np.array_equal(acc_r[0], rr[0])
        gk, ga, gz = cref.ipa_open_fold(curve, gk, ga, gz, xi, xinv)
        f = z_fold_factor(sf, z, xis[: r + 1], n)
        exp = np.stack([_int_to_fe(sf, _fe_to_int(sf, f) * pow(zi, i, _MODULI[sf]) % _MODULI[sf]) for i in range(gz.shape[0])])
        assert np.array_equal(gz, exp)
    assert gk.shape[0] == G
    for g in range(G):
        assert np.array_equal(gk[g], sh[g][0][0]) and np.array_equal(ga[g], sh[g][1][0]) and np.array_equal(gz[g], sh[g][2][0])


def _rows_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from accumulation_b200.sharded import gather_rows
        row = torch.arange(12, dtype=torch.int64) + 100 * rank
        out = gather_rows(row, world)
        ok = tuple(out.shape) == (world, 12) and all(int(out[r, 3]) == 100 * r + 3 for r in range(world))
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_gather_rows_over_gloo():
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

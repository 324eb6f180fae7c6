"""GPU parity tests of the multi-GPU IPA opening (SURVEY.md 8e "IPA open folding"; accumulation_b200/sharded.py
ShardedIpaOpen over accmsm_ipa_open_begin_shard / accmsm_ipa_open_round_partial_dev): the key, the coefficients and
the z-vector are sharded cyclically, and (l, r) of every round, the final key and c are bit-exact with the oracle's
round-by-round opening (the reference algorithm, ark-poly-commit ipa_pc open, src/ipa_pc_as/mod.rs:454-462).
On one GPU the ranks are driven as virtual ranks through the same per-rank steps; with >= 2 GPUs the NCCL path runs."""
import os
import socket

import numpy as np
import pytest

import accumulation_b200 as ab
from accumulation_b200.mirror import _MODULI, _fe_to_int, _int_to_fe
from accumulation_b200.sharded import ShardedIpaOpen, cyclic_shard
from oracle import cref
from tests.test_gpu_ipa_open import oracle_open, sponge_stand_in
from tests.util import same_point

pytestmark = pytest.mark.gpu


def _case(curve, k, seed):
    sf = cref.scalar_field(curve)
    n = 1 << k
    pts = cref.gen_points(curve, seed, n + 1)
    key, h = pts[:n], pts[n]
    xi0 = cref.gen_scalars(sf, seed + 1, 1, True).reshape(4)
    hp, hp_inf = cref.point_mul(curve, h, 0, cref.from_mont(sf, xi0.reshape(1, 4)).reshape(4))
    assert hp_inf == 0
    coeffs = cref.gen_scalars(sf, seed + 2, n, True)
    z = cref.gen_scalars(sf, seed + 3, 1, True).reshape(4)
    return sf, key, h, xi0, hp, coeffs, z


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k,world,indexed,fold", [(3, 8, False, 0), (6, 2, False, 0), (7, 4, True, 0), (9, 8, True, 2), (10, 4, False, 3),
                                                  (4, 1, True, 0)])
def test_sharded_open_virtual_ranks_vs_oracle(ctx, curve, k, world, indexed, fold):
    import torch
    sf, key, h, xi0, hp, coeffs, z = _case(curve, k, 400 + k)
    squeeze = sponge_stand_in(sf)
    el, er, efk, ec, echs = oracle_open(curve, key, coeffs, z, hp, squeeze)
    ranks = []
    for r in range(world):
        shard = cyclic_shard(key, r, world)
        xy = np.concatenate([shard, h.reshape(1, 8)]) if indexed else shard
        bases = ctx.register_bases(curve, xy)
        if k - (world.bit_length() - 1) >= 2:
            bases.precompute()
        ranks.append(ShardedIpaOpen(ctx, curve, bases, k, rank=r, world=world, hiding_index=shard.shape[0] if indexed else None))
    if fold:
        ctx.set_ipa_fold(fold, 2)
    try:
        for r, so in enumerate(ranks):
            so.begin(cyclic_shard(coeffs, r, world), z, None if indexed else hp, xi0 if indexed else None)
        l_vec, r_vec, chs, xi = [], [], [], None
        for _ in range(ranks[0].k_local):
            allp = torch.stack([so.round_partials().clone() for so in ranks])       # the all-gather of the NCCL path
            l, r_ = ranks[0].combine(allp)
            for so in ranks[1:2]:
                l2, r2 = so.combine(allp)
                assert same_point(l, l2) and same_point(r_, r2)
            xi = squeeze(xi, l, r_)
            xinv = _int_to_fe(sf, pow(_fe_to_int(sf, xi), -1, _MODULI[sf]))
            for so in ranks:
                so.fold(xi, xinv)
            l_vec.append(l); r_vec.append(r_); chs.append(xi)
        pairs = [so.finish_local() for so in ranks]
        if world > 1:
            fk_all, c_all = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
            tl, tr, fk, c, tch = ranks[0].tail(fk_all, c_all, z, chs, squeeze, xi, None if indexed else hp, xi0 if indexed else None,
                                               h if indexed else None)
            l_vec, r_vec, chs = l_vec + tl, r_vec + tr, chs + tch
        else:
            fk, c = pairs[0]
    finally:
        ctx.set_ipa_fold()
        for so in ranks:
            so.key.release()
    assert len(l_vec) == k
    assert all(same_point(x, y) for x, y in zip(l_vec + r_vec, el + er))
    assert all(np.array_equal(x, y) for x, y in zip(chs, echs))
    assert np.array_equal(fk, efk) and np.array_equal(c, ec)


def _nccl_worker(rank, world, port, k, curve, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        ctx = ab.Context(rank)
        sf, key, h, xi0, hp, coeffs, z = _case(curve, k, 500 + k)
        shard = cyclic_shard(key, rank, world)
        bases = ctx.register_bases(curve, np.concatenate([shard, h.reshape(1, 8)]))
        bases.precompute()
        so = ShardedIpaOpen(ctx, curve, bases, k, rank=rank, world=world, hiding_index=shard.shape[0])
        squeeze = sponge_stand_in(sf)
        res = so.open(cyclic_shard(coeffs, rank, world), z, squeeze, xi0_mont=xi0)
        if rank == 0:
            el, er, efk, ec, echs = oracle_open(curve, key, coeffs, z, hp, squeeze)
            ok = all(same_point(x, y) for x, y in zip(res[0] + res[1], el + er)) and np.array_equal(res[2], efk) and np.array_equal(res[3], ec)
            q.put(bool(ok))
        bases.release(); ctx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_open_over_nccl():
    import torch
    import torch.multiprocessing as mp
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs (runs under gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_nccl_worker, args=(r, world, port, 9, 0, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

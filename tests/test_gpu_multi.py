"""GPU parity tests of the device-group ctx (accmsm_init_multi): the SAME host-pointer entry points
over a key sharded by point range across devices inside the library -- MSM / batch / commit (K2), the IPA decider tail
(K3), hp-as decide and product-polynomial commitments and the element-wise vector calls (K4), registered sparse
matrices (K5), IpaPC::open sessions.

On a single-GPU box the group lists device 0 several times ("virtual shards": every code path of the group layer --
range cutting, per-child uploads from slices of the caller's buffers, partial stores into child 0's gather buffer, the
combine -- runs, only the NVLink hop is local); with >= 2 GPUs a second parametrisation uses real peers.  The scenarios
are the ones the single-device suite checks against the oracle, re-run with the group ctx in place of `ctx`."""
import numpy as np
import pytest

import accumulation_b200 as ab
from oracle import cref
from tests import test_gpu_fused as tf
from tests import test_gpu_ipa_open as to
from tests import test_gpu_msm as tm
from tests import test_gpu_vec as tv
from tests.util import same_point

pytestmark = pytest.mark.gpu


def _device_lists():
    lists = [[0, 0, 0]]
    try:
        import torch
        if torch.cuda.device_count() >= 2:
            lists.append(list(range(min(torch.cuda.device_count(), 8))))
    except Exception:
        pass
    return lists


@pytest.fixture(scope="module", params=_device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def gctx(request):
    """min_shard = 16: keys of >= 32 points are really cut; shorter ones exercise the one-owner paths"""
    c = ab.Context(devices=request.param, min_shard=16)
    assert c.device_count() == len(request.param)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gkeys(gctx):
    out = {}
    for curve in (0, 1):
        pts = cref.gen_points(curve, 100 + curve, 1 << 14)
        out[curve] = (pts, gctx.register_bases(curve, pts))
    yield out
    for _, b in out.values():
        b.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 5000, 1 << 14])
def test_group_msm_sizes(gctx, gkeys, curve, n):
    tm.test_msm_sizes_vs_oracle(gctx, gkeys, curve, n)


@pytest.mark.parametrize("curve", [0, 1])
def test_group_msm_offsets_canonical_and_distributions(gctx, gkeys, curve):
    tm.test_msm_canonical_scalars_and_offset(gctx, gkeys, curve)
    tm.test_msm_reference_fixture_distributions(gctx, gkeys, curve, 4096)
    pts, B = gkeys[curve]
    sf = cref.scalar_field(curve)
    # a range that starts and ends inside shards, and one that lies inside a single shard
    for off, n in ((5461 - 7, 5461 + 20), (6000, 100), (16383, 1)):
        sc = cref.gen_scalars(sf, off, n, True)
        assert same_point(gctx.msm(B, sc, offset=off), cref.commit(curve, pts[off:off + n], sc))


def test_group_golden_vectors(gctx):
    tm.test_msm_golden_vectors(gctx)
    tm.test_ipa_golden_vectors(gctx)
    tv.test_vec_golden_vectors(gctx)


@pytest.mark.parametrize("curve", [0, 1])
def test_group_commit_hiding_generator_in_any_shard(gctx, gkeys, curve):
    tm.test_commit_with_randomizer(gctx, gkeys, curve)
    pts, B = gkeys[curve]
    sf = cref.scalar_field(curve)
    el = cref.gen_scalars(sf, 91, 3000, True)
    r = cref.gen_scalars(sf, 92, 1, True).reshape(4)
    for h in (0, 2999, 3000, 5460, 5461, 11000, 16383):      # owners: first / middle / last shard, inside and outside [0, n)
        assert same_point(gctx.commit(B, el, hiding_index=h, randomizer_mont=r), cref.commit(curve, pts[:3000], el, pts[h], r)), h


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n,k,precompute", [(2048, 3, False), (300, 10, False), (6000, 8, True), (1, 2, False)])
def test_group_msm_batch(gctx, curve, n, k, precompute):
    tm.test_msm_batch(gctx, curve, n, k, precompute)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("c", [0, 8, 13])
def test_group_precomputed_window_tables(gctx, curve, c):
    tm.test_precomputed_window_table(gctx, curve, c)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k", [0, 1, 4, 10, 14])
def test_group_ipa_decide_tail(gctx, gkeys, curve, k):
    tm.test_ipa_decide_tail_vs_oracle(gctx, gkeys, curve, k)


@pytest.mark.parametrize("curve", [0, 1])
def test_group_synthetic_bases_and_download(gctx, curve):
    """every child generates its own range of the seeded key; downloads are stitched back in order"""
    n = 5000
    single = ab.Context(0)
    try:
        ref = single.download_bases(single.register_synthetic_bases(curve, 77, n, first_index=9))
    finally:
        single.close()
    B = gctx.register_synthetic_bases(curve, 77, n, first_index=9)
    assert np.array_equal(gctx.download_bases(B), ref)
    assert np.array_equal(gctx.download_bases(B, 1600, 2000), ref[1600:3600])
    sc = cref.gen_scalars(cref.scalar_field(curve), 5, n, True)
    assert same_point(gctx.msm(B, sc), cref.commit(curve, ref, sc))
    B.release()


@pytest.mark.parametrize("curve,L,zk", [(0, 1 << 16, True), (0, 1 << 16, False), (1, 3000, True), (0, 1, True)])
def test_group_hp_as_decide(gctx, curve, L, zk):
    tf.test_hp_as_decide_fused(gctx, curve, L, zk)


@pytest.mark.parametrize("curve,n_in,L,zk", [(0, 2, 1 << 16, False), (0, 3, 5000, True), (1, 2, 777, True), (0, 1, 100, False), (0, 6, 300, False)])
def test_group_hp_as_product_poly_comm(gctx, curve, n_in, L, zk):
    tf.test_hp_as_product_poly_comm_fused(gctx, curve, n_in, L, zk)


@pytest.mark.parametrize("dense,zk", [(False, False), (False, True), (True, True)])
def test_group_r1cs_nark_matvec_commit(gctx, dense, zk):
    tf.test_r1cs_nark_matvec_commit_fused(gctx, dense, zk)


@pytest.mark.parametrize("field", [0, 1])
def test_group_vector_kernels(gctx, field):
    for n in (1, 257, 1 << 16):
        tv.test_hadamard_scale_vs_oracle(gctx, field, n)
    tv.test_lincomb_ragged_vs_oracle(gctx, field)
    for n_in, zk in ((1, False), (2, True), (3, True)):
        tv.test_tvecs_vs_oracle(gctx, field, n_in, zk)
    tv.test_csr_matvec_vs_oracle(gctx, field)
    tv.test_compute_coeffs_combine_evaluate(gctx, field, 5)


@pytest.mark.parametrize("curve", [0, 1])
def test_group_hp_as_accept_reject(gctx, curve):
    tv.test_hp_as_decide_accept_reject(gctx, curve)


@pytest.mark.parametrize("curve,k,precompute", [(0, 1, False), (0, 5, False), (1, 10, True), (0, 13, True)])
def test_group_ipa_open_sessions(gctx, curve, k, precompute):
    """sessions of a group run on child 0 over the whole key assembled from the shards by peer copies"""
    to.test_ipa_open_vs_oracle_and_verifies(gctx, curve, k, precompute)


@pytest.mark.parametrize("curve,k,m", [(0, 10, 2), (1, 6, 3)])
def test_group_ipa_pc_as_prove_open_combined(gctx, curve, k, m):
    tf.test_ipa_pc_as_prove_open_combined(gctx, curve, k, m)


def test_group_device_ctx_and_dev_entry_points(gctx, gkeys):
    """device-pointer entry points need a single-device ctx: the group refuses them, its children serve them"""
    import torch
    pts, B = gkeys[0]
    d = torch.zeros((16,), dtype=torch.int64, device="cuda:0")
    with pytest.raises(ab.AccmsmError):
        gctx.msm_dev(B, d.data_ptr(), 1)
    with pytest.raises(ab.AccmsmError):
        gctx.combine_partials_dev(0, d.data_ptr(), 1)
    kid = gctx.device_ctx(0)
    n = 300
    Bk = kid.register_bases(0, pts[:n])
    sc = cref.gen_scalars(cref.FQ, 3, n, False)
    d_sc = torch.from_numpy(sc.view(np.int64)).to("cuda:0")
    assert same_point(kid.msm_dev(Bk, d_sc.data_ptr(), n, montgomery=False), cref.msm_ark(0, pts[:n], sc))
    Bk.release()
    with pytest.raises(ab.AccmsmError):
        gctx.device_ctx(gctx.device_count())


def test_group_min_shard_keeps_short_keys_on_one_device():
    """default policy (2^16 points per shard): a 2^12 key is not cut; results are the same"""
    g = ab.Context(devices=[0, 0])
    try:
        pts = cref.gen_points(0, 7, 1 << 12)
        B = g.register_bases(0, pts)
        sc = cref.gen_scalars(cref.FQ, 8, 1 << 12, True)
        assert same_point(g.msm(B, sc), cref.commit(0, pts, sc))
        launches_one = g.device_ctx(1).kernel_launches()
        assert launches_one == 0                    # the second child never ran a kernel
        B.release()
    finally:
        g.close()


def test_group_error_codes(gctx, gkeys):
    pts, B = gkeys[0]
    with pytest.raises(ab.AccmsmError):
        gctx.msm(B, np.zeros((10, 4), np.uint64), offset=(1 << 14) - 5)
    with pytest.raises(ab.AccmsmError):
        gctx.commit(B, np.zeros((10, 4), np.uint64), hiding_index=1 << 14, randomizer_mont=np.zeros(4, np.uint64))
    with pytest.raises(ab.AccmsmError):
        gctx.ipa_final_key(B, np.zeros((15, 4), np.uint64))
    B2 = gctx.register_bases(0, pts[:10])
    B2.release()
    with pytest.raises(ab.AccmsmError):
        gctx.msm(B2, np.zeros((1, 4), np.uint64))

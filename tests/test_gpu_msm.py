"""GPU parity tests of the MSM / commitment path (K1 + K2 + K3 fused) through the C-ABI.

Bit-exact against (i) the committed golden fixtures (independent Python big-int oracle), (ii) the C oracle
restating ark-ec 0.2 VariableBaseMSM on the same seeded inputs, and (iii) at BASELINE.json's full size
(2^20) size-independent properties: linearity, a checksum of partial MSMs, oracle agreement.
Shapes follow the reference's own fixtures."""
import numpy as np
import pytest

import accumulation_b200 as ab
from oracle import cref, pyref
from tests.util import (fe_canon, fe_mont, ints, load_golden, point_result, points_mont, same_point,
                        scalar_distributions)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def keys(ctx):
    """Per curve: 2^14 seeded bases registered once (like a trimmed commitment key)."""
    out = {}
    for curve in (0, 1):
        pts = cref.gen_points(curve, 100 + curve, 1 << 14)
        out[curve] = (pts, ctx.register_bases(curve, pts))
    yield out
    for _, b in out.values():
        b.release()


def test_msm_golden_vectors(ctx):
    for case in load_golden("msm"):
        curve = case["curve"]
        bases = points_mont(curve, case["bases"])
        exp = point_result(curve, case["result"])
        B = ctx.register_bases(curve, bases)
        try:
            got = ctx.msm(B, fe_canon(ints(case["scalars"])), montgomery=False)
            assert same_point(got, exp), case["name"]
            got = ctx.msm(B, fe_mont(cref.scalar_field(curve), ints(case["scalars"])), montgomery=True)
            assert same_point(got, exp), case["name"] + " (montgomery scalars)"
        finally:
            B.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 1 << 10, 5000, 1 << 12, 1 << 14])
def test_msm_sizes_vs_oracle(ctx, keys, curve, n):
    pts, B = keys[curve]
    sc = cref.gen_scalars(cref.scalar_field(curve), 7 * n + curve, n, montgomery=True)
    assert same_point(ctx.msm(B, sc), cref.commit(curve, pts[:n], sc))


@pytest.mark.parametrize("curve", [0, 1])
def test_msm_canonical_scalars_and_offset(ctx, keys, curve):
    pts, B = keys[curve]
    sf = cref.scalar_field(curve)
    sc = cref.gen_scalars(sf, 5, 3000, montgomery=False)
    assert same_point(ctx.msm(B, sc, montgomery=False), cref.msm_ark(curve, pts[:3000], sc))
    scm = cref.gen_scalars(sf, 77, 1000, True)
    assert same_point(ctx.msm(B, scm, offset=123), cref.commit(curve, pts[123:1123], scm))
    # n = 0 is the identity (0, 1, infinity) like ark-ec
    xy, inf = ctx.msm(B, np.zeros((0, 4), np.uint64))
    assert inf == 1 and same_point((xy, inf), point_result(curve, None))


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [33, 4096])
def test_msm_reference_fixture_distributions(ctx, keys, curve, n):
    """constant vectors, (a,..,a,0), zero / one / q-1, one-hot, 128-bit challenges: one hot bucket per window."""
    pts, B = keys[curve]
    for name, sc in scalar_distributions(curve, n, 900 + n).items():
        assert same_point(ctx.msm(B, sc), cref.commit(curve, pts[:n], sc)), name


@pytest.mark.parametrize("curve", [0, 1])
def test_msm_duplicate_and_opposite_bases(ctx, curve):
    """examples/scaling-as.rs:91-92 duplicates accumulators: P + P and P + (-P) inside one bucket."""
    sf = cref.scalar_field(curve)
    base = cref.gen_points(curve, 3, 64)
    bm = pyref.base_modulus(curve)
    neg = base.copy()
    y = cref.from_mont(cref.base_field(curve), base[:, 4:])
    neg[:, 4:] = cref.to_mont(cref.base_field(curve), cref.ints_to_arr([(bm - v) % bm for v in cref.arr_to_ints(y)]))
    pts = np.concatenate([base, base, neg, base[:1].repeat(200, axis=0)])
    one = cref.gen_scalars(sf, 4, 1, True)
    rnd = cref.gen_scalars(sf, 5, pts.shape[0], True)
    B = ctx.register_bases(curve, pts)
    try:
        for sc in (np.repeat(one, pts.shape[0], axis=0), rnd):
            assert same_point(ctx.msm(B, sc), cref.commit(curve, pts, sc))
    finally:
        B.release()


@pytest.mark.parametrize("curve", [0, 1])
def test_identity_bases_contribute_nothing(ctx, curve):
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 8, 300)
    inf = (np.arange(300) % 7 == 0).astype(np.uint8)
    sc = cref.gen_scalars(sf, 9, 300, False)
    B = ctx.register_bases(curve, pts, inf)
    try:
        assert same_point(ctx.msm(B, sc, montgomery=False), cref.msm_ark(curve, pts, sc, bases_inf=inf))
    finally:
        B.release()


@pytest.mark.parametrize("curve", [0, 1])
def test_commit_with_randomizer(ctx, keys, curve):
    """PedersenCommitment::commit(ck, elems, Some(r)) = MSM + r * hiding_generator (App. A.2)."""
    pts, B = keys[curve]
    sf = cref.scalar_field(curve)
    for n in (0, 1, 777):
        el = cref.gen_scalars(sf, 21 + n, n, True)
        r = cref.gen_scalars(sf, 22 + n, 1, True).reshape(4)
        got = ctx.commit(B, el, hiding_index=1000, randomizer_mont=r)
        assert same_point(got, cref.commit(curve, pts[:n], el, pts[1000], r))
    ck = ab.CommitterKey.new(ctx, curve, pts[:500], pts[500])
    el = cref.gen_scalars(sf, 30, 500, True)
    r = cref.gen_scalars(sf, 31, 1, True).reshape(4)
    assert same_point(ab.PedersenCommitment.commit(ck, el, r), cref.commit(curve, pts[:500], el, pts[500], r))
    assert same_point(ab.PedersenCommitment.commit(ck, el), cref.commit(curve, pts[:500], el))
    # elems longer than the key are truncated like ark-ec's zip
    el2 = cref.gen_scalars(sf, 32, 600, True)
    assert same_point(ab.PedersenCommitment.commit(ck, el2), cref.commit(curve, pts[:500], el2[:500]))
    ck.bases.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n,k,precompute", [(2048, 3, False), (300, 10, False), (6000, 8, True), (1, 2, False)])
def test_msm_batch(ctx, curve, n, k, precompute):
    """k scalar vectors over one key in shared passes of the pipeline (hp_as::decide: 3, NARK prove: 3-8);
    includes an all-zero vector (identity result) and a constant vector among the jobs."""
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 140 + curve, max(n, 1) + 5)
    B = ctx.register_bases(curve, pts)
    if precompute:
        B.precompute(12)
    sc = cref.gen_scalars(sf, 40 + n, n * k, True).reshape(k, n, 4)
    sc[1] = 0
    if k > 2:
        sc[2] = sc[2][0]
    xy, inf = ctx.msm_batch(B, sc, offset=3)
    for j in range(k):
        assert same_point((xy[j], inf[j]), cref.commit(curve, pts[3:3 + n], sc[j])), j
    B.release()


@pytest.mark.parametrize("c", [4, 7, 8, 11, 13, 15, 16])
def test_window_bits_sweep(ctx, keys, c):
    """every signed-digit window width gives the same point (the final-carry bug lives here, App. D.4)."""
    pts, B = keys[0]
    q = pyref.scalar_modulus(0)
    n = 600
    vals = [q - 1, q - 2, (1 << 254), (1 << 254) - 1, (1 << 255) % q, 1, 0] + [pyref.SplitMix64(c).field(q) for _ in range(n - 7)]
    sc = cref.ints_to_arr(vals)
    ctx.set_window_bits(c)
    try:
        assert same_point(ctx.msm(B, sc, montgomery=False), cref.msm_ark(0, pts[:n], sc))
    finally:
        ctx.set_window_bits(0)


def test_ipa_golden_vectors(ctx):
    for case in load_golden("ipa"):
        curve = case["curve"]
        sf = cref.scalar_field(curve)
        ch = fe_mont(sf, ints(case["challenges"]))
        key = points_mont(curve, case["key"])
        exp = point_result(curve, case["final_key"])
        B = ctx.register_bases(curve, key)
        try:
            assert same_point(ctx.ipa_final_key(B, ch), exp)
            ok, xy, inf = ctx.ipa_check_final_key(B, ch, exp[0], exp[1])
            assert ok and same_point((xy, inf), exp)
            bad = exp[0].copy(); bad[5] ^= np.uint64(1 << 17)
            assert not ctx.ipa_check_final_key(B, ch, bad, 0)[0]
            assert (ctx.compute_coeffs(sf, ch) == fe_mont(sf, ints(case["coeffs"]))).all()
        finally:
            B.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k", [0, 1, 4, 10, 14])
def test_ipa_decide_tail_vs_oracle(ctx, keys, curve, k):
    """decide = check: final_key = cm_commit(key, h.compute_coeffs()) == proof.final_comm_key
    (src/ipa_pc_as/mod.rs:836-845); reject on any corrupted challenge / key / expected point."""
    pts, B = keys[curve]
    sf = cref.scalar_field(curve)
    ch = cref.gen_scalars(sf, 1000 + k, k, True)
    got = ctx.ipa_final_key(B, ch)
    ok, exp_xy, exp_inf = cref.ipa_check_final_key(curve, pts[: 1 << k], ch, got[0], got[1])
    assert ok and same_point(got, (exp_xy, exp_inf))
    if k <= 10:   # two independent computations of the same point: MSM(key, coeffs) == key folded k times
        assert same_point(got, cref.ipa_fold_key(curve, pts[: 1 << k], ch))
    assert ab.InnerProductArgPC.check_final_key(ab.CommitterKey(B, B.n), ch, exp_xy, exp_inf)
    if k:
        bad = ch.copy(); bad[k // 2, 1] ^= np.uint64(4)
        assert not ctx.ipa_check_final_key(B, bad, exp_xy, exp_inf)[0]


@pytest.mark.parametrize("precompute", [False, True])
def test_msm_full_size_2_20(ctx, precompute):
    """BASELINE config 5 headline size: oracle agreement, linearity and the split-sum checksum at 2^20."""
    n = 1 << 20
    pts = cref.gen_points(0, 0xACC5, n)
    B = ctx.register_bases(0, pts)
    if precompute:
        B.precompute()
    try:
        a = cref.gen_scalars(cref.FQ, 1, n, True)
        b = cref.gen_scalars(cref.FQ, 2, n, True)
        ca, cb = ctx.msm(B, a), ctx.msm(B, b)
        assert same_point(ca, cref.commit(0, pts, a))
        cab = ctx.msm(B, cref.fe_add(cref.FQ, a, b))
        assert same_point(cref.point_add(0, ca[0], ca[1], cb[0], cb[1]), cab)          # linearity
        h = n // 2 + 12345                                                             # checksum of partial sums
        lo, hi = ctx.msm(B, a[:h]), ctx.msm(B, a[h:], offset=h)
        assert same_point(cref.point_add(0, lo[0], lo[1], hi[0], hi[1]), ca)
        const = np.repeat(a[:1], n, axis=0)                                            # hot bucket at full size
        assert same_point(ctx.msm(B, const), cref.commit(0, pts, const))
        # ipa decide tail at degree 2^20 (config 4, one GPU)
        ch = cref.gen_scalars(cref.FQ, 3, 20, True)
        got = ctx.ipa_final_key(B, ch)
        assert same_point(got, cref.commit(0, pts, cref.compute_coeffs(cref.FQ, ch)))
    finally:
        B.release()


def _splitmix64(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


@pytest.mark.parametrize("curve", [0, 1])
def test_synthetic_bases(ctx, curve):
    """register_synthetic_bases: base i = s_i * G, reproducible per global index (shards line up)."""
    seed, n = 0xACC5, 300
    B = ctx.register_synthetic_bases(curve, seed, n)
    pts = ctx.download_bases(B)
    bf = cref.base_field(curve)
    G = fe_mont(bf, list(pyref.generator(curve))).reshape(8)
    for i in (0, 1, 17, n - 1):
        assert cref.on_curve(curve, pts[i])
        words = [_splitmix64(seed ^ _splitmix64(i * 4 + j)) for j in range(4)]
        s = sum(w << (64 * j) for j, w in enumerate(words))
        s = (s & ((1 << 254) - 1)) | 1
        exp, inf = cref.point_mul(curve, G, 0, cref.from_int(s))
        assert inf == 0 and np.array_equal(exp, pts[i])
    B2 = ctx.register_synthetic_bases(curve, seed, 100, first_index=150)
    assert np.array_equal(ctx.download_bases(B2), pts[150:250])
    sc = cref.gen_scalars(cref.scalar_field(curve), 5, n, True)
    assert same_point(ctx.msm(B, sc), cref.commit(curve, pts, sc))
    B.release(); B2.release()


def test_device_resident_scalars_and_sharded_flow(ctx, keys):
    """accmsm_msm_dev and the partial/combine pair used by the one-process-per-GPU sharding (8e), driven from
    one process: 3 point-range shards -> 3 partials -> combine == whole MSM; same for the IPA decider tail."""
    import torch
    from accumulation_b200.sharded import ShardedMSM, shard_range
    pts, B = keys[0]
    n = 1 << 13
    sc = cref.gen_scalars(cref.FQ, 300, n, True)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    exp = cref.commit(0, pts[:n], sc)
    from accumulation_b200.sharded import _stream_handle
    st = _stream_handle()      # torch's default stream by its explicit handle (0 would mean the ctx stream)
    assert same_point(ctx.msm_dev(B, d_sc.data_ptr(), n, stream=st), exp)
    assert same_point(ctx.msm_dev(B, d_sc.data_ptr(), n), exp)
    world = 3
    parts = torch.zeros((world, 16), dtype=torch.int64, device="cuda")
    for r in range(world):
        lo, cnt = shard_range(n, r, world)
        ctx.msm_partial_dev(B, d_sc[lo:].data_ptr(), cnt, parts[r].data_ptr(), offset=lo, stream=st)
    assert same_point(ctx.combine_partials_dev(0, parts.data_ptr(), world, stream=st), exp)
    # IPA: coefficient ranges expanded per shard
    k = 13
    ch = cref.gen_scalars(cref.FQ, 301, k, True)
    whole = ctx.ipa_final_key(B, ch)
    shards = []
    for r in range(world):
        lo, cnt = shard_range(n, r, world)
        Br = ctx.register_bases(0, pts[lo:lo + cnt])
        shards.append(Br)
        ctx.ipa_final_key_partial_dev(Br, ch, lo, cnt, parts[r].data_ptr(), stream=st)
    assert same_point(ctx.combine_partials_dev(0, parts.data_ptr(), world, stream=st), whole)
    for Br in shards:
        Br.release()
    # world = 1 ShardedMSM wrapper (what bench.py drives)
    sh = ShardedMSM(ctx, 0, pts[:n], n, 0, 1)
    assert same_point(sh.msm_dev(d_sc), exp)
    h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory()
    assert same_point(sh.msm_host(h_sc, torch.empty_like(d_sc)), exp)
    assert same_point(sh.ipa_final_key(ch, k), whole)
    sh.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("c", [0, 4, 8, 11, 13])
def test_precomputed_window_table(ctx, curve, c):
    """accmsm_precompute_bases: table[w][i] = 2^(c w) P_i, one shared bucket set.  Same points bit for bit
    as the plain path and the oracle, for every size / distribution / offset / commit / IPA entry point."""
    sf = cref.scalar_field(curve)
    N = 1 << 13
    pts = cref.gen_points(curve, 500 + curve, N)
    B = ctx.register_bases(curve, pts).precompute(c)
    try:
        for n in (1, 33, 1000, 4096, 5000, N):          # short MSMs over a long key use the table too
            sc = cref.gen_scalars(sf, 600 + n, n, True)
            assert same_point(ctx.msm(B, sc), cref.commit(curve, pts[:n], sc)), n
        n = 4500
        for name, sc in scalar_distributions(curve, n, 700).items():
            assert same_point(ctx.msm(B, sc), cref.commit(curve, pts[:n], sc)), name
        sc = cref.gen_scalars(sf, 701, 4200, False)
        assert same_point(ctx.msm(B, sc, montgomery=False, offset=777), cref.msm_ark(curve, pts[777:777 + 4200], sc))
        el = cref.gen_scalars(sf, 702, 6000, True)
        r = cref.gen_scalars(sf, 703, 1, True).reshape(4)
        assert same_point(ctx.commit(B, el, hiding_index=N - 1, randomizer_mont=r), cref.commit(curve, pts[:6000], el, pts[N - 1], r))
        ch = cref.gen_scalars(sf, 704, 13, True)
        got = ctx.ipa_final_key(B, ch)
        assert same_point(got, cref.commit(curve, pts, cref.compute_coeffs(sf, ch)))
        assert ctx.ipa_check_final_key(B, ch, got[0], got[1])[0]
    finally:
        B.release()


@pytest.mark.parametrize("c", [17, 20])
def test_precomputed_wide_windows(ctx, c):
    """window tables with c > 16 (more buckets than points per window; tiled multi-CTA scan)."""
    N = 1 << 17
    pts = cref.gen_points(0, 800 + c, N)
    B = ctx.register_bases(0, pts).precompute(c)
    try:
        sc = cref.gen_scalars(cref.FQ, 801, N, True)
        assert same_point(ctx.msm(B, sc), cref.commit(0, pts, sc))
        n = (1 << 16) + 4321
        sc = cref.gen_scalars(cref.FQ, 802, n, False)
        assert same_point(ctx.msm(B, sc, montgomery=False, offset=999), cref.msm_ark(0, pts[999:999 + n], sc))
        const = np.repeat(cref.gen_scalars(cref.FQ, 803, 1, True), N, axis=0)
        assert same_point(ctx.msm(B, const), cref.commit(0, pts, const))
    finally:
        B.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [0, 1, 7, 42, 3000])
def test_msm_oneshot_unregistered_bases(ctx, curve, n):
    """the literal VariableBaseMSM::multi_scalar_mul(&bases, &scalars) signature (commitment linear combinations:
    src/hp_as/mod.rs:391-406; 2 k + 2 terms of succinct_check at k = 20 is n = 42)"""
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 900 + n, max(n, 1))[:n]
    sc = cref.gen_scalars(sf, 901 + n, n, False)
    inf = (np.arange(n) % 5 == 1).astype(np.uint8)
    assert same_point(ctx.msm_oneshot(curve, pts, sc, montgomery=False), cref.msm_ark(curve, pts, sc) if n else point_result(curve, None))
    if n:
        assert same_point(ctx.msm_oneshot(curve, pts, sc, montgomery=False, infinity=inf), cref.msm_ark(curve, pts, sc, bases_inf=inf))
        scm = cref.to_mont(sf, sc)
        assert same_point(ctx.msm_oneshot(curve, pts, scm[: max(n - 1, 0)]), cref.msm_ark(curve, pts[: n - 1], sc[: n - 1]) if n > 1 else point_result(curve, None))


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n,m", [(0, 3), (1, 1), (43, 3), (39, 11), (300, 8), (5000, 2)])
def test_msm_oneshot_batch(ctx, curve, n, m):
    """m one-shot MSMs of equal length over their own bases in shared passes (the succinct-check equations of all inputs of
    one ipa-pc-as prove: 2 k + 3 terms each, 43 at k = 20): every result equals the single-call result and the oracle's;
    m > 8 spans several passes; identity bases and Montgomery-form scalars as in the single call"""
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 1300 + n, max(n * m, 1))[: n * m].reshape(m, n, 8)
    sc = cref.gen_scalars(sf, 1301 + n, n * m, False).reshape(m, n, 4)
    if n and m > 1:
        sc[1, 0] = 0                                      # a zero scalar, a scalar == 1 (ark's shortcut), a repeated base
        sc[1, n - 1] = cref.from_int(1)
        pts[m - 1, n - 1] = pts[m - 1, 0]
    inf = ((np.arange(n * m) % 7) == 2).astype(np.uint8).reshape(m, n)
    xy, oinf = ctx.msm_oneshot_batch(curve, pts, sc, montgomery=False)
    assert xy.shape == (m, 8) and oinf.shape == (m,)
    for j in range(m):
        exp = cref.msm_ark(curve, pts[j], sc[j]) if n else point_result(curve, None)
        assert same_point((xy[j], int(oinf[j])), exp)
        if n:
            assert same_point((xy[j], int(oinf[j])), ctx.msm_oneshot(curve, pts[j], sc[j], montgomery=False))
    if n:
        xy2, oinf2 = ctx.msm_oneshot_batch(curve, pts, cref.to_mont(sf, sc.reshape(-1, 4)).reshape(m, n, 4), montgomery=True, infinity=inf)
        for j in range(m):
            assert same_point((xy2[j], int(oinf2[j])), cref.msm_ark(curve, pts[j], sc[j], bases_inf=inf[j]))


def test_pinned_host_buffers(ctx, keys):
    """accmsm_host_alloc: page-locked scalar buffers give the same result (and the PCIe-rate H2D path)"""
    pts, B = keys[0]
    n = 5000
    sc = cref.gen_scalars(cref.FQ, 950, n, True)
    pin = ab.pinned_array((n, 4))
    pin[:] = sc
    assert same_point(ctx.msm(B, pin), cref.commit(0, pts[:n], sc))
    ab.release_pinned(pin)


def test_error_codes_never_abort(ctx, keys):
    """error convention of the boundary: 0 or a negative ACCMSM_E_* code plus a message, never a crash"""
    import ctypes as C
    pts, B = keys[0]
    lib, h = ctx._lib, ctx._h
    out = np.empty(8, np.uint64); inf = C.c_uint8(0)
    sc = cref.gen_scalars(cref.FQ, 1, 10, True)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.accmsm_msm(h, C.c_uint64(987654321), C.c_size_t(0), C.c_size_t(10), p(sc), 1, p(out), C.byref(inf)) == -3      # E_HANDLE
    assert lib.accmsm_msm(h, C.c_uint64(B.handle), C.c_size_t(B.n - 3), C.c_size_t(10), p(sc), 1, p(out), C.byref(inf)) == -2  # range
    assert b"range" in lib.accmsm_last_error(h)
    assert lib.accmsm_msm(h, C.c_uint64(B.handle), C.c_size_t(0), C.c_size_t(10), None, 1, p(out), C.byref(inf)) == -2          # null scalars
    assert lib.accmsm_commit(h, C.c_uint64(B.handle), C.c_size_t(10), p(sc), C.c_size_t(B.n + 5), p(sc), p(out), C.byref(inf)) == -2
    assert lib.accmsm_ipa_final_key(h, C.c_uint64(B.handle), p(sc), 31, p(out), C.byref(inf)) == -2                             # k too large
    assert lib.accmsm_ipa_final_key(h, C.c_uint64(B.handle), p(cref.gen_scalars(cref.FQ, 2, 20, True)), 20, p(out), C.byref(inf)) == -2   # key shorter than 2^k
    assert lib.accmsm_set_window_bits(h, 1) == -2 and lib.accmsm_set_window_bits(h, 40) == -2
    assert lib.accmsm_precompute_bases(h, C.c_uint64(B.handle), 30) == -2
    assert lib.accmsm_release_bases(h, C.c_uint64(555)) == -3
    sess = C.c_uint64(0)
    assert lib.accmsm_ipa_open_round(h, C.c_uint64(4242), p(out), C.byref(inf), p(out), C.byref(inf)) == -3
    assert lib.accmsm_ipa_open_begin(h, C.c_uint64(B.handle), p(sc), C.c_size_t(10), 2, p(sc), p(out), C.byref(sess)) == -2     # 10 coeffs > 2^2
    with pytest.raises(ab.AccmsmError):
        ctx.msm(B, sc, offset=B.n)
    # the context still works afterwards
    assert same_point(ctx.msm(B, sc), cref.commit(0, pts[:10], sc))


def test_randomised_shapes_stress(ctx):
    """seeded sweep over random (curve, key size, table or not, offset, length, scalar form, distribution): every result
    bit-exact with the oracle.  Exercises the size-dependent dispatch (warp-per-bucket vs balanced accumulation, tiled
    scan, one- / two-level reduction, window rule) at sizes no other test names explicitly."""
    rng = np.random.default_rng(0xACC)
    for trial in range(40):
        curve = int(rng.integers(0, 2))
        sf = cref.scalar_field(curve)
        N = int(2 ** rng.uniform(1, 15.5))
        pts = cref.gen_points(curve, 5000 + trial, N)
        B = ctx.register_bases(curve, pts)
        mode = int(rng.integers(0, 3))
        if mode == 1:
            B.precompute()
        elif mode == 2:
            B.precompute(int(rng.integers(4, 19)))
        try:
            for _ in range(3):
                n = int(rng.integers(1, N + 1))
                off = int(rng.integers(0, N - n + 1))
                mont = bool(rng.integers(0, 2))
                sc = cref.gen_scalars(sf, int(rng.integers(1, 1 << 30)), n, mont)
                kind = int(rng.integers(0, 4))
                if kind == 1:
                    sc[:] = sc[0]
                elif kind == 2:
                    sc[rng.integers(0, n, max(n // 3, 1))] = 0
                elif kind == 3 and not mont:
                    sc[:, 2:] = 0
                exp = cref.commit(curve, pts[off:off + n], sc) if mont else cref.msm_ark(curve, pts[off:off + n], sc)
                got = ctx.msm(B, sc, montgomery=mont, offset=off)
                assert same_point(got, exp), (trial, curve, N, mode, n, off, mont, kind)
        finally:
            B.release()


def test_experimental_batch_affine_rounds(monkeypatch):
    """the batch-affine pre-reduction (off by default: measured slower than the XYZZ path so far) must
    still give the same points: shared inversion, doubling / cancelling / pass-through pairs, identity markers"""
    monkeypatch.setenv("ACCMSM_AFFINE_ROUNDS", "3")
    c2 = ab.Context(0)
    try:
        for curve in (0, 1):
            sf = cref.scalar_field(curve)
            n = 20000
            base = cref.gen_points(curve, 990 + curve, n)
            bm = pyref.base_modulus(curve)
            neg = base[:50].copy()
            y = cref.from_mont(cref.base_field(curve), neg[:, 4:])
            neg[:, 4:] = cref.to_mont(cref.base_field(curve), cref.ints_to_arr([(bm - v) % bm for v in cref.arr_to_ints(y)]))
            pts = np.concatenate([base, base[:50], neg, base[:1].repeat(33, axis=0)])      # duplicates and opposite points
            B = c2.register_bases(curve, pts).precompute(10)
            N = pts.shape[0]
            rnd = cref.gen_scalars(sf, 991, N, True)
            const = np.repeat(rnd[:1], N, axis=0)
            for sc in (rnd, const):
                assert same_point(c2.msm(B, sc), cref.commit(curve, pts, sc))
            B.release()
    finally:
        c2.close()


@pytest.mark.parametrize("curve,pinned", [(0, False), (0, True), (1, True)])
def test_chunked_upload_overlapping_the_digit_kernel(ctx, curve, pinned):
    """Host scalar vectors of >= 16 MiB go up in 4 MiB chunks on a copy stream, cut into two point segments that add into one
    bucket set (msm_host_scalars: the second segment travels while the first is sorted and accumulated, fix-up on the side
    stream); ragged last chunk, pageable (staged by several threads) and page-locked sources, both curves, result bit-exact
    with the oracle and with the same MSM from device-resident scalars."""
    import torch
    n = (1 << 19) + 4097
    key = ctx.register_synthetic_bases(curve, 77, n)
    key.precompute()
    sc = cref.gen_scalars(cref.scalar_field(curve), 78, n, True)
    if pinned:
        buf = ab.pinned_array((n, 4))
        buf[:] = sc
        src = buf
    else:
        src = sc
    got = ctx.msm(key, src)
    again = ctx.msm(key, src)                         # staging buffer and events are reused
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    dev = ctx.msm_dev(key, d_sc.data_ptr(), n)
    assert same_point(got, dev) and same_point(again, dev)
    pts = ctx.download_bases(key)
    assert same_point(got, cref.commit(curve, pts, sc))
    if pinned:
        ab.release_pinned(buf)
    key.release()


@pytest.mark.parametrize("kind", ["constant", "a_a_a_0", "two_values", "trunc128"])
def test_pipelined_upload_with_degenerate_scalar_vectors(ctx, kind):
    """The two-segment pipeline of a large host-scalar MSM (both segments add into one bucket set; fix-up and the gated
    fallback sort on the side stream) on the reference's degenerate scalar vectors: a constant vector (src/hp_as/mod.rs:991)
    and (a, ..., a, 0) (examples/scaling-nark.rs:48-52) close the skew gate of the radix sort in BOTH segments, so the
    counting sort feeds the `into` accumulation with one hot bucket per window; two values and 128-bit challenges likewise
    stress single partitions.  Bit-exact with the oracle and with the device-resident (unsegmented) MSM."""
    import torch
    n = (1 << 19) + 333
    key = ctx.register_synthetic_bases(0, 79, n)
    key.precompute()
    rnd = cref.gen_scalars(cref.FQ, 80, n, False)
    if kind == "constant":
        sc = np.repeat(rnd[:1], n, axis=0)
    elif kind == "a_a_a_0":
        sc = np.repeat(rnd[:1], n, axis=0); sc[-1] = 0
    elif kind == "two_values":
        sc = np.where((np.arange(n) % 3 == 0)[:, None], rnd[0], rnd[1])
    else:
        sc = rnd.copy(); sc[:, 2:] = 0
    sc = np.ascontiguousarray(sc, dtype=np.uint64)
    got = ctx.msm(key, sc, montgomery=False)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    dev = ctx.msm_dev(key, d_sc.data_ptr(), n, montgomery=False)
    assert same_point(got, dev)
    assert same_point(got, cref.msm_ark(0, ctx.download_bases(key), sc))
    key.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("precompute", [False, True])
def test_reduction_tree_meets_equal_and_opposite_points(ctx, curve, precompute):
    """Every base is the SAME point and every scalar a distinct small digit, so every bucket holds exactly that point (or
    its negative): the bucket reduction then adds P + P and P + (-P) at every level of its shuffle trees -- the
    exceptional branches of the (cooperative) addition that random inputs never reach."""
    sf = cref.scalar_field(curve)
    g = cref.gen_points(curve, 4242, 1)
    for n, neg_every in ((512, 0), (512, 2), (300, 3), (31, 0), (2, 2)):
        pts = np.repeat(g, n, axis=0)
        vals = [i + 1 for i in range(n)]
        q = pyref.Q_PALLAS_SCALAR if sf == cref.FQ else pyref.P_PALLAS_BASE
        ints_ = [(q - v) if (neg_every and i % neg_every == 0) else v for i, v in enumerate(vals)]
        sc = fe_mont(sf, ints_)
        key = ctx.register_bases(curve, pts)
        if precompute:
            key.precompute()
        assert same_point(ctx.msm(key, sc), cref.commit(curve, pts, sc))
        key.release()

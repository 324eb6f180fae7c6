"""CPU suite: the device headers (fp.cuh / ec.cuh) compiled for the HOST with an emulated carry flag
(tests/host/host_shim.cpp) must agree limb-for-limb with the oracle.  This checks the 8x32-bit Montgomery
multiplier and the XYZZ group law without a GPU; it is test scaffolding, never part of libaccmsm.so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cref, pyref

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "host_shim.cpp")
SO = os.path.join(HERE, "host", "libhostshim.so")


@pytest.fixture(scope="module")
def shim():
    deps = [SRC] + [os.path.join(HERE, "..", "accumulation_b200", "csrc", f) for f in ("fp.cuh", "ec.cuh", "hostfp.hpp")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    return C.CDLL(SO)


def fe_op(lib, field, code, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    bp = None if b is None else np.ascontiguousarray(b, dtype=np.uint64).ctypes.data_as(C.c_void_p)
    lib.host_fe_op(C.c_int(field), C.c_int(code), a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), C.c_size_t(a.size // 4))
    return out


def fe_mul2(lib, field, sub, a, b, c, d):
    arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (a, b, c, d)]
    out = np.empty_like(arrs[0])
    lib.host_fe_mul2(C.c_int(field), C.c_int(sub), *[x.ctypes.data_as(C.c_void_p) for x in arrs], out.ctypes.data_as(C.c_void_p),
                     C.c_size_t(arrs[0].size // 4))
    return out


@pytest.mark.parametrize("field", [0, 1])
def test_field_limb_algorithms(shim, field):
    m = [pyref.P_PALLAS_BASE, pyref.Q_PALLAS_SCALAR][field]
    n = 20000
    a = cref.gen_scalars(field, 1, n, True)
    b = cref.gen_scalars(field, 2, n, True)
    edge = cref.ints_to_arr([0, 1, m - 1, (1 << 256) % m, m - 2, 2, (1 << 255) % m, 0xffffffff, 1 << 32, m >> 1,
                             (1 << 254) - 1, (1 << 224), m - (1 << 32)])
    ea, eb = np.repeat(edge, len(edge), axis=0), np.tile(edge, (len(edge), 1))
    a, b = np.concatenate([a, ea]), np.concatenate([b, eb])
    assert (fe_op(shim, field, 0, a, b) == cref.fe_mul(field, a, b)).all()
    assert (fe_op(shim, field, 1, a, b) == cref.fe_add(field, a, b)).all()
    assert (fe_op(shim, field, 2, a, b) == cref.fe_sub(field, a, b)).all()
    assert (fe_op(shim, field, 3, a) == cref.fe_mul(field, a, a)).all()          # dedicated squaring (36 limb products)
    # dual product with one reduction: a b + c d and a b - c d (the Y3 of every addition law in ec.cuh)
    c, d = np.roll(a, 7, axis=0), np.roll(b, 13, axis=0)
    ab, cd = cref.fe_mul(field, a, b), cref.fe_mul(field, c, d)
    assert (fe_mul2(shim, field, 0, a, b, c, d) == cref.fe_add(field, ab, cd)).all()
    assert (fe_mul2(shim, field, 1, a, b, c, d) == cref.fe_sub(field, ab, cd)).all()
    assert (fe_op(shim, field, 5, a) == cref.from_mont(field, a)).all()
    assert (fe_op(shim, field, 6, a) == cref.to_mont(field, a)).all()
    assert (fe_op(shim, field, 7, a) == cref.fe_sub(field, np.zeros_like(a), a)).all()
    assert (fe_op(shim, field, 4, a[:200]) == cref.fe_inv(field, a[:200])).all()
    assert (fe_op(shim, field, 4, ea) == cref.fe_inv(field, ea)).all()
    # binary-GCD inversion (used by the single-thread normalisation in k_finish) == Fermat == oracle
    assert (fe_op(shim, field, 8, a[:3000]) == cref.fe_inv(field, a[:3000])).all()
    assert (fe_op(shim, field, 8, ea) == cref.fe_inv(field, ea)).all()
    small = cref.to_mont(field, cref.ints_to_arr(list(range(0, 70)) + [m - i for i in range(1, 70)] + [1 << i for i in range(0, 255, 7)]))
    assert (fe_op(shim, field, 8, small) == cref.fe_inv(field, small)).all()


@pytest.mark.parametrize("field", [0, 1])
def test_tonelli_shanks_square_roots(shim, field):
    """Fp::sqrt (device header, host build): squares have a root whose square is the input; non-squares are reported"""
    m = [pyref.P_PALLAS_BASE, pyref.Q_PALLAS_SCALAR][field]
    a = cref.gen_scalars(field, 11, 300, True)
    sq = cref.fe_mul(field, a, a)
    assert (fe_op(shim, field, 9, sq) == sq).all()
    vals = cref.arr_to_ints(cref.from_mont(field, a))
    got = fe_op(shim, field, 9, a)
    for v, g, orig in zip(vals, got, a):
        is_sq = pow(v, (m - 1) // 2, m) == 1
        assert (g == orig).all() if is_sq else int(g[0]) & 0xffffffff == 0xdeadbeef
    edge = cref.to_mont(field, cref.ints_to_arr([0, 1, 4, m - 1, 5]))      # -1 is a square (m = 1 mod 4), 5 is not
    got = fe_op(shim, field, 9, edge)
    assert (got[:4] == edge[:4]).all() and int(got[4][0]) & 0xffffffff == 0xdeadbeef


def ec_sum(lib, curve, pts, neg, mode):
    pts = np.ascontiguousarray(pts, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    inf = C.c_uint8(0)
    negp = None if neg is None else np.ascontiguousarray(neg, dtype=np.uint8).ctypes.data_as(C.c_void_p)
    lib.host_ec_sum(C.c_int(curve), pts.ctypes.data_as(C.c_void_p), negp, C.c_size_t(pts.shape[0]), C.c_int(mode),
                    out.ctypes.data_as(C.c_void_p), C.byref(inf))
    return out, inf.value


@pytest.mark.parametrize("curve", [0, 1])
def test_xyzz_group_law(shim, curve):
    sm = pyref.scalar_modulus(curve)
    n = 60
    pts = cref.gen_points(curve, 3 + curve, n)
    neg = (np.arange(n) % 3 == 0).astype(np.uint8)
    sc = np.array([cref.from_int(sm - 1) if neg[i] else cref.from_int(1) for i in range(n)])
    exp, einf = cref.msm_ark(curve, pts, sc)
    for mode in (0, 1, 2):   # mixed adds, full adds, two partials merged
        got, ginf = ec_sum(shim, curve, pts, neg, mode)
        assert ginf == einf and (got == exp).all()
    P, Qp = pts[:1], pts[1:2]
    # the exceptional cases the reference's fixtures hit: P+P, P-P, P+P+P, (P-P)+Q, 2P-2P
    for seq, negs, scal in [([P, P], [0, 0], [2, 0]), ([P, P], [0, 1], [0, 0]), ([P, P, P], [0, 0, 0], [3, 0]),
                            ([P, P, Qp], [0, 1, 0], [0, 1]), ([P, P, P, P], [0, 0, 1, 1], [0, 0])]:
        exp, einf = cref.msm_ark(curve, np.concatenate([P, Qp]), cref.ints_to_arr(scal))
        for mode in (0, 1, 2):
            got, ginf = ec_sum(shim, curve, np.concatenate(seq), np.array(negs, dtype=np.uint8), mode)
            assert ginf == einf and (got == exp).all()      # identity image is (0, R, inf = 1)
    got, ginf = ec_sum(shim, curve, np.repeat(P, 20, axis=0), None, 3)   # 20 doublings of P
    exp, einf = cref.point_mul(curve, P[0], 0, cref.from_int(1 << 20))
    assert ginf == 0 and (got == exp).all()


@pytest.mark.parametrize("field", [0, 1])
def test_library_host_field_code(shim, field):
    """hostfp.hpp (4 x 64-bit limbs; the library's host thread uses it for the affine conversion of returned sums and for the
    inverse of an IpaPC::open round challenge) against the oracle: products, inverses, XYZZ -> affine with one inversion."""
    m = [pyref.P_PALLAS_BASE, pyref.Q_PALLAS_SCALAR][field]
    n = 3000
    a = cref.gen_scalars(field, 11, n, True)
    b = cref.gen_scalars(field, 12, n, True)
    edge = cref.ints_to_arr([0, 1, m - 1, (1 << 256) % m, m - 2, 2, (1 << 255) % m, 0xffffffff, 1 << 32, m >> 1, (1 << 254) - 1, m - (1 << 32)])
    ea, eb = np.repeat(edge, len(edge), axis=0), np.tile(edge, (len(edge), 1))
    a, b = np.concatenate([a, ea]), np.concatenate([b, eb])

    def fast(op, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.empty_like(x)
        yp = None if y is None else np.ascontiguousarray(y, dtype=np.uint64).ctypes.data_as(C.c_void_p)
        shim.host_fast_op(C.c_int(field), C.c_int(op), x.ctypes.data_as(C.c_void_p), yp, out.ctypes.data_as(C.c_void_p), C.c_size_t(x.shape[0]))
        return out

    assert (fast(0, a, b) == cref.fe_mul(field, a, b)).all()
    inv = fast(1, a[:400])
    one = cref.to_mont(field, cref.ints_to_arr([1]))[0]
    prod = cref.fe_mul(field, inv, a[:400])
    assert all((p == one).all() for p in prod)
    assert (fast(1, cref.ints_to_arr([0])) == 0).all()                       # inv(0) = 0, like the device code
    # XYZZ -> affine: points (x, y) scaled by random z (X = x z^2, Y = y z^3, ZZ = z^2, ZZZ = z^3), identities in between
    k = 37
    curve = field
    pts = cref.gen_points(curve, 21 + curve, k)
    z = cref.gen_scalars(field, 13, k, True)
    zz = cref.fe_mul(field, z, z)
    zzz = cref.fe_mul(field, zz, z)
    raw = np.zeros((k, 16), dtype=np.uint64)
    raw[:, 0:4] = cref.fe_mul(field, pts[:, 0:4], zz)
    raw[:, 4:8] = cref.fe_mul(field, pts[:, 4:8], zzz)
    raw[:, 8:12], raw[:, 12:16] = zz, zzz
    ident = [0, 5, 36]
    for i in ident:
        raw[i, 8:16] = 0
    for kk in (1, 2, k):          # the batch sizes share one inversion
        out = np.empty((kk, 8), dtype=np.uint64)
        inf = np.empty(kk, dtype=np.uint8)
        shim.host_fast_xyzz_to_affine(C.c_int(field), raw.ctypes.data_as(C.c_void_p), C.c_size_t(kk), out.ctypes.data_as(C.c_void_p), inf.ctypes.data_as(C.c_void_p))
        for i in range(kk):
            if i in ident:
                assert inf[i] == 1 and (out[i, :4] == 0).all() and (out[i, 4:] == one).all()     # ark's identity image (0, 1, true)
            else:
                assert inf[i] == 0 and (out[i] == pts[i]).all()

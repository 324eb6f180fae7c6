"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (the C restatement).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this.  PARITY UNPINNED at the reference boundary (see oracle/oracle.c header).
All arrays are numpy uint64, little-endian 4-limb field elements; points are (n, 8) = x||y.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

FP, FQ = 0, 1          # field ids
PALLAS, VESTA = 0, 1   # curve ids


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_on_curve.restype = C.c_int
        _lib.oracle_ipa_check_final_key.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def scalar_field(curve):
    return FQ if curve == PALLAS else FP


def base_field(curve):
    return FP if curve == PALLAS else FQ


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int):
    lib().oracle_set_num_threads(C.c_int(int(n)))


def _binop(name, field, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    getattr(lib(), name)(C.c_int(field), _p(a), _p(b), _p(out), C.c_size_t(a.size // 4))
    return out


def fe_mul(field, a, b):
    return _binop("oracle_fe_mul", field, a, b)


def fe_add(field, a, b):
    return _binop("oracle_fe_add", field, a, b)


def fe_sub(field, a, b):
    return _binop("oracle_fe_sub", field, a, b)


def _unop(name, field, a):
    a = _u64(a)
    out = np.empty_like(a)
    getattr(lib(), name)(C.c_int(field), _p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fe_inv(field, a):
    return _unop("oracle_fe_inv", field, a)


def to_mont(field, a):
    return _unop("oracle_fe_to_mont", field, a)


def from_mont(field, a):
    return _unop("oracle_fe_from_mont", field, a)


def gen_scalars(field, seed, n, montgomery=True):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().oracle_gen_scalars(C.c_int(field), C.c_uint64(seed), C.c_size_t(n), C.c_int(int(montgomery)), _p(out))
    return out


def gen_points(curve, seed, n):
    out = np.empty((n, 8), dtype=np.uint64)
    lib().oracle_gen_points(C.c_int(curve), C.c_uint64(seed), C.c_size_t(n), _p(out))
    return out


def on_curve(curve, xy) -> bool:
    xy = _u64(xy)
    return bool(lib().oracle_on_curve(C.c_int(curve), _p(xy)))


def point_mul(curve, xy, inf, scalar_canon):
    xy, k = _u64(xy), _u64(scalar_canon)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().oracle_point_mul(C.c_int(curve), _p(xy), C.c_uint8(int(inf)), _p(k), _p(out), C.byref(oinf))
    return out, int(oinf.value)


def point_add(curve, a_xy, a_inf, b_xy, b_inf):
    a_xy, b_xy = _u64(a_xy), _u64(b_xy)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().oracle_point_add(C.c_int(curve), _p(a_xy), C.c_uint8(int(a_inf)), _p(b_xy), C.c_uint8(int(b_inf)),
                           _p(out), C.byref(oinf))
    return out, int(oinf.value)


def msm_ark(curve, bases_xy, scalars_canon, bases_inf=None):
    bases_xy, scalars_canon = _u64(bases_xy), _u64(scalars_canon)
    inf = None if bases_inf is None else np.ascontiguousarray(bases_inf, dtype=np.uint8)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().oracle_msm_ark(C.c_int(curve), _p(bases_xy), _p(inf), C.c_size_t(bases_xy.size // 8),
                         _p(scalars_canon), C.c_size_t(scalars_canon.size // 4), _p(out), C.byref(oinf))
    return out, int(oinf.value)


def commit(curve, bases_xy, elems_mont, hiding_xy=None, randomizer_mont=None):
    bases_xy, elems_mont = _u64(bases_xy), _u64(elems_mont)
    h = None if hiding_xy is None else _u64(hiding_xy)
    r = None if randomizer_mont is None else _u64(randomizer_mont)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().oracle_commit(C.c_int(curve), _p(bases_xy), C.c_size_t(bases_xy.size // 8), _p(elems_mont),
                        C.c_size_t(elems_mont.size // 4), _p(h), _p(r), _p(out), C.byref(oinf))
    return out, int(oinf.value)


def compute_coeffs(field, challenges_mont):
    ch = _u64(challenges_mont)
    k = ch.size // 4
    out = np.empty((1 << k, 4), dtype=np.uint64)
    lib().oracle_compute_coeffs(C.c_int(field), _p(ch), C.c_int(k), _p(out))
    return out


def succinct_evaluate(field, challenges_mont, z_mont):
    ch, z = _u64(challenges_mont), _u64(z_mont)
    out = np.empty(4, dtype=np.uint64)
    lib().oracle_succinct_evaluate(C.c_int(field), _p(ch), C.c_int(ch.size // 4), _p(z), _p(out))
    return out


def poly_evaluate(field, coeffs_mont, z_mont):
    cf, z = _u64(coeffs_mont), _u64(z_mont)
    out = np.empty(4, dtype=np.uint64)
    lib().oracle_poly_evaluate(C.c_int(field), _p(cf), C.c_size_t(cf.size // 4), _p(z), _p(out))
    return out


def ipa_check_final_key(curve, key_xy, challenges_mont, expected_xy, expected_inf=0):
    key_xy, ch, exp = _u64(key_xy), _u64(challenges_mont), _u64(expected_xy)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    ok = lib().oracle_ipa_check_final_key(C.c_int(curve), _p(key_xy), C.c_size_t(key_xy.size // 8), _p(ch),
                                          C.c_int(ch.size // 4), _p(exp), C.c_uint8(int(expected_inf)),
                                          _p(out), C.byref(oinf))
    return bool(ok), out, int(oinf.value)


def ipa_fold_key(curve, key_xy, challenges_mont):
    key_xy, ch = _u64(key_xy), _u64(challenges_mont)
    out = np.empty(8, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().oracle_ipa_fold_key(C.c_int(curve), _p(key_xy), C.c_size_t(key_xy.size // 8), _p(ch),
                              C.c_int(ch.size // 4), _p(out), C.byref(oinf))
    return out, int(oinf.value)


def powers(field, z_mont, n):
    z = _u64(z_mont)
    out = np.empty((n, 4), dtype=np.uint64)
    lib().oracle_powers(C.c_int(field), _p(z), C.c_size_t(n), _p(out))
    return out


def ipa_open_round_lr(curve, key_xy, coeffs_mont, z_vec_mont, h_prime_xy):
    key_xy, cf, zv, hp = _u64(key_xy), _u64(coeffs_mont), _u64(z_vec_mont), _u64(h_prime_xy)
    n = cf.size // 4
    l, r = np.empty(8, dtype=np.uint64), np.empty(8, dtype=np.uint64)
    li, ri = C.c_uint8(0), C.c_uint8(0)
    lib().oracle_ipa_open_round_lr(C.c_int(curve), _p(key_xy), _p(cf), _p(zv), C.c_size_t(n), _p(hp), _p(l), C.byref(li),
                                   _p(r), C.byref(ri))
    return (l, int(li.value)), (r, int(ri.value))


def ipa_open_fold(curve, key_xy, coeffs_mont, z_vec_mont, xi_mont, xi_inv_mont):
    """in place on copies; returns the folded (key, coeffs, z) halves"""
    key, cf, zv = _u64(key_xy).copy(), _u64(coeffs_mont).copy(), _u64(z_vec_mont).copy()
    n = cf.size // 4
    xi, xinv = _u64(xi_mont), _u64(xi_inv_mont)
    lib().oracle_ipa_open_fold(C.c_int(curve), _p(key), _p(cf), _p(zv), C.c_size_t(n), _p(xi), _p(xinv))
    h = n // 2
    return key.reshape(-1, 8)[:h].copy(), cf.reshape(-1, 4)[:h].copy(), zv.reshape(-1, 4)[:h].copy()


def ipa_succinct_check(curve, comm, z_mont, v_mont, l_xy, r_xy, xi_mont, h_prime_xy, final_key_xy, c_mont) -> bool:
    comm_xy, comm_inf = comm
    l_xy, r_xy, xi = _u64(l_xy), _u64(r_xy), _u64(xi_mont)
    k = xi.size // 4
    return bool(lib().oracle_ipa_succinct_check(C.c_int(curve), _p(_u64(comm_xy)), C.c_uint8(int(comm_inf)), _p(_u64(z_mont)),
                                                _p(_u64(v_mont)), _p(l_xy), _p(r_xy), C.c_int(k), _p(xi), _p(_u64(h_prime_xy)),
                                                _p(_u64(final_key_xy)), _p(_u64(c_mont))))


def combine_check_polys(field, challenges_mont, alphas_mont, random_poly_mont=None):
    ch = _u64(challenges_mont)           # (m, k, 4)
    m, k = ch.shape[0], ch.shape[1]
    al = _u64(alphas_mont)
    rp = None if random_poly_mont is None else _u64(random_poly_mont)
    out = np.empty((1 << k, 4), dtype=np.uint64)
    lib().oracle_combine_check_polys(C.c_int(field), _p(ch), C.c_int(m), C.c_int(k), _p(al), _p(rp),
                                     C.c_size_t(0 if rp is None else rp.size // 4), _p(out))
    return out


def hadamard(field, a, b):
    return _binop("oracle_hadamard", field, a, b)


def scale(field, v, c):
    v, c = _u64(v), _u64(c)
    out = np.empty_like(v)
    lib().oracle_scale(C.c_int(field), _p(v), _p(c), _p(out), C.c_size_t(v.size // 4))
    return out


def _ptr_array(vecs):
    arr = (C.c_void_p * len(vecs))()
    for i, v in enumerate(vecs):
        arr[i] = v.ctypes.data if v.size else None
    return arr


def combine_vectors(field, vecs, challenges, hiding=None):
    vecs = [_u64(v).reshape(-1, 4) for v in vecs]
    ch = _u64(challenges)
    lens = np.array([v.shape[0] for v in vecs], dtype=np.uint64)
    hid = None if hiding is None else _u64(hiding).reshape(-1, 4)
    out_len = max([int(x) for x in lens] + [0 if hid is None else hid.shape[0]])
    out = np.zeros((out_len, 4), dtype=np.uint64)
    lib().oracle_combine_vectors(C.c_int(field), _ptr_array(vecs), _p(lens), C.c_int(len(vecs)), _p(ch), _p(hid),
                                 C.c_size_t(0 if hid is None else hid.shape[0]), _p(out), C.c_size_t(out_len))
    return out


def tvecs(field, a_vecs, b_vecs, mu, length, hiding_a=None, hiding_b=None):
    a_vecs = [_u64(v).reshape(-1, 4) for v in a_vecs]
    b_vecs = [_u64(v).reshape(-1, 4) for v in b_vecs]
    n = len(a_vecs)
    al = np.array([v.shape[0] for v in a_vecs], dtype=np.uint64)
    bl = np.array([v.shape[0] for v in b_vecs], dtype=np.uint64)
    mu = _u64(mu)
    ha = None if hiding_a is None else _u64(hiding_a).reshape(-1, 4)
    hb = None if hiding_b is None else _u64(hiding_b).reshape(-1, 4)
    out = np.zeros((2 * n - 1, length, 4), dtype=np.uint64)
    lib().oracle_tvecs(C.c_int(field), _ptr_array(a_vecs), _p(al), _ptr_array(b_vecs), _p(bl), C.c_int(n), _p(mu),
                       C.c_size_t(length), _p(ha), C.c_size_t(0 if ha is None else ha.shape[0]), _p(hb),
                       C.c_size_t(0 if hb is None else hb.shape[0]), _p(out))
    return out


def csr_matvec(field, row_ptr, cols, coeffs_mont, inp, wit):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    coeffs, inp, wit = _u64(coeffs_mont), _u64(inp), _u64(wit)
    n_rows = row_ptr.size - 1
    out = np.empty((n_rows, 4), dtype=np.uint64)
    lib().oracle_csr_matvec(C.c_int(field), _p(row_ptr), _p(cols), _p(coeffs), C.c_size_t(n_rows), _p(inp),
                            C.c_size_t(inp.size // 4), _p(wit), C.c_size_t(wit.size // 4), _p(out))
    return out


# ---- helpers shared by tests: numpy limbs <-> python ints -----------------------------------------
def to_int(limbs) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(limbs).reshape(-1)[:4]))


def from_int(x: int):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def ints_to_arr(xs):
    return np.array([[(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for x in xs], dtype=np.uint64).reshape(-1, 4)


def arr_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in a]

"""Shared helpers for the parity tests: golden-fixture decoding and oracle-side conversions."""
import json
import os

import numpy as np

from oracle import cref, pyref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, f"{name}.json")) as f:
        return json.load(f)


def ints(hexes):
    return [int(h, 16) for h in hexes]


def fe_mont(field, values):
    """python ints (canonical) -> (n, 4) uint64 Montgomery images"""
    if len(values) == 0:
        return np.zeros((0, 4), dtype=np.uint64)
    return cref.to_mont(field, cref.ints_to_arr(values))


def fe_canon(values):
    if len(values) == 0:
        return np.zeros((0, 4), dtype=np.uint64)
    return cref.ints_to_arr(values)


def fe_ints(field, arr_mont):
    """(n, 4) Montgomery images -> python ints (canonical)"""
    arr = np.asarray(arr_mont, dtype=np.uint64).reshape(-1, 4)
    if arr.shape[0] == 0:
        return []
    return cref.arr_to_ints(cref.from_mont(field, arr))


def points_mont(curve, pts):
    """[[x_hex, y_hex], ...] (canonical) -> (n, 8) uint64 Montgomery"""
    bf = cref.base_field(curve)
    if len(pts) == 0:
        return np.zeros((0, 8), dtype=np.uint64)
    flat = []
    for p in pts:
        flat += [int(p[0], 16), int(p[1], 16)]
    return cref.to_mont(bf, cref.ints_to_arr(flat)).reshape(-1, 8)


def point_result(curve, res):
    """golden result (None or [x_hex, y_hex]) -> (xy Montgomery, inf) in the ark-ec affine image"""
    bf = cref.base_field(curve)
    if res is None:
        ident = np.concatenate([np.zeros(4, dtype=np.uint64), cref.to_mont(bf, cref.from_int(1).reshape(1, 4)).reshape(4)])
        return ident, 1
    return points_mont(curve, [res]).reshape(8), 0


def same_point(a, b):
    return int(a[1]) == int(b[1]) and np.array_equal(np.asarray(a[0], dtype=np.uint64), np.asarray(b[0], dtype=np.uint64))


def scalar_distributions(curve, n, seed):
    """The scalar shapes the reference's own fixtures push through the MSM."""
    sf = cref.scalar_field(curve)
    q = pyref.scalar_modulus(curve)
    one_rand = cref.gen_scalars(sf, seed + 1, 1, True)
    rnd = cref.gen_scalars(sf, seed, n, True)
    trunc = cref.from_mont(sf, rnd).copy()
    trunc[:, 2:] = 0                      # 128-bit challenges (src/hp_as/mod.rs:29)
    dists = {
        "uniform": rnd,
        "constant": np.repeat(one_rand, n, axis=0),          # vec![F::rand(rng); len]  (src/hp_as/mod.rs:991)
        "a_a_a_0": np.concatenate([np.repeat(one_rand, max(n - 1, 0), axis=0), np.zeros((min(n, 1), 4), np.uint64)]),
        "zero": np.zeros((n, 4), dtype=np.uint64),
        "one": cref.to_mont(sf, np.tile(cref.from_int(1), (n, 1))),
        "q_minus_1": cref.to_mont(sf, np.tile(cref.from_int(q - 1), (n, 1))),
        "one_hot": np.concatenate([np.zeros((max(n - 1, 0), 4), np.uint64), one_rand])[:n],
        "trunc128": cref.to_mont(sf, trunc),
    }
    return dists

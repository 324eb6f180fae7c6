"""Timing drivers shaped like the reference's examples (examples/scaling-pc.rs:32-102, scaling-as.rs:38-138,
scaling-nark.rs:58-110): for every log size in [min, max] the hot-path part of each step is timed on the GPU and, up to
--cpu-max, on the host CPU oracle (restated arkworks algorithms), and checked bit for bit.

  pc   : IpaPC commit / open / check at degree 2^k - 1     (BASELINE config 1 is k = 10, config 4 is the check at k = 20)
  hp   : hp-as prove (t-vectors + product-polynomial commitments, 2 inputs) and decide at vector length 2^k   (config 2: k = 16)
  nark : r1cs-nark A z, B z, C z + their three commitments on the scaling-nark circuit with 2^k constraints    (config 3: k = 16)

The host transcript (Poseidon sponge in the reference) is a Blake2s stand-in on both sides; everything that is not on
the hot path (succinct checks, challenge derivation, structure checks) is left out of the timings on both sides.
usage: python examples/scaling.py {pc|hp|nark} <log_min> <log_max> [--cpu-max K]"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab
from accumulation_b200.mirror import CommitterKey, InnerProductArgPC, ASForHadamardProducts, R1CSNark, _fe_to_int, _int_to_fe, _MODULI
from oracle import cref   # checker + CPU column only

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["pc", "hp", "nark"]); ap.add_argument("log_min", type=int); ap.add_argument("log_max", type=int)
ap.add_argument("--cpu-max", type=int, default=14); ap.add_argument("--curve", type=int, default=0)
args = ap.parse_args()
curve = args.curve
sf = ab.scalar_field(curve)
ctx = ab.Context(0)


def squeeze(prev, l, r):
    h = hashlib.blake2s(b"" if prev is None else np.asarray(prev).tobytes())
    h.update(l[0].tobytes()); h.update(r[0].tobytes())
    return _int_to_fe(sf, int.from_bytes(h.digest()[:16], "little") | 1)


def timed(fn, reps=3):
    best, out = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


def same(a, b):
    return int(a[1]) == int(b[1]) and np.array_equal(np.asarray(a[0]), np.asarray(b[0]))


for k in range(args.log_min, args.log_max + 1):
    n = 1 << k
    cpu = k <= args.cpu_max
    rec = {"what": args.what, "log": k}
    bases = ctx.register_synthetic_bases(curve, 0xACC0 + k, n + 1)
    bases.precompute()
    ck = CommitterKey(bases, n)
    pts = ctx.download_bases(bases) if cpu else None
    if args.what == "pc":
        coeffs = cref.gen_scalars(sf, k, n, True)
        z = cref.gen_scalars(sf, 100 + k, 1, True).reshape(4)
        hp = ctx.download_bases(bases, n, 1).reshape(8)
        rec["commit_ms"], comm = timed(lambda: InnerProductArgPC.cm_commit(ck, coeffs))
        xi0 = cref.gen_scalars(sf, 200 + k, 1, True).reshape(4)
        hgen = hp                                                     # the key's hiding generator (base n)
        hp, _ = cref.point_mul(curve, hgen, 0, cref.from_mont(sf, xi0.reshape(1, 4)).reshape(4))   # h' = xi_0 * h (host, like upstream)
        rec["open_ms"], proof = timed(lambda: InnerProductArgPC.open(ck, coeffs, z, None, squeeze, log_d=k, xi0=xi0), reps=2)
        l_vec, r_vec, fk, c, chs = proof
        rec["check_final_key_ms"], ok = timed(lambda: InnerProductArgPC.check_final_key(ck, np.array(chs), fk, 0))
        rec["accept"] = bool(ok)
        if cpu:
            t0 = time.perf_counter(); ecomm = cref.commit(curve, pts[:n], coeffs); rec["cpu_commit_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            t0 = time.perf_counter()
            key_, cf_, zv_, xi = pts[:n], coeffs, cref.powers(sf, z, n), None
            el, er = [], []
            while cf_.shape[0] > 1:
                l, r = cref.ipa_open_round_lr(curve, key_, cf_, zv_, hp)
                xi = squeeze(xi, l, r)
                key_, cf_, zv_ = cref.ipa_open_fold(curve, key_, cf_, zv_, xi, cref.fe_inv(sf, xi.reshape(1, 4)).reshape(4))
                el.append(l); er.append(r)
            rec["cpu_open_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            t0 = time.perf_counter(); eok, _, _ = cref.ipa_check_final_key(curve, pts[:n], np.array(chs), fk, 0); rec["cpu_check_final_key_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            rec["bit_exact"] = bool(same(comm, ecomm) and all(same(a, b) for a, b in zip(l_vec + r_vec, el + er)) and np.array_equal(fk, key_[0]) and eok)
    elif args.what == "hp":
        a = [cref.gen_scalars(sf, 10 * k + i, n, True) for i in range(2)]
        b = [cref.gen_scalars(sf, 10 * k + 5 + i, n, True) for i in range(2)]
        mu = cref.gen_scalars(sf, 7, 3, True)
        rec["prove_tvecs_commit_ms"], (low, high) = timed(lambda: ASForHadamardProducts.compute_t_vecs_and_product_poly_comm(ck, a, b, mu, n))
        r = [cref.gen_scalars(sf, 20 + i, 1, True).reshape(4) for i in range(3)]
        prod = ctx.hadamard(sf, a[0], b[0])
        inst = [ctx.commit(bases, v, hiding_index=n, randomizer_mont=rr) for v, rr in zip((a[0], b[0], prod), r)]
        rec["decide_ms"], ok = timed(lambda: ASForHadamardProducts.decide(ck, inst, (a[0], b[0], tuple(r))))
        rec["accept"] = bool(ok)
        if cpu:
            t0 = time.perf_counter(); t = cref.tvecs(sf, a, b, mu, n); elow, ehigh = cref.commit(curve, pts[:n], t[0]), cref.commit(curve, pts[:n], t[2])
            rec["cpu_prove_tvecs_commit_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            t0 = time.perf_counter(); p2 = cref.hadamard(sf, a[0], b[0]); e = [cref.commit(curve, pts[:n], v, pts[n], rr) for v, rr in zip((a[0], b[0], p2), r)]
            rec["cpu_decide_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            rec["bit_exact"] = bool(same(low[0], elow) and same(high[0], ehigh) and all(same(x, y) for x, y in zip(inst, e)))
    else:
        from tests.test_gpu_fused import scaling_nark_matrices
        mats, n_in, n_wit = scaling_nark_matrices(sf, n)
        t_idx, nark = timed(lambda: R1CSNark(ck, mats), reps=1)
        rec["index_register_matrices_ms"] = t_idx
        inp, wit = cref.gen_scalars(sf, 31, n_in, True), cref.gen_scalars(sf, 32, n_wit, True)
        bl = cref.gen_scalars(sf, 33, 3, True)
        rec["matvec_commit_ms"], (vecs, comms) = timed(lambda: nark.matvec_commit(inp, wit, bl))
        if cpu:
            t0 = time.perf_counter()
            ev = [cref.csr_matvec(sf, *m, inp, wit) for m in mats]
            ec = [cref.commit(curve, pts[:n], v, pts[n], bl[i]) for i, v in enumerate(ev)]
            rec["cpu_matvec_commit_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            rec["bit_exact"] = bool(all(np.array_equal(x, y) for x, y in zip(vecs, ev)) and all(same(x, y) for x, y in zip(comms, ec)))
        nark.release()
    for key_name in list(rec):
        if key_name.endswith("_ms"):
            rec[key_name] = round(rec[key_name], 3)
    print(json.dumps(rec), flush=True)
    bases.release()

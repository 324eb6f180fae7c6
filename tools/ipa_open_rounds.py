"""Development tool: wall-clock of every library call of one IpaPC::open through the one-call-per-round API
(accmsm_ipa_open_fold_round, hiding generator by index -- the flow bench.py and mirror.InnerProductArgPC.open use), for several
materialisation policies.  No per-stage events are read between rounds, so the numbers are those of the production flow.
    python tools/ipa_open_rounds.py 18 5,14 5,11"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab
from accumulation_b200.mirror import _int_to_fe

k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
policies = sys.argv[2:] or ["5,14", "5,11"]
ctx = ab.Context(0)
n = 1 << k
key = ctx.register_synthetic_bases(0, 1, n + 1)
key.precompute()
rng = np.random.default_rng(1)
coeffs = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); coeffs[:, 3] &= np.uint64((1 << 62) - 1)
z, xi0 = coeffs[0].copy(), coeffs[1].copy()

def squeeze(prev, l, r):
    h = hashlib.blake2s(b"" if prev is None else prev.tobytes())
    h.update(l[0].tobytes()); h.update(r[0].tobytes())
    return _int_to_fe(1, int.from_bytes(h.digest()[:16], "little") | 1)

for pol in policies:
    ctx.set_ipa_fold(*[int(v) for v in pol.split(",")])
    best = None
    for rep in range(3):
        t = [time.perf_counter()]
        sess = ctx.ipa_open_begin(key, coeffs, k, z, None)
        ctx.ipa_open_use_hiding_generator(sess, n, xi0)
        t.append(time.perf_counter())
        xi, lr = None, ctx.ipa_open_round(sess)
        t.append(time.perf_counter())
        host = 0.0
        while lr is not None:
            h0 = time.perf_counter()
            xi = squeeze(xi, lr[0], lr[1])
            host += time.perf_counter() - h0
            lr = ctx.ipa_open_fold_round(sess, xi)
            t.append(time.perf_counter())
        ctx.ipa_open_finish(sess)
        t.append(time.perf_counter())
        ms = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
        if best is None or sum(ms) < sum(best[0]):
            best = (ms, host * 1e3)
    ms, host = best
    print(f"k={k} fold={pol}: total {sum(ms):.3f} ms (host sponge stand-in {host:.3f} ms); begin {ms[0]:.3f}; calls: " + " ".join(f"{x:.2f}" for x in ms[1:]), flush=True)

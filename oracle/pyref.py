"""TEST INFRASTRUCTURE ONLY -- independent big-int oracle for the accmsm hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product path (`accumulation_b200/`) must never import it.

PARITY UNPINNED at the reference boundary: arkworks-rs/accumulation holds no golden vectors,
KATs or fixtures for this path and its arithmetic lives in un-vendored
crates (ark-ec/ark-ff/ark-poly 0.2.0, ark-poly-commit@accumulation-experimental).  What pins
this oracle instead: the curve KATs of SURVEY.md App. B (G=(-1,2), 2G, q*G = O), the algebraic
identities of section 8c(4), and bit-agreement between this file (affine formulas, Python ints,
`pow(x,-1,p)`) and the C restatement `oracle/oracle.c` (Jacobian formulas, 4x64 Montgomery).

Everything here is plain Python integers in CANONICAL (non-Montgomery) form unless a function
name says `mont`.  Pure-Python loops: use for small cases only.
"""
from __future__ import annotations

# --- SURVEY.md App. B constants -------------------------------------------------------------
P_PALLAS_BASE = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Fp
Q_PALLAS_SCALAR = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Fq
R_MONT = 1 << 256
CURVE_B = 5
PALLAS, VESTA = 0, 1


def base_modulus(curve: int) -> int:
    return P_PALLAS_BASE if curve == PALLAS else Q_PALLAS_SCALAR


def scalar_modulus(curve: int) -> int:
    return Q_PALLAS_SCALAR if curve == PALLAS else P_PALLAS_BASE


def to_mont(x: int, m: int) -> int:
    return (x * R_MONT) % m


def from_mont(x: int, m: int) -> int:
    return (x * pow(R_MONT, -1, m)) % m


def limbs4(x: int):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs4(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# --- affine group law, identity = None -------------------------------------------------------
def generator(curve: int):
    m = base_modulus(curve)
    return (m - 1, 2)


def on_curve(pt, curve: int) -> bool:
    if pt is None:
        return True
    m = base_modulus(curve)
    x, y = pt
    return (y * y - x * x * x - CURVE_B) % m == 0


def neg(pt, curve: int):
    if pt is None:
        return None
    m = base_modulus(curve)
    return (pt[0], (-pt[1]) % m)


def add(p1, p2, curve: int):
    m = base_modulus(curve)
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % m == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, m) % m
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, m) % m
    x3 = (lam * lam - x1 - x2) % m
    y3 = (lam * (x1 - x3) - y1) % m
    return (x3, y3)


def mul(k: int, pt, curve: int):
    acc = None
    addend = pt
    while k > 0:
        if k & 1:
            acc = add(acc, addend, curve)
        addend = add(addend, addend, curve)
        k >>= 1
    return acc


def msm_naive(bases, scalars, curve: int):
    """sum s_i * P_i with ark-ec's truncation to min(len). Canonical ints."""
    acc = None
    for pt, s in zip(bases, scalars):
        acc = add(acc, mul(s, pt, curve), curve)
    return acc


def msm_bucket(bases, scalars, curve: int, c: int = 8):
    """Faster independent MSM (unsigned windows, affine buckets) for mid-size golden vectors."""
    n = min(len(bases), len(scalars))
    nwin = (255 + c - 1) // c
    total = None
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = add(total, total, curve)
        buckets = [None] * (1 << c)
        for i in range(n):
            d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = add(buckets[d], bases[i], curve)
        run = None
        acc = None
        for d in range((1 << c) - 1, 0, -1):
            run = add(run, buckets[d], curve)
            acc = add(acc, run, curve)
        total = add(total, acc, curve)
    return total


def sqrt_mod(a: int, m: int):
    """Tonelli-Shanks (two-adicity 32 for both moduli). Returns None for non-residues."""
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    s, t = 0, m - 1
    while t % 2 == 0:
        s += 1
        t //= 2
    z = 2
    while pow(z, (m - 1) // 2, m) == 1:
        z += 1
    c = pow(z, t, m)
    x = pow(a, (t + 1) // 2, m)
    b = pow(a, t, m)
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % m
            i += 1
        e = pow(c, 1 << (s - i - 1), m)
        x = x * e % m
        c = e * e % m
        b = b * c % m
        s = i
    return x


class SplitMix64:
    """Seeded generator used by every synthetic input in tests/bench."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def field(self, m: int) -> int:
        """Rejection-sample a 255-bit value < m (mask the top bit like ark-ff's UniformRand)."""
        while True:
            v = 0
            for i in range(4):
                v |= self.next() << (64 * i)
            v &= (1 << 255) - 1
            if v < m:
                return v


def random_point(rng: SplitMix64, curve: int):
    m = base_modulus(curve)
    while True:
        x = rng.field(m)
        y = sqrt_mod((x * x * x + CURVE_B) % m, m)
        if y is None:
            continue
        if rng.next() & 1:
            y = (-y) % m
        return (x, y)


# --- SuccinctCheckPolynomial (reference call sites src/ipa_pc_as/mod.rs:400,418)
def compute_coeffs(challenges, m: int):
    k = len(challenges)
    coeffs = [1] * (1 << k)
    for i, xi in enumerate(challenges, start=1):
        e = 1 << (k - i)
        for start in range(e, 1 << k, 2 * e):
            for j in range(start, start + e):
                coeffs[j] = coeffs[j] * xi % m
    return coeffs


def succinct_evaluate(challenges, z: int, m: int) -> int:
    k = len(challenges)
    out = 1
    for i, xi in enumerate(challenges, start=1):
        out = out * (1 + xi * pow(z, 1 << (k - i), m)) % m
    return out


def horner(coeffs, z: int, m: int) -> int:
    acc = 0
    for cf in reversed(coeffs):
        acc = (acc * z + cf) % m
    return acc


# --- hp_as vector ops (src/hp_as/mod.rs:278-349, 482-512) ------------------------------------
def compute_hp(a, b, m):
    return [(x * y) % m for x, y in zip(a, b)]


def compute_t_vecs(a_vecs, b_vecs, mu, length, m, hiding=None):
    n = len(a_vecs)
    t = [[0] * length for _ in range(2 * n - 1)]
    for li in range(length):
        ac = [(mu[i] * a_vecs[i][li]) % m if li < len(a_vecs[i]) else 0 for i in range(n)]
        bc = [b_vecs[i][li] if li < len(b_vecs[i]) else 0 for i in range(n)]
        bc.reverse()
        if hiding is not None:
            ha, hb = hiding
            if li < len(ha):
                ac[0] = (ac[0] + ha[li] * mu[n]) % m
            if li < len(hb):
                bc[0] = (bc[0] + hb[li] * mu[1]) % m
        for i in range(n):
            for j in range(n):
                t[i + j][li] = (t[i + j][li] + ac[i] * bc[j]) % m
    return t


def combine_vectors(vectors, challenges, m, hiding=None):
    out = list(hiding) if hiding is not None else []
    for ni, v in enumerate(vectors):
        for li, e in enumerate(v):
            prod = challenges[ni] * e % m
            if li >= len(out):
                out.append(prod)
            else:
                out[li] = (out[li] + prod) % m
    return out


def scale_vector(v, c, m):
    return [(x * c) % m for x in v]


# --- r1cs_nark matrix_vec_mul (src/r1cs_nark_as/r1cs_nark/mod.rs:443-462) --------------------
def matrix_vec_mul(rows, inp, wit, m):
    out = []
    for row in rows:
        acc = 0
        for coeff, col in row:
            z = inp[col] if col < len(inp) else wit[col - len(inp)]
            acc = (acc + coeff * z) % m
        out.append(acc)
    return out


# --- IPA key folding identity ----------------------------------------------
def fold_key(key, challenges, curve: int):
    """key_l += xi * key_r per round; final_comm_key == MSM(key, compute_coeffs(challenges))."""
    key = list(key)
    for xi in challenges:
        h = len(key) // 2
        key = [add(key[i], mul(xi, key[i + h], curve), curve) for i in range(h)]
    assert len(key) == 1
    return key[0]


# --- IpaPC::open / succinct_check (src/ipa_pc_as/mod.rs:198-205,454-462) ------
def ipa_open(key, coeffs, z: int, h_prime, challenges, curve: int):
    """k rounds with the round challenges given (the host sponge squeezes them from (l, r) in the real
    protocol).  Returns (l_vec, r_vec, final_comm_key, c)."""
    q = scalar_modulus(curve)
    key, a = list(key), [c % q for c in coeffs]
    b = [pow(z, i, q) for i in range(len(a))]
    l_vec, r_vec = [], []
    for xi in challenges:
        h = len(a) // 2
        ip_l = sum(x * y for x, y in zip(a[h:], b[:h])) % q
        ip_r = sum(x * y for x, y in zip(a[:h], b[h:])) % q
        l_vec.append(add(msm_naive(key[:h], a[h:], curve), mul(ip_l, h_prime, curve), curve))
        r_vec.append(add(msm_naive(key[h:], a[:h], curve), mul(ip_r, h_prime, curve), curve))
        xinv = pow(xi, -1, q)
        a = [(a[i] + xinv * a[i + h]) % q for i in range(h)]
        b = [(b[i] + xi * b[i + h]) % q for i in range(h)]
        key = [add(key[i], mul(xi, key[i + h], curve), curve) for i in range(h)]
    return l_vec, r_vec, key[0], a[0]


def ipa_succinct_check(comm, z: int, v: int, l_vec, r_vec, challenges, h_prime, final_key, c: int, curve: int) -> bool:
    q = scalar_modulus(curve)
    acc = add(comm, mul(v % q, h_prime, curve), curve)
    for xi, l, r in zip(challenges, l_vec, r_vec):
        acc = add(acc, add(mul(pow(xi, -1, q), l, curve), mul(xi, r, curve), curve), curve)
    v_prime = succinct_evaluate(challenges, z, q) * c % q
    return acc == add(mul(c, final_key, curve), mul(v_prime, h_prime, curve), curve)

"""Wire / seed compatibility with the arkworks 0.2 stack the reference builds on.

Host-side formats only -- byte shuffling, no field arithmetic on the hot path:

* ark-serialize 0.2 `CanonicalSerialize` images of the types that cross the commitment boundary
  (derives at src/ipa_pc_as/data_structures.rs:55,76, src/hp_as/data_structures.rs:13,53,94):
  field element 32 B little-endian canonical; short-Weierstrass affine point compressed 33 B (x + flag byte: bit 7 =
  y is the larger root, bit 6 = infinity) / uncompressed 65 B (x, then y with the infinity flag); `Vec<T>` = u64 LE
  length + items; `Option<T>` = one byte + item; `ipa_pc::Commitment{comm, shifted_comm}`,
  `ipa_pc::Proof{l_vec, r_vec, final_comm_key, c, hiding_comm, rand}` field by field in declaration order.
* `TestRng`: `ark_std::test_rng()` = `rand::rngs::StdRng::from_seed([1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0 x 16])`.
  With rand 0.7 (ark-std 0.2) StdRng is ChaCha20 (rand_chacha 0.2: 64-bit block counter, stream 0, 4-block buffer,
  `next_u64` = two consecutive little-endian words); `rounds=12` gives rand 0.8's StdRng.
* `rand_fe` / `rand_point`: ark-ff 0.2 `UniformRand` for `Fp256` (4 x next_u64, top REPR_SHAVE_BITS = 1 bit masked,
  rejection against the modulus, the accepted integer IS the Montgomery image) and ark-ec 0.2 `GroupProjective::rand`
  (x = Fq::rand, greatest = rng.gen::<bool>(), `get_point_from_x`, cofactor 1).

The large direction of the key format -- decompressing n x 33 B into an HBM-resident key -- runs on the device:
`Context.register_bases_compressed` (csrc/wire.cuh).  Everything here follows the published sources of the pinned
dependency versions as recalled offline; `tools/ref_fixtures/ (cargo run --release)` is the recipe that produces fixtures from a real
arkworks build, and `tests/test_ref_fixtures.py` replays them when they are present.  Until such fixtures exist the
parity of this module with arkworks is UNPINNED (the ChaCha20 core alone is pinned, by RFC 8439's block test vector).
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

P_BASE = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001      # Pallas base field = Vesta scalar field
Q_SCALAR = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001    # Pallas scalar field = Vesta base field
MODULI = (P_BASE, Q_SCALAR)            # field ids as everywhere: 0 = Fp, 1 = Fq
R = 1 << 256
FLAG_POSITIVE_Y, FLAG_INFINITY = 1 << 7, 1 << 6


def base_field(curve: int) -> int:
    return 0 if curve == 0 else 1


# ---------------------------------------------------------------------------------------------------------------
# limbs <-> integers (Montgomery memory image <-> canonical value)
# ---------------------------------------------------------------------------------------------------------------
def mont_limbs_to_int(field: int, limbs) -> int:
    m = MODULI[field]
    v = sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(limbs, dtype=np.uint64).reshape(4)))
    return v * pow(R, -1, m) % m


def int_to_mont_limbs(field: int, value: int) -> np.ndarray:
    v = value % MODULI[field] * R % MODULI[field]
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


# ---------------------------------------------------------------------------------------------------------------
# CanonicalSerialize
# ---------------------------------------------------------------------------------------------------------------
def ser_fe(value: int) -> bytes:
    return int(value).to_bytes(32, "little")


def de_fe(field: int, data: bytes, pos: int = 0) -> Tuple[int, int]:
    v = int.from_bytes(data[pos:pos + 32], "little")
    if v >= MODULI[field]:
        raise ValueError("field element not canonical")
    return v, pos + 32


def _sqrt(field: int, a: int) -> Optional[int]:
    """Tonelli-Shanks (two-adicity 32); either root"""
    m = MODULI[field]
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    t = (m - 1) >> 32
    z = pow(5, t, m)
    w = pow(a, (t - 1) // 2, m)
    x, b, v = a * w % m, a * w * w % m, 32
    while b != 1:
        k, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % m
            k += 1
        tt = pow(z, 1 << (v - k - 1), m)
        z, b, x, v = tt * tt % m, b * tt * tt % m, x * tt % m, k
    return x


def point_from_x(curve: int, x: int, greatest: bool) -> Optional[Tuple[int, int]]:
    """ark-ec `GroupAffine::get_point_from_x`: the larger (greatest) or smaller root of y^2 = x^3 + 5"""
    f = base_field(curve)
    m = MODULI[f]
    y = _sqrt(f, (x * x * x + 5) % m)
    if y is None:
        return None
    ny = (m - y) % m
    return (x, y if (y < ny) != greatest else ny)


def ser_point_compressed(curve: int, pt: Optional[Tuple[int, int]]) -> bytes:
    """pt = (x, y) canonical integers, or None for the point at infinity -> 33 bytes"""
    if pt is None:
        return bytes(32) + bytes([FLAG_INFINITY])
    m = MODULI[base_field(curve)]
    x, y = pt
    return ser_fe(x) + bytes([FLAG_POSITIVE_Y if y > (m - y) % m else 0])


def de_point_compressed(curve: int, data: bytes, pos: int = 0) -> Tuple[Optional[Tuple[int, int]], int]:
    flags = data[pos + 32]
    if flags & 0x3F or (flags & FLAG_POSITIVE_Y and flags & FLAG_INFINITY):
        raise ValueError("bad point flags")
    x, _ = de_fe(base_field(curve), data, pos)
    if flags & FLAG_INFINITY:
        if x:
            raise ValueError("infinity with non-zero x")
        return None, pos + 33
    pt = point_from_x(curve, x, bool(flags & FLAG_POSITIVE_Y))
    if pt is None:
        raise ValueError("x is not on the curve")
    return pt, pos + 33


def ser_point_uncompressed(curve: int, pt: Optional[Tuple[int, int]]) -> bytes:
    if pt is None:
        return bytes(64) + bytes([FLAG_INFINITY])
    return ser_fe(pt[0]) + ser_fe(pt[1]) + bytes([0])


def de_point_uncompressed(curve: int, data: bytes, pos: int = 0) -> Tuple[Optional[Tuple[int, int]], int]:
    f = base_field(curve)
    x, _ = de_fe(f, data, pos)
    y, _ = de_fe(f, data, pos + 32)
    flags = data[pos + 64]
    if flags & FLAG_INFINITY:
        return None, pos + 65
    if (y * y - x * x * x - 5) % MODULI[f]:
        raise ValueError("point not on the curve")
    return (x, y), pos + 65


def ser_vec(items: Sequence[bytes]) -> bytes:
    return struct.pack("<Q", len(items)) + b"".join(items)


def de_len(data: bytes, pos: int) -> Tuple[int, int]:
    return struct.unpack_from("<Q", data, pos)[0], pos + 8


def ser_option(item: Optional[bytes]) -> bytes:
    return b"\x00" if item is None else b"\x01" + item


def ser_ipa_commitment(curve: int, comm, shifted_comm=None) -> bytes:
    """ipa_pc::Commitment { comm: G, shifted_comm: Option<G> }"""
    return ser_point_compressed(curve, comm) + ser_option(None if shifted_comm is None else ser_point_compressed(curve, shifted_comm))


def ser_ipa_proof(curve: int, l_vec, r_vec, final_comm_key, c: int, hiding_comm=None, rand: Optional[int] = None) -> bytes:
    """ipa_pc::Proof { l_vec: Vec<G>, r_vec: Vec<G>, final_comm_key: G, c: G::ScalarField, hiding_comm: Option<G>,
    rand: Option<G::ScalarField> } (fields read at src/ipa_pc_as/mod.rs:217,587)"""
    return (ser_vec([ser_point_compressed(curve, p) for p in l_vec]) + ser_vec([ser_point_compressed(curve, p) for p in r_vec]) +
            ser_point_compressed(curve, final_comm_key) + ser_fe(c) +
            ser_option(None if hiding_comm is None else ser_point_compressed(curve, hiding_comm)) +
            ser_option(None if rand is None else ser_fe(rand)))


def de_ipa_proof(curve: int, data: bytes):
    sf = 1 - base_field(curve)
    pos = 0
    vecs = []
    for _ in range(2):
        n, pos = de_len(data, pos)
        v = []
        for _ in range(n):
            pt, pos = de_point_compressed(curve, data, pos)
            v.append(pt)
        vecs.append(v)
    fk, pos = de_point_compressed(curve, data, pos)
    c, pos = de_fe(sf, data, pos)
    hiding = rand = None
    if data[pos]:
        hiding, pos = de_point_compressed(curve, data, pos + 1)
    else:
        pos += 1
    if data[pos]:
        rand, pos = de_fe(sf, data, pos + 1)
    else:
        pos += 1
    if pos != len(data):
        raise ValueError("trailing bytes")
    return vecs[0], vecs[1], fk, c, hiding, rand


def key_to_compressed(curve: int, xy_mont) -> bytes:
    """(n, 8) uint64 Montgomery x || y (the memory image the C-ABI takes) -> n x 33 B"""
    f = base_field(curve)
    out = bytearray()
    for row in np.asarray(xy_mont, dtype=np.uint64).reshape(-1, 8):
        out += ser_point_compressed(curve, (mont_limbs_to_int(f, row[:4]), mont_limbs_to_int(f, row[4:])))
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------
# ark_std::test_rng()
# ---------------------------------------------------------------------------------------------------------------
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha_block(key_words: Sequence[int], counter: int, stream_words: Sequence[int] = (0, 0), rounds: int = 20,
                 state12_15: Optional[Sequence[int]] = None) -> List[int]:
    """one 64-byte ChaCha block as 16 little-endian words.  Words 12, 13 = 64-bit block counter, 14, 15 = stream id
    (djb's layout, what rand_chacha uses); state12_15 overrides all four (RFC 8439's 32-bit counter + 96-bit nonce)."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words)
    st += list(state12_15) if state12_15 is not None else [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream_words[0], stream_words[1]]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, st)]


TEST_RNG_SEED = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)


class TestRng:
    """rand's StdRng (BlockRng over a ChaCha core with a 4-block buffer) seeded like ark_std::test_rng()"""
    __test__ = False        # not a pytest class

    def __init__(self, seed: bytes = TEST_RNG_SEED, rounds: int = 20):
        self.key = list(struct.unpack("<8I", seed))
        self.rounds = rounds
        self.counter = 0
        self.buf: List[int] = []
        self.index = 64          # empty buffer

    def _generate(self):
        self.buf = []
        for b in range(4):
            self.buf += chacha_block(self.key, self.counter + b, rounds=self.rounds)
        self.counter += 4

    def next_u32(self) -> int:
        if self.index >= 64:
            self._generate(); self.index = 0
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self) -> int:
        """rand_core BlockRng::next_u64: two consecutive words, low first; an odd word left at the end of the buffer is
        paired with the first word of the refill"""
        if self.index < 63:
            lo, hi = self.buf[self.index], self.buf[self.index + 1]
            self.index += 2
            return (hi << 32) | lo
        if self.index >= 64:
            self._generate()
            self.index = 2
            return (self.buf[1] << 32) | self.buf[0]
        lo = self.buf[63]
        self._generate()
        self.index = 1
        return (self.buf[0] << 32) | lo

    def gen_bool(self) -> bool:
        """rand 0.7 `Standard` for bool: the sign bit of next_u32"""
        return bool(self.next_u32() >> 31)


def rand_fe_mont_image(rng: TestRng, field: int) -> int:
    """ark-ff 0.2 `UniformRand for Fp256`: the accepted 255-bit integer, which arkworks uses AS the Montgomery image"""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= 0xFFFFFFFFFFFFFFFF >> 1              # REPR_SHAVE_BITS = 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < MODULI[field]:
            return v


def rand_fe(rng: TestRng, field: int) -> int:
    """canonical value of `F::rand(rng)`"""
    return rand_fe_mont_image(rng, field) * pow(R, -1, MODULI[field]) % MODULI[field]


def rand_point(rng: TestRng, curve: int) -> Tuple[int, int]:
    """ark-ec 0.2 `GroupProjective::rand` (cofactor 1): x = Fq::rand, greatest = bool, first x that is on the curve"""
    f = base_field(curve)
    while True:
        x = rand_fe(rng, f)
        greatest = rng.gen_bool()
        pt = point_from_x(curve, x, greatest)
        if pt is not None:
            return pt

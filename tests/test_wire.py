"""Wire / seed compatibility: ark-serialize images and the `test_rng()` generator.

CPU suite: the ChaCha20 core against RFC 8439's block test vector, the BlockRng word pairing, ark-ff's UniformRand rule,
round trips of every serialised type, rejection of invalid encodings.  GPU suite: a compressed commitment key is
decompressed on the device (Tonelli-Shanks per point) bit-exactly, also through a device group, and keys serialise back.
Parity with arkworks itself stays unpinned until fixtures from tools/ref_fixtures/ (cargo run --release) exist (tests/test_ref_fixtures.py)."""
import numpy as np
import pytest

from accumulation_b200 import wire
from oracle import cref


def test_chacha20_block_rfc8439_vector():
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    out = wire.chacha_block(key, 0, state12_15=[1, 0x09000000, 0x4A000000, 0])
    exp = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
           0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert out == exp


def test_block_rng_word_pairing_and_refill():
    a, b = wire.TestRng(), wire.TestRng()
    words = [a.next_u32() for _ in range(200)]
    # u64 = (word[2i+1] << 32) | word[2i]; a u32 in between shifts the pairing by one word, also across the 64-word refill
    assert [b.next_u64() for _ in range(10)] == [(words[2 * i + 1] << 32) | words[2 * i] for i in range(10)]
    assert b.next_u32() == words[20]
    got = [b.next_u64() for _ in range(40)]
    assert got == [(words[22 + 2 * i] << 32) | words[21 + 2 * i] for i in range(40)]
    # the stream is the concatenation of consecutive ChaCha20 blocks with a 64-bit counter starting at 0
    key = list(np.frombuffer(wire.TEST_RNG_SEED, dtype="<u4"))
    assert words[:16] == wire.chacha_block([int(k) for k in key], 0) and words[64:80] == wire.chacha_block([int(k) for k in key], 4)
    assert wire.TestRng(rounds=12).next_u64() != wire.TestRng().next_u64()


@pytest.mark.parametrize("field", [0, 1])
def test_uniform_rand_rule(field):
    """4 x next_u64, top bit masked, rejected if >= modulus, accepted value is the Montgomery IMAGE"""
    rng, raw = wire.TestRng(), wire.TestRng()
    m = wire.MODULI[field]
    for _ in range(50):
        v = wire.rand_fe(rng, field)
        while True:
            limbs = [raw.next_u64() for _ in range(4)]
            img = sum(l << (64 * i) for i, l in enumerate(limbs)) & ((1 << 255) - 1)
            if img < m:
                break
        assert v == img * pow(1 << 256, -1, m) % m
        assert wire.mont_limbs_to_int(field, wire.int_to_mont_limbs(field, v)) == v
        assert sum(int(x) << (64 * i) for i, x in enumerate(wire.int_to_mont_limbs(field, v))) == img


@pytest.mark.parametrize("curve", [0, 1])
def test_point_and_proof_round_trips(curve):
    rng = wire.TestRng()
    f, sf = wire.base_field(curve), 1 - wire.base_field(curve)
    m = wire.MODULI[f]
    pts = [wire.rand_point(rng, curve) for _ in range(12)]
    for x, y in pts:
        assert (y * y - x * x * x - 5) % m == 0
    for pt in pts + [None]:
        c, u = wire.ser_point_compressed(curve, pt), wire.ser_point_uncompressed(curve, pt)
        assert len(c) == 33 and len(u) == 65
        assert wire.de_point_compressed(curve, c) == (pt, 33) and wire.de_point_uncompressed(curve, u) == (pt, 65)
    x, y = pts[0]
    assert wire.ser_point_compressed(curve, (x, y))[32] != wire.ser_point_compressed(curve, (x, m - y))[32]     # the sign flag
    proof = wire.ser_ipa_proof(curve, pts[:4], pts[4:8], pts[8], wire.rand_fe(rng, sf), hiding_comm=pts[9], rand=wire.rand_fe(rng, sf))
    assert len(proof) == 2 * (8 + 4 * 33) + 33 + 32 + 34 + 33
    l, r, fk, c, h, rd = wire.de_ipa_proof(curve, proof)
    assert l == pts[:4] and r == pts[4:8] and fk == pts[8] and h == pts[9]
    assert wire.ser_ipa_proof(curve, l, r, fk, c, h, rd) == proof
    bare = wire.ser_ipa_proof(curve, pts[:2], pts[2:4], pts[4], 7)
    assert bare[-2:] == b"\x00\x00" and wire.de_ipa_proof(curve, bare)[4:] == (None, None)
    assert len(wire.ser_ipa_commitment(curve, pts[0])) == 34 and len(wire.ser_ipa_commitment(curve, pts[0], pts[1])) == 67


@pytest.mark.parametrize("curve", [0, 1])
def test_invalid_encodings_are_rejected(curve):
    f = wire.base_field(curve)
    m = wire.MODULI[f]
    with pytest.raises(ValueError):
        wire.de_fe(f, wire.ser_fe(m))                                                   # not canonical
    x = next(x for x in range(2, 200) if wire.point_from_x(curve, x, True) is None)      # x^3 + 5 is a non-residue
    with pytest.raises(ValueError):
        wire.de_point_compressed(curve, wire.ser_fe(x) + b"\x00")
    good = wire.ser_point_compressed(curve, wire.rand_point(wire.TestRng(), curve))
    for flags in (0xC0, 0x01, 0x20):
        with pytest.raises(ValueError):
            wire.de_point_compressed(curve, good[:32] + bytes([flags]))
    with pytest.raises(ValueError):
        wire.de_point_compressed(curve, wire.ser_fe(3) + bytes([wire.FLAG_INFINITY]))


def test_key_image_matches_the_oracle_points():
    """the oracle's seeded points, through the Montgomery memory image and back"""
    for curve in (0, 1):
        pts = cref.gen_points(curve, 9, 20)
        data = wire.key_to_compressed(curve, pts)
        f = wire.base_field(curve)
        for i in range(20):
            pt, _ = wire.de_point_compressed(curve, data, 33 * i)
            assert np.array_equal(wire.int_to_mont_limbs(f, pt[0]), pts[i, :4]) and np.array_equal(wire.int_to_mont_limbs(f, pt[1]), pts[i, 4:])


# ---------------------------------------------------------------------------------------------------------------
# GPU: decompression on the device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("curve", [0, 1])
def test_compressed_key_registration_on_device(ctx, curve):
    from tests.util import same_point
    n = 3000
    pts = cref.gen_points(curve, 31 + curve, n)
    data = wire.key_to_compressed(curve, pts)
    B = ctx.register_bases_compressed(curve, data)
    assert B.n == n and np.array_equal(ctx.download_bases(B), pts)                    # both roots chosen by the flag, bit-exact
    assert ctx.serialize_bases(B) == data and ctx.serialize_bases(B, 100, 50) == data[3300:4950]
    sc = cref.gen_scalars(cref.scalar_field(curve), 5, n, True)
    assert same_point(ctx.msm(B, sc), cref.commit(curve, pts, sc))
    B.release()
    # a key with identity entries: they deserialise, contribute nothing, and serialise back
    with_inf = bytearray(data[:33 * 10]); with_inf[33 * 3:33 * 4] = wire.ser_point_compressed(curve, None)
    B = ctx.register_bases_compressed(curve, bytes(with_inf))
    sc10 = sc[:10].copy()
    keep = [i for i in range(10) if i != 3]
    assert same_point(ctx.msm(B, sc10), cref.commit(curve, pts[keep], sc10[keep]))
    assert ctx.serialize_bases(B) == bytes(with_inf)
    B.release()
    assert ctx.register_bases_compressed(curve, b"").n == 0


@pytest.mark.gpu
@pytest.mark.parametrize("curve", [0, 1])
def test_compressed_key_invalid_encodings_fail_the_call(ctx, curve):
    import accumulation_b200 as ab
    f = wire.base_field(curve)
    good = wire.key_to_compressed(curve, cref.gen_points(curve, 3, 8))
    x_bad = next(x for x in range(2, 200) if wire.point_from_x(curve, x, True) is None)
    for bad_rec in (wire.ser_fe(x_bad) + b"\x00", (wire.MODULI[f] + 1).to_bytes(32, "little") + b"\x00", good[:32] + b"\xc0",
                    good[:32] + b"\x01", wire.ser_fe(3) + bytes([wire.FLAG_INFINITY])):
        data = bytearray(good); data[33 * 5:33 * 6] = bad_rec
        with pytest.raises(ab.AccmsmError):
            ctx.register_bases_compressed(curve, bytes(data))


@pytest.mark.gpu
def test_compressed_key_on_a_device_group():
    import accumulation_b200 as ab
    from tests.util import same_point
    g = ab.Context(devices=[0, 0, 0], min_shard=16)
    try:
        pts = cref.gen_points(0, 77, 1000)
        data = wire.key_to_compressed(0, pts)
        B = g.register_bases_compressed(0, data)
        assert np.array_equal(g.download_bases(B), pts) and g.serialize_bases(B) == data
        sc = cref.gen_scalars(cref.FQ, 6, 1000, True)
        assert same_point(g.msm(B, sc), cref.commit(0, pts, sc))
        B.release()
    finally:
        g.close()

"""Pins the CPU oracle: the reference's own golden table + the known answers of SURVEY.md Appendix B.
(The reference ships no other fixtures: .gitignore:2-4 excludes *.wav/*.dat, `make test` asserts nothing.)"""
import ctypes as C
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_freezer_reproduces_reference_polar_tables(oracle):
    gold = json.load(open(os.path.join(HERE, "golden", "polar_tables.sha256")))
    for table, name in ((0, "frozen_64800_43072"), (1, "frozen_64512_43072")):
        fr = np.zeros(2048, np.uint32)
        oracle.lib().ref_frozen_table(table, _p(fr))
        assert hashlib.sha256(fr.astype("<u4").tobytes()).hexdigest() == gold[name]["sha256"]
        assert sum(bin(int(x)).count("1") for x in fr) == gold[name]["frozen"]


def test_frozen_table_structure(oracle):
    fr = np.zeros(2048, np.uint32)
    oracle.lib().ref_frozen_table(0, _p(fr))
    bits = np.unpackbits(fr.view(np.uint8), bitorder="little")
    assert bits.sum() == 21728 and (1 - bits).sum() == 43808          # mesg_bits mode 6 (decode.cc:310)
    assert int(np.argmin(bits)) == 4063                                # first free index
    assert int(np.nonzero(bits)[0][-1]) == 61440                       # last frozen index
    assert (fr == 0xFFFFFFFF).sum() == 438 and (fr == 0).sum() == 914
    per_block = bits.reshape(16, 4096).sum(axis=1).tolist()
    assert per_block == [4089, 3676, 3394, 1547, 2939, 1075, 810, 110, 2374, 665, 501, 75, 385, 50, 37, 1]
    assert bits[64800:].sum() == 0                                      # lengthen(): the padded tail is non-frozen


def test_mls(oracle):
    for poly, first, period, ones in ((0b10001001, "0000001000100110", 127, 64), (0b100101011, "0000000100101111", 255, 128),
                                      (0b100101010001, "0000000000100101", 2047, 1024)):
        out = np.zeros(2 * period, np.uint8)
        oracle.lib().ref_mls(poly, 2 * period, _p(out))
        assert "".join(map(str, out[:16])) == first
        assert (out[:period] == out[period:]).all() and out[:period].sum() == ones


def test_xorshift_crc_base37(oracle):
    L = oracle.lib()
    x = np.zeros(6, np.uint32)
    L.ref_xorshift(6, _p(x))
    assert x.tolist() == [723471715, 2497366906, 2064144800, 2008045182, 3532304609, 374114282]
    assert L.ref_base37(b"CALLSIGN") == 1263905687425 and L.ref_base37(b"ANONYMOUS") == 40981513255571
    assert 37 ** 9 == 129961739795077
    md = (1263905687425 << 8) | 6
    assert md == 323559855980806
    assert L.ref_crc16_u64(md << 9) == 0xE724
    # CRC-32 residue over data || crc (LSB first) is zero
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, 5380, dtype=np.uint8)
    crc = L.ref_crc32_bytes(_p(data), 5380)
    bits = np.concatenate([np.unpackbits(data, bitorder="little"), np.array([(crc >> i) & 1 for i in range(32)], np.uint8)])
    assert L.ref_crc32_bits(_p(bits), bits.size) == 0


def test_bch_generator(oracle):
    g = np.zeros(185, np.uint8)
    oracle.lib().ref_bch_generator(_p(g))
    assert g[184] == 1 and g[0] == 1                                   # degree 184
    # g(x) divides x^255 + 1
    rem = np.zeros(256, np.uint8)
    rem[255] = 1
    rem[0] ^= 1
    for d in range(255, 183, -1):
        if rem[d]:
            rem[d - 184:d + 1] ^= g
    assert not rem.any()
    # systematic generator matrix: identity part + rows are codewords (parity-check by re-encoding)
    G = np.zeros((71, 255), np.int8)
    oracle.lib().ref_bch_genmat(_p(G))
    assert (G[:, :71] == np.eye(71, dtype=np.int8)).all()
    msg = np.random.default_rng(1).integers(0, 2, 71).astype(np.uint8)
    par = np.zeros(184, np.uint8)
    oracle.lib().ref_bch_encode(_p(msg), _p(par))
    assert ((msg @ G.astype(np.int64)) % 2 == np.concatenate([msg, par])).all()


def test_polar_conventions(oracle):
    """shorten()==truncate, systematic property, frozen u all zero."""
    pl = oracle.make_payload(3)
    code = np.zeros(64800, np.uint8)
    oracle.lib().ref_payload_to_code(_p(pl), 6, _p(code))
    fr = np.zeros(2048, np.uint32)
    oracle.lib().ref_frozen_table(0, _p(fr))
    frozen = np.unpackbits(fr.view(np.uint8), bitorder="little").astype(bool)
    scr = np.zeros(5380, np.uint32)
    x = np.zeros(5380, np.uint32)
    oracle.lib().ref_xorshift(5380, _p(x))
    scrambled = pl ^ (x & 255).astype(np.uint8)
    data_bits = np.unpackbits(scrambled, bitorder="little")
    full = np.concatenate([code, np.zeros(736, np.uint8)])
    assert (full[~frozen][:43040] == data_bits).all()                  # systematic: message visible at free positions
    # u = x * F^{(x)n} (involution): frozen positions of u are zero
    u = full.copy()
    h = 1
    while h < 65536:
        v = u.reshape(-1, 2, h)
        v[:, 0, :] ^= v[:, 1, :]
        h *= 2
    assert not u[frozen].any()


def test_scalars():
    assert abs(2 * np.sin(np.pi / 8) - 0.76536686) < 1e-7
    assert abs(np.sqrt(1280 / 432) - 1.7213259) < 1e-6
    assert abs(np.sqrt(2560 / 127) - 4.4897083) < 1e-6


def test_osd_pruned_equals_literal(oracle):
    """The branch-and-bound OSD returns the literal order-4 sweep's codeword and `unique` flag."""
    rng = np.random.default_rng(5)
    G = np.zeros((71, 255), np.int8)
    oracle.lib().ref_bch_genmat(_p(G))
    for trial in range(12):
        msg = rng.integers(0, 2, 71)
        cw = (msg @ G.astype(np.int64)) % 2
        sigma = [0.3, 0.6, 0.9, 1.2][trial % 4]
        y = (1 - 2 * cw) + sigma * rng.standard_normal(255)
        soft = np.clip(np.rint(127 * y), -128, 127).astype(np.int8)
        o1, o2 = np.zeros(32, np.uint8), np.zeros(32, np.uint8)
        v1, v2 = C.c_int64(0), C.c_int64(0)
        u1 = oracle.lib().ref_osd(_p(soft), 1, _p(o1), C.byref(v1))
        u2 = oracle.lib().ref_osd(_p(soft), 0, _p(o2), C.byref(v2))
        assert v1.value == 1031347
        assert u1 == u2 and (not u1 or (o1 == o2).all()), (trial, sigma)
        assert v2.value <= v1.value

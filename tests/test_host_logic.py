"""Host-side product logic (modem_b200/csrc/host_tables.cc) against the oracle, and the lane-array emulation of the
CUDA list decoder's schedule / map algebra (tests/scl_emulator.cc) — all on CPU."""
import ctypes as C

import numpy as np


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_frozen_sets_match(oracle, hostlib):
    for fn, table in (("host_frozen", 0), ("host_frozen_alt", 1)):
        a, b = np.zeros(2048, np.uint32), np.zeros(2048, np.uint32)
        getattr(hostlib, fn)(_p(a))
        oracle.lib().ref_frozen_table(table, _p(b))
        assert (a == b).all()


def test_schedule_structure(hostlib):
    n = hostlib.host_schedule(None, 0)
    raw = np.zeros(n, np.uint32)
    hostlib.host_schedule(_p(raw), n)
    # OP_R1 (6) carries one extra word: the pc to continue at when the rate-1 attempt succeeds
    keep, skip_of, pc = [], {}, 0
    while pc < n:
        keep.append(pc)
        if raw[pc] & 7 == 6:
            skip_of[pc] = int(raw[pc + 1])
            pc += 1
        pc += 1
    keep = np.array(keep)
    ops = raw[keep]
    op, lvl, idx, depth = ops & 7, (ops >> 3) & 31, ((ops >> 8) & 0x3FFFFF) * 32, (ops >> 30) + 1
    assert op[-1] == 7 and (op[:-1] != 7).all()
    fr = np.zeros(2048, np.uint32)
    hostlib.host_frozen(_p(fr))
    # every non-all-frozen word is decoded exactly once, in order; rate-0 nodes cover exactly the all-frozen words
    words = idx[op == 2] // 32
    assert (np.diff(words) > 0).all() and set(words) == set(np.nonzero(fr != 0xFFFFFFFF)[0])
    covered = np.zeros(2048, bool)
    for l, i in zip(lvl[op == 3], idx[op == 3]):
        covered[i // 32:(i + (1 << l)) // 32] = True
    assert (covered == (fr == 0xFFFFFFFF)).all()
    # levels 16..14 are virtual: their 7 nodes only combine (C); the 8 level-13 nodes are produced by TOP ops
    top = op == 5
    assert top.sum() == 8 and (lvl[top] == 13).all() and (idx[top] == np.arange(8) * 8192).all()
    assert lvl[(op == 0) | (op == 1)].max() == 13 and (lvl[op == 4] >= 14).sum() == 7
    # every internal node below has one F step, one G and one C; F steps are explicit or fused into the op above them
    f_steps = depth[op == 0].sum() + (depth[op == 1] - 1).sum() + (depth[top] - 1).sum()
    assert f_steps == (op == 1).sum() == (op == 4).sum() - 7
    assert (lvl[op == 1] - depth[op == 1] + 1).min() >= 6 and depth.max() == 2 and (depth[top] == 2).all()
    assert (depth[(op != 0) & (op != 1) & ~top] == 1).all()
    # rate-1 attempts: on all-free nodes above the words only, never on the left child of an all-free node, and the skip
    # target is the op after the node's own C (the next op that touches a later index or a higher level)
    pos = {int(k): i for i, k in enumerate(keep)}
    assert len(skip_of) > 100
    for pc, tgt in skip_of.items():
        i = pos[pc]
        l, ix = int(lvl[i]), int(idx[i])
        assert 6 <= l <= 12 and (fr[ix // 32:(ix + (1 << l)) // 32] == 0).all()
        par = ix & ~((2 << l) - 1)
        assert ix != par or not (fr[par // 32:(par + (2 << l)) // 32] == 0).all()   # a left child: its parent is not all-free
        assert op[i + 1] == 0 and lvl[i + 1] == l and idx[i + 1] == ix               # the attempt is followed by the node's own F
        k = pos[tgt]
        assert op[k - 1] == 4 and lvl[k - 1] == l and idx[k - 1] == ix              # ... and skips to just after its own C


def test_fft_plans_of_all_sample_rates():
    """The product's Stockham passes and radix-2/3/4/5/7 butterflies (modem_b200/csrc/fft.cuh, compiled for the host by
    tests/fft_host.cu) against numpy for the symbol and half-symbol lengths of 8 / 16 / 44.1 / 48 kHz."""
    import os
    from modem_b200 import build
    build.build()
    lib = C.CDLL(build.FFTHOST)
    rng = np.random.default_rng(1)
    for n in (640, 1280, 2560, 3528, 3840, 7056, 7680):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        out = np.zeros(n, np.complex64)
        assert lib.fft_host(n, _p(x), _p(out)) == 0
        ref = np.fft.fft(x.astype(np.complex128))
        assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-6, n
    assert lib.fft_host(1000, _p(x), _p(out)) == -1


def test_mode_geometry_is_consistent(oracle):
    """Mode table of the Python mirror (decode.cc:302-374) against the oracle's row counts and the code lengths."""
    import modem_b200 as M
    bits = {6: 3, 7: 3, 8: 2, 9: 2, 10: 3, 11: 3, 12: 2, 13: 2}
    for mode, (rows, cols) in M.MODE_GEOMETRY.items():
        assert rows == oracle.MODE_ROWS[mode]
        assert rows * cols * bits[mode] == (64800 if mode < 10 else 64512)
        assert rows <= 126 and cols <= 512 and rows * cols <= 32400
    for rate, sym in ((8000, 1280), (16000, 2560), (44100, 7056), (48000, 7680)):
        assert oracle.frame_samples(6, rate) == 2 * rate + 55 * (sym + sym // 8)


def test_tables_match_oracle(oracle, hostlib):
    rows = np.zeros(71 * 8, np.uint32)
    hostlib.host_bch_rows(_p(rows))
    bits = np.unpackbits(rows.view(np.uint8), bitorder="little").reshape(71, 256)[:, :255]
    G = np.zeros((71, 255), np.int8)
    oracle.lib().ref_bch_genmat(_p(G))
    assert (bits == G).all()
    for poly, n in ((0b10001001, 127), (0b100101011, 255)):
        a, b = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        hostlib.host_mls(poly, n, _p(a))
        oracle.lib().ref_mls(poly, n, _p(b))
        assert (a == b).all()
    hostlib.host_crc16.restype = C.c_uint
    hostlib.host_crc16.argtypes = [C.c_uint64]
    assert hostlib.host_crc16(323559855980806 << 9) == 0xE724


def test_crc_pieces_join_to_the_bit_serial_crc(oracle, hostlib):
    """The list decoder's epilogue computes CRC-32 in eight pieces and joins them with the matrices of crc32_pieces():
    emulate that on random message bits and compare with the oracle's bit-serial CRC (decode.cc:534-537)."""
    rng = np.random.default_rng(5)
    for table in (0, 1):
        fr = np.zeros(2048, np.uint32)
        (hostlib.host_frozen_alt if table else hostlib.host_frozen)(_p(fr))
        free = np.unpackbits((~fr).view(np.uint8), bitorder="little").reshape(2048, 32).astype(bool)
        before = np.concatenate([[0], np.cumsum(free.sum(1))])
        pc = np.zeros(16 + 256, np.uint32)
        hostlib.host_crc_pieces(table, _p(pc))
        bounds = pc[:9].astype(int)
        assert bounds[0] == 0 and (np.diff(bounds) > 0).all() and before[bounds[8]] >= 43072 > before[bounds[8] - 1]
        for trial in range(3):
            code = rng.integers(0, 2, (2048, 32)).astype(np.uint8)
            mesg = code[free][:43072]
            want = oracle.lib().ref_crc32_bits(_p(np.ascontiguousarray(mesg)), 43072)
            total = 0
            for j in range(8):
                bits = code[bounds[j]:bounds[j + 1]][free[bounds[j]:bounds[j + 1]]]
                bits = bits[:max(0, 43072 - before[bounds[j]])]
                reg = oracle.lib().ref_crc32_bits(_p(np.ascontiguousarray(bits)), len(bits))
                for c in range(32):
                    if (reg >> c) & 1:
                        total ^= int(pc[16 + 32 * j + c])
            assert total == want, (table, trial)


def _noisy(oracle, seed, sigma):
    rng = np.random.default_rng(seed)
    pl = oracle.make_payload(500 + seed)
    code = np.zeros(64800, np.uint8)
    oracle.lib().ref_payload_to_code(_p(pl), 6, _p(code))
    y = (1.0 - 2.0 * code) + sigma * rng.standard_normal(64800)
    llr = np.concatenate([2 * y / max(sigma, 0.3) ** 2, np.full(736, 9000.0)]).astype(np.float32)
    return pl, llr


def test_emulator_matches_oracle_bit_exact(oracle, hostlib):
    """Same survivors (all 8 lanes), same fp32 metrics, at SNRs from clean to hopeless."""
    for seed, sigma in enumerate([0.0, 0.55, 0.7, 0.76, 0.85]):
        pl, llr = _noisy(oracle, seed, sigma)
        best, lanes, met, payload, flips = oracle.polar_decode(llr)
        el, em = np.zeros((8, 65536), np.uint8), np.zeros(8, np.float32)
        hostlib.emu_polar_decode(_p(llr), _p(el), _p(em), None)
        assert (el == lanes).all() and (em == met).all(), sigma
        if sigma <= 0.7:
            assert best == 0 and (payload == pl).all()


def test_emulator_second_code_table(oracle, hostlib):
    """frozen_64512_43072 (modes 10..13): schedule incl. TOP ops, map algebra and metrics against the oracle."""
    rng = np.random.default_rng(3)
    for seed, sigma in enumerate([0.0, 0.65, 0.8]):
        pl = oracle.make_payload(600 + seed)
        code = np.zeros(64512, np.uint8)
        oracle.lib().ref_payload_to_code(_p(pl), 10, _p(code))
        y = (1.0 - 2.0 * code) + sigma * rng.standard_normal(64512)
        llr = np.concatenate([2 * y / max(sigma, 0.3) ** 2, np.full(65536 - 64512, 9000.0)]).astype(np.float32)
        best, lanes, met, payload, flips = oracle.polar_decode(llr, table=1)
        el, em = np.zeros((8, 65536), np.uint8), np.zeros(8, np.float32)
        hostlib.emu_polar_decode_alt(_p(llr), _p(el), _p(em), None)
        assert (el == lanes).all() and (em == met).all(), sigma
        if sigma <= 0.65:
            assert best == 0 and (payload == pl).all()


def test_emulator_ties_and_zero_llrs(oracle, hostlib):
    """Degenerate inputs: all-equal magnitudes (every fork ties) and exact zeros."""
    pl, llr = _noisy(oracle, 9, 0.0)
    for variant in range(2):
        x = np.sign(llr).astype(np.float32) * 4.0
        if variant:
            x[::7] = 0.0
        x[64800:] = 9000.0
        best, lanes, met, payload, flips = oracle.polar_decode(x)
        el, em = np.zeros((8, 65536), np.uint8), np.zeros(8, np.float32)
        hostlib.emu_polar_decode(_p(x), _p(el), _p(em), None)
        assert (el == lanes).all() and (em == met).all()


def _scaled(oracle, seed, sigma, scale, mode=6):
    rng = np.random.default_rng(seed)
    pl = oracle.make_payload(500 + seed)
    nb = 64800 if mode < 10 else 64512
    code = np.zeros(nb, np.uint8)
    oracle.lib().ref_payload_to_code(_p(pl), mode, _p(code))
    y = (1.0 - 2.0 * code) + sigma * rng.standard_normal(nb)
    return np.concatenate([scale * y, np.full(65536 - nb, 9000.0)]).astype(np.float32)


def test_emulator_path_classes_and_rate1_attempts(oracle, hostlib):
    """The kernel's class bookkeeping (one stored copy per distinct path, slots found through the lane maps) and its rate-1
    attempts, over the regimes that matter: LLR scales at which the eight lanes stay copies of one path for the whole
    codeword (clean channel), split late, or split at once; with and without random refusals of the attempts and leaf
    shortcuts (what another codeword of the same warp can force).  The emulator poisons every slot that must not be read."""
    hostlib.emu_set_fail_seed.argtypes = [C.c_uint32]
    cases = [(0.0, 1000.0, 6), (0.05, 700.0, 6), (0.3, 60.0, 6), (0.5, 30.0, 6), (0.76, 300.0, 6), (0.7, 8.0, 6), (0.02, 900.0, 10), (0.65, 10.0, 10)]
    try:
        for seed, (sigma, scale, mode) in enumerate(cases):
            llr = _scaled(oracle, seed, sigma, scale, mode)
            best, lanes, met, payload, flips = oracle.polar_decode(llr, table=int(mode >= 10))
            for fail in (0, 4242 + seed):
                hostlib.emu_set_fail_seed(fail)
                el, em, st = np.zeros((8, 65536), np.uint8), np.zeros(8, np.float32), np.zeros(16, np.int64)
                (hostlib.emu_polar_decode_alt if mode >= 10 else hostlib.emu_polar_decode)(_p(llr), _p(el), _p(em), _p(st))
                assert (el == lanes).all() and (em == met).all(), (sigma, scale, mode, fail)
                classes = st[8:16]
                if scale >= 700 and not fail:
                    assert classes[0] == classes.sum() and st[3] == st[2] > 100   # one class throughout, every attempt succeeds
                if sigma >= 0.7:
                    assert classes[7] > 0.9 * classes.sum()                      # eight distinct paths almost from the start
    finally:
        hostlib.emu_set_fail_seed(0)

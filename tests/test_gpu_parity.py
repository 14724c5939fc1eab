"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs.
Bit-exact for integer/byte results (sync position, header bits, survivors, payload, status) and for the list decoder's
fp32 path metrics; stated tolerances for the floating-point stages ahead of the hard decision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# tolerances of the fp32 front end (FFT twiddle/ordering, atan2f/sincosf implementations, summation order differ
# between the device and the scalar oracle):
# SURVEY.md §8(c) asks for cons 1e-3, precision 1e-3, LLR 2e-3.  profiles/r2_parity_histogram.md holds the measured
# distributions (48 windows per channel condition): clean frames sit three orders of magnitude below those values.  On noisy
# frames two legitimate fp32 effects act, and the SURVEY values are enforced wherever they do not:
#  (1) the timing metric has a flat top; under noise its first maximum can sit one sample earlier or later in one of the two
#      fp32 pipelines.  The fine synchronisation absorbs that (sc_pos is identical), but the correlator phase is then read one
#      sample apart: the CFO estimate differs by ~4e-7 rad per sample (0.0005 Hz), which turns every point by ~5e-4 rad per
#      symbol and leaks ~3e-4 between carriers.  Frames whose (index_max, pos_err) equal the oracle's do not show it.
#  (2) a Theil-Sen median lands on the neighbouring order statistic in a row or two (inputs differ in the last bits): that row
#      turns by ~1e-3 rad as a whole, which moves its largest LLRs by up to 7e-3 of the mean.
TOL_CONS = 1e-3       # |cons_gpu - cons_ref| before derotation, points have |.| ~ 1 (frames with effect (1): 3e-3)
TOL_SLOPE_REL = 2e-2  # Theil-Sen slope relative to max|slope| of the frame: effect (2)
TOL_YINT = 2e-3       # rad: effects (1) and (2), measured max 1.3e-3
TOL_PRECISION = 1e-3  # relative (measured max 8e-4; frames with effect (1): 3e-3)
TOL_LLR = 2e-3        # |llr_gpu - llr_ref| / mean |llr_ref| on every row that was derotated like the oracle's ...
TOL_LLR_MAX = 1e-2    # ... and on the rows with effect (2), and on frames with effect (1)
TOL_SAME_ROW = 2e-4   # "derotated by the same line": max |cons_gpu - cons_ref| of the row after derotation (a jump is >= 3e-4)


def llr_close(got, ref, cons_gpu=None, cons_ref=None, mod_bits=3, same_sync=True):
    """LLR parity per constellation row: tight where the row was derotated like the oracle's, the rotation bound where its
    phase line jumped; at least half of the rows of a frame without effect (1) must be of the first kind (measured: 78 .. 100 %)."""
    scale = np.abs(ref[:64512]).mean()
    d = np.abs(got - ref) / scale
    assert d.max() < TOL_LLR_MAX, float(d.max())
    if not same_sync:
        return
    if cons_gpu is None:
        assert np.quantile(d, 0.85) < TOL_LLR, float(np.quantile(d, 0.85))
        return
    rows, cols = cons_ref.shape
    per_row = d[: rows * cols * mod_bits].reshape(rows, cols * mod_bits).max(axis=1)
    same = (np.abs(cons_gpu - cons_ref) / np.maximum(1.0, np.abs(cons_ref))).max(axis=1) < TOL_SAME_ROW
    assert same.mean() >= 0.5, float(same.mean())
    assert per_row[same].max() < TOL_LLR, float(per_row[same].max())
    assert d[rows * cols * mod_bits:].max() == 0.0     # lengthen(): the padding constant


def _noisy_llr(oracle, seed, sigma):
    import ctypes as C
    rng = np.random.default_rng(seed)
    pl = oracle.make_payload(700 + seed)
    code = np.zeros(64800, np.uint8)
    oracle.lib().ref_payload_to_code(pl.ctypes.data_as(C.c_void_p), 6, code.ctypes.data_as(C.c_void_p))
    y = (1.0 - 2.0 * code) + sigma * rng.standard_normal(64800)
    return pl, np.concatenate([2 * y / max(sigma, 0.3) ** 2, np.full(736, 9000.0)]).astype(np.float32)


def test_device_tables_match_oracle(rx, oracle):
    import ctypes as C
    for tb in (0, 1):
        fr = np.zeros(2048, np.uint32)
        oracle.lib().ref_frozen_table(tb, fr.ctypes.data_as(C.c_void_p))
        assert (rx.table(2 * tb) == fr).all()


def test_polar_list_decoder_bit_exact(rx, oracle):
    """All 8 survivors, their fp32 metrics, the CRC pick, the payload bytes and the flip count equal the oracle's."""
    sig = [0.0, 0.3, 0.5, 0.6, 0.65, 0.7, 0.72, 0.74, 0.76, 0.78, 0.8, 0.9, 1.2, 0.55, 0.68, 0.71, 0.73]
    pls, llrs = zip(*[_noisy_llr(oracle, i, s) for i, s in enumerate(sig)])
    payload, st, xb = rx.polar_decode(np.stack(llrs), want_xbits=True)
    n_ok = 0
    for i, s in enumerate(sig):
        best, lanes, met, opay, flips = oracle.polar_decode(llrs[i])
        glanes = np.unpackbits(xb[i].view(np.uint8), bitorder="little").reshape(8, 65536)
        assert (glanes == lanes).all(), s
        assert (st["metrics"][i] == met).all(), s
        assert st["best_lane"][i] == best and st["flips"][i] == flips
        assert (payload[i] == opay).all()
        assert st["status"][i] == (0 if best >= 0 else 6)
        n_ok += best >= 0
    assert 6 <= n_ok < len(sig)   # the sweep straddles the decoding threshold


def test_polar_degenerate_inputs(rx, oracle):
    """ties on every fork, exact zeros, all-erased, huge magnitudes"""
    pl, base = _noisy_llr(oracle, 50, 0.0)
    cases = []
    a = np.sign(base) * 4.0
    a[64800:] = 9000
    cases.append(a)
    b = a.copy()
    b[::7] = 0.0
    cases.append(b)
    c = np.zeros(65536, np.float32)
    c[64800:] = 9000
    cases.append(c)
    cases.append(base * 1e6)
    llr = np.stack(cases).astype(np.float32)
    payload, st, xb = rx.polar_decode(llr, want_xbits=True)
    for i in range(len(cases)):
        best, lanes, met, opay, flips = oracle.polar_decode(llr[i])
        glanes = np.unpackbits(xb[i].view(np.uint8), bitorder="little").reshape(8, 65536)
        assert (glanes == lanes).all() and (st["metrics"][i] == met).all() and st["best_lane"][i] == best
        assert (payload[i] == opay).all()


def _theil_sen_exact(y):
    """What DSP::TheilSenEstimator computes, restated with numpy fp32 (IEEE) arithmetic: the element of rank count/2 of
    all pairwise quotients (y_j - y_i) / (x_j - x_i), then of y_i - slope * x_i (decode.cc:488; oracle/ref_dsp.hh)."""
    y = np.asarray(y, np.float32)
    n = y.shape[0]
    i, j = np.triu_indices(n, 1)
    q = ((y[j] - y[i]).astype(np.float32) / (j - i).astype(np.float32)).astype(np.float32)
    slope = np.partition(q, q.size // 2)[q.size // 2]
    x = (np.arange(n) - n // 2).astype(np.float32)
    z = (y - (slope * x).astype(np.float32)).astype(np.float32)
    return slope, np.partition(z, n // 2)[n // 2]


def test_theil_sen_is_the_exact_order_statistic(rx):
    """Bit-exact against the brute-force order statistic on synthetic rows: Gaussian phase noise at several levels,
    lines with outliers, heavy ties (quantised phases, erased carriers), constant and exactly linear rows."""
    rng = np.random.default_rng(11)
    x = np.arange(432) - 216
    rows = []
    for sigma in (1e-4, 3e-3, 0.03, 0.1, 0.25):
        for slope in (0.0, 1.3e-4, -9e-4):
            rows.append(np.clip(0.05 + slope * x + sigma * rng.standard_normal(432), -np.pi / 8, np.pi / 8))
    r = 0.02 * rng.standard_normal(432); r[::9] = rng.uniform(-0.39, 0.39, 48); rows.append(r)          # outliers
    r = 0.02 * rng.standard_normal(432); r[rng.random(432) < 0.4] = 0.0; rows.append(r)                 # erased carriers
    rows.append(np.round(0.05 * rng.standard_normal(432) * 16) / 16)                                    # few distinct values
    rows.append(np.round(rng.uniform(-0.39, 0.39, 432) * 4) / 4)
    rows.append(np.zeros(432)); rows.append(np.full(432, 0.125)); rows.append(x / 1024.0)               # constant, linear
    rows.append(rng.uniform(-0.39, 0.39, 432))                                                          # no line at all
    r = np.zeros(432); r[200:] = 0.3; rows.append(r)                                                    # step
    y = np.stack(rows).astype(np.float32)
    slope, yint = rx.theil_sen(y)
    for k in range(y.shape[0]):
        es, ey = _theil_sen_exact(y[k])
        assert slope[k] == es and yint[k] == ey, (k, slope[k], es, yint[k], ey)
    # the carrier counts of the other operation modes (decode.cc:313-368) and odd sizes
    for cols in (512, 400, 384, 360, 256, 33, 8):
        xc = np.arange(cols) - cols // 2
        yc = np.stack([np.clip(0.02 + s1 * xc + sg * rng.standard_normal(cols), -np.pi / 4, np.pi / 4)
                       for sg in (1e-3, 0.05, 0.2) for s1 in (0.0, 2e-4)] + [np.zeros(cols), np.round(rng.standard_normal(cols) * 2) / 8]).astype(np.float32)
        slope, yint = rx.theil_sen(yc)
        for k in range(yc.shape[0]):
            es, ey = _theil_sen_exact(yc[k])
            assert slope[k] == es and yint[k] == ey, (cols, k, slope[k], es, yint[k], ey)


def test_theil_sen_exact_on_pipeline_rows(rx, oracle):
    """The same check on the phase errors the device itself produced for noisy frames (tap PHASE -> tap TS)."""
    import modem_b200 as M
    imp = oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-20, seed=5)
    pcm, ns, sent = oracle.encode_batch(3, seed0=7100, channels=2, imp=imp)
    payload, st = rx.decode(pcm, channels=2)
    assert (st["status"] == 0).all()
    for f in range(3):
        y = rx.taps(M.TAP_PHASE, f, 1)[0]
        ts = rx.taps(M.TAP_TS, f, 1)[0]   # mode 6 geometry
        for row in range(0, 50, 7):
            es, ey = _theil_sen_exact(y[row])
            assert ts[row, 0] == es and ts[row, 1] == ey, (f, row)


def _compare_frames(rx, oracle, pcm, channels, sent, strict_payload=True):
    import modem_b200 as M
    payload, st = rx.decode(pcm, channels=channels)
    for i in range(pcm.shape[0]):
        ost, opay, tp = oracle.decode(pcm[i], channels=channels)
        s = st[i]
        assert s["status"] == ost, (i, s["status"], ost)
        if tp.detections:
            # the symbol position and the integer CFO are exact; the split of the position between the coarse arg-max of
            # the (flat-topped) timing metric and the fine correction may move by one sample under noise
            assert (s["sc_pos"], s["shift"]) == (tp.sc_pos, tp.shift) and abs(int(s["pos_err"]) - tp.pos_err) <= 1
            assert abs(s["cfo_rad"] - tp.cfo_rad) < 1e-5
            dsoft = np.abs(rx.taps(M.TAP_SOFT, i, 1)[0][:255].astype(int) - oracle.taps_np(tp, "soft")[:255].astype(int))
            assert dsoft.max() <= 1 and (dsoft != 0).sum() <= 8   # rint() of a float that differs in the last ulps
            assert ((int(s["md_hi"]) << 32) | int(s["md_lo"])) == tp.md and s["mode"] == tp.mode
        if ost in (0, 6):
            md = int(s["mode"])
            same_sync = (int(s["index_max"]), int(s["pos_err"])) == (tp.index_max, tp.pos_err)   # effect (1) absent
            loose = 1 if same_sync else 3
            assert np.abs(rx.taps(M.TAP_CONS_RAW, i, 1, md)[0] - oracle.taps_np(tp, "cons_raw")).max() < TOL_CONS * loose
            oc = oracle.taps_np(tp, "cons")   # derotated: the phase-line difference acts on |cons| (> 1 under multipath) at |x| <= 256
            gc = rx.taps(M.TAP_CONS, i, 1, md)[0]
            assert (np.abs(gc - oc) / np.maximum(1.0, np.abs(oc))).max() < 4e-3   # a row with effect (2) at the band edge
            ts = rx.taps(M.TAP_TS, i, 1, md)[0]
            osl = oracle.taps_np(tp, "slope")
            assert np.abs(ts[:, 0] - osl).max() <= TOL_SLOPE_REL * np.abs(osl).max() + 1e-7
            assert np.abs(ts[:, 1] - oracle.taps_np(tp, "yint")).max() < TOL_YINT
            assert (np.abs(ts[:, 2] - oracle.taps_np(tp, "precision")) / oracle.taps_np(tp, "precision")).max() < TOL_PRECISION * loose
            ollr = oracle.taps_np(tp, "llr")
            mi = M.MODE_GEOMETRY[md]
            llr_close(rx.taps(M.TAP_LLR, i, 1)[0], ollr, gc, oc, mod_bits=(64800 if md < 10 else 64512) // (mi[0] * mi[1]), same_sync=same_sync)
        if ost == 0:
            assert (payload[i] == opay).all() and (payload[i] == sent[i]).all()
            assert s["best_lane"] == tp.best_lane
        elif strict_payload:
            assert (payload[i] == opay).all()   # failed window: de-scrambled zero buffer on both sides
    return st


def test_pipeline_clean_mono_config1_and_2(rx, oracle):
    """BASELINE configs[0]/[1]: 8000 Hz 16-bit real WAV, clean loop-back — payload bit-exact, taps within tolerance."""
    pcm, ns, sent = oracle.encode_batch(24, seed0=1000)
    st = _compare_frames(rx, oracle, pcm, 1, sent)
    assert (st["status"] == 0).all() and (st["flips"] == 0).all()


def test_pipeline_impaired_iq_config3(rx, oracle):
    """BASELINE configs[2]: multipath + CFO 234.567 Hz + SFO 147 ppm + AWGN -30 dB on the analytic signal."""
    imp = oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=77)
    pcm, ns, sent = oracle.encode_batch(16, seed0=2000, channels=2, imp=imp)
    st = _compare_frames(rx, oracle, pcm, 2, sent)
    assert (st["status"] == 0).all()


@pytest.mark.parametrize("mode", [7, 8, 9, 10, 11, 12, 13])
def test_pipeline_other_modes(oracle, mode):
    """Modes 7..13 (decode.cc:313-368): QPSK and 8PSK on 256..512 carriers, 42..126 rows, both frozen sets — clean mono
    frames bit-exact with taps in tolerance, and impaired analytic frames decoding to the oracle's bytes."""
    import modem_b200 as M
    stride = oracle.frame_samples(mode) + 64   # slack: a negative sampling-frequency offset stretches the stream
    pcm, ns, sent = oracle.encode_batch(5, seed0=100 * mode, mode=mode, stride=stride)
    rxm = M.Receiver(max_frames=8, max_samples=stride, keep_taps=True)
    try:
        st = _compare_frames(rxm, oracle, pcm, 1, sent)
        assert (st["status"] == 0).all() and (st["mode"] == mode).all() and (st["flips"] == 0).all()
        imp = oracle.impair(multipath=True, cfo_hz=-77.7, sfo_ppm=-60, awgn_db=-26, seed=mode)
        pcm, ns, sent = oracle.encode_batch(5, seed0=100 * mode + 50, channels=2, mode=mode, imp=imp, stride=pcm.shape[1])
        _compare_frames(rxm, oracle, pcm, 2, sent)
    finally:
        rxm.close()


def _compare_frames_rate(rxh, oracle, pcm, channels, sent, rate):
    payload, st = rxh.decode(pcm, channels=channels)
    for i in range(pcm.shape[0]):
        ost, opay, tp = oracle.decode(pcm[i], channels=channels, rate=rate)
        s = st[i]
        assert s["status"] == ost, (i, s["status"], ost)
        assert (s["sc_pos"], s["shift"]) == (tp.sc_pos, tp.shift) and abs(int(s["pos_err"]) - tp.pos_err) <= 1
        assert abs(s["cfo_rad"] - tp.cfo_rad) < 1e-5 and s["mode"] == tp.mode
        if ost == 0:
            assert (payload[i] == opay).all() and (payload[i] == sent[i]).all() and s["best_lane"] == tp.best_lane
    return st


@pytest.mark.parametrize("rate,mode", [(16000, 6), (16000, 9), (16000, 12), (44100, 6), (44100, 13), (48000, 6), (48000, 10)])
def test_pipeline_other_sample_rates(oracle, rate, mode):
    """16000 / 44100 / 48000 Hz (decode.cc:171-173,590-606): symbol lengths 2560 / 7056 / 7680 (FFT radices 2, 3, 4, 5, 7),
    Hilbert<41/113/125>, correlator lengths 1280 / 3528 / 3840 — clean mono and impaired analytic frames against the
    oracle at the same rate."""
    import modem_b200 as M
    stride = oracle.frame_samples(mode, rate) + 512
    pcm, ns, sent = oracle.encode_batch(3, seed0=rate // 10 + mode, rate=rate, mode=mode, stride=stride)
    rxh = M.Receiver(max_frames=3, max_samples=stride, rate=rate)
    try:
        st = _compare_frames_rate(rxh, oracle, pcm, 1, sent, rate)
        assert (st["status"] == 0).all() and (st["flips"] == 0).all()
        imp = oracle.impair(multipath=True, cfo_hz=91.3, sfo_ppm=80, awgn_db=-28, seed=16 + mode)
        pcm, ns, sent = oracle.encode_batch(3, seed0=rate // 10 + 100 + mode, rate=rate, channels=2, mode=mode, imp=imp, stride=stride)
        st = _compare_frames_rate(rxh, oracle, pcm, 2, sent, rate)
        assert (st["status"] == 0).all()
    finally:
        rxh.close()


def test_mixed_modes_in_one_batch(oracle):
    """Windows of different modes (both code tables) in one call: each decodes as it does alone."""
    import modem_b200 as M
    frames = [oracle.encode_batch(2, seed0=7000 + m, mode=m) for m in (13, 6, 10, 8, 6, 11)]
    stride = max(f[0].shape[1] for f in frames)
    pcm = np.zeros((12, stride), np.int16)
    sent = np.concatenate([f[2] for f in frames])
    ns = np.zeros(12, np.int32)
    for k, f in enumerate(frames):
        pcm[2 * k:2 * k + 2, :f[0].shape[1]] = f[0]
        ns[2 * k:2 * k + 2] = f[0].shape[1]
    rxm = M.Receiver(max_frames=12, max_samples=stride)
    try:
        payload, st = rxm.decode(pcm, n_samples=ns)
        assert (st["status"] == 0).all() and (payload == sent).all()
        assert list(st["mode"]) == [13, 13, 6, 6, 10, 10, 8, 8, 6, 6, 11, 11]
    finally:
        rxm.close()


def test_polar_second_code_table(rx, oracle):
    """frozen_64512_43072 (modes 10..13): all 8 survivors and metrics equal the oracle's."""
    import ctypes as C
    rng = np.random.default_rng(5)
    for k, sigma in enumerate((0.0, 0.6, 0.72, 0.8)):
        pl = oracle.make_payload(800 + k)
        code = np.zeros(64512, np.uint8)
        oracle.lib().ref_payload_to_code(pl.ctypes.data_as(C.c_void_p), 10, code.ctypes.data_as(C.c_void_p))
        y = (1.0 - 2.0 * code) + sigma * rng.standard_normal(64512)
        llr = np.concatenate([2 * y / max(sigma, 0.3) ** 2, np.full(65536 - 64512, 9000.0)]).astype(np.float32)
        payload, st, xb = rx.polar_decode(llr[None], want_xbits=True, table=1)
        best, lanes, met, opay, flips = oracle.polar_decode(llr, table=1)
        glanes = np.unpackbits(xb[0].view(np.uint8), bitorder="little").reshape(8, 65536)
        assert (glanes == lanes).all() and (st["metrics"][0] == met).all()
        assert st["best_lane"][0] == best and (payload[0] == opay).all() and st["flips"][0] == flips
    rx.polar_decode(llr[None], table=0)   # leave the shared fixture on table 0


def test_pipeline_awgn_near_threshold(rx, oracle):
    """Frames around the waterfall: every window the oracle decodes must decode to the same bytes."""
    n_ok = 0
    for db in (-16.0, -14.5, -13.5):
        imp = oracle.impair(awgn_db=db, seed=int(-db * 10))
        pcm, ns, sent = oracle.encode_batch(8, seed0=3000, channels=2, imp=imp)
        payload, st = rx.decode(pcm, channels=2)
        for i in range(8):
            ost, opay, tp = oracle.decode(pcm[i], channels=2, want_taps=False)
            if ost == 0:
                assert st["status"][i] == 0 and (payload[i] == opay).all()
                n_ok += 1
    assert n_ok > 0


def test_skip_semantics_and_ragged_windows(oracle):
    import modem_b200 as M
    pls = np.stack([oracle.make_payload(40 + i) for i in range(3)])
    multi = oracle.encode(pls)                       # three frames in one stream (encode.cc:289)
    single = oracle.encode(pls[0])
    stride = multi.shape[0]
    pcm = np.zeros((4, stride), np.int16)
    pcm[0] = multi
    pcm[1, :single.shape[0]] = single
    pcm[2, :30000] = single[:30000]                  # cut inside the payload
    ns = np.array([stride, single.shape[0], 30000, 5000], np.int32)   # window 3: silence
    rx2 = M.Receiver(max_frames=4, max_samples=stride, keep_taps=False)
    for skip in range(4):
        payload, st = rx2.decode(pcm, n_samples=ns, skip=skip)
        for i in range(4):
            ost, opay, tp = oracle.decode(pcm[i, :ns[i]], skip=skip, want_taps=True)
            assert st["status"][i] == ost, (skip, i)
            assert (payload[i] == opay).all()
            assert st["detections"][i] == tp.detections
        if skip < 3:
            assert (payload[0] == pls[skip]).all()
    rx2.close()


def test_garbage_and_empty_inputs(rx, oracle):
    rng = np.random.default_rng(3)
    pcm = np.zeros((6, 95200), np.int16)
    pcm[0] = (rng.standard_normal(95200) * 8000).clip(-32767, 32767)
    pcm[1] = 32767
    pcm[2, ::2] = 20000
    pcm[3] = (np.sin(2 * np.pi * 2000 / 8000 * np.arange(95200)) * 20000)
    good = oracle.encode(oracle.make_payload(60))
    pcm[4] = good
    pcm[4, 11000:13000] = (rng.standard_normal(2000) * 10000)       # destroy the header symbol
    pcm[5] = good[::-1]
    payload, st = rx.decode(pcm)
    for i in range(6):
        ost, opay, tp = oracle.decode(pcm[i], want_taps=False)
        assert st["status"][i] == ost and (payload[i] == opay).all(), i
    payload, st = rx.decode(np.zeros((0, 95200), np.int16))
    assert payload.shape == (0, 5380)


def test_full_size_round_trip_and_determinism(rx, oracle):
    """Size-independent properties at scale: encode -> decode returns the sent bytes, bit flips 0, and repeated runs
    give identical bytes (2048 distinct windows; the 10k-window case is bench.py's gate)."""
    pcm, ns, sent = oracle.encode_batch(2048, seed0=50000)
    p1, s1 = rx.decode(pcm)
    p2, s2 = rx.decode(pcm)
    assert (s1["status"] == 0).all() and (p1 == sent).all() and (s1["flips"] == 0).all()
    assert (p1 == p2).all() and (s1["metrics"] == s2["metrics"]).all()
    for i in range(0, 2048, 256):   # the oracle agrees on a strided sample
        ost, opay, tp = oracle.decode(pcm[i], want_taps=False)
        assert ost == 0 and (opay == p1[i]).all()


def test_batches_larger_than_the_handle_are_chunked(oracle):
    """n_frames > max_frames: the call walks the batch in chunks of max_frames; results equal the one-shot decode."""
    import modem_b200 as M
    pcm, ns, sent = oracle.encode_batch(21, seed0=8100)
    small = M.Receiver(max_frames=8)
    try:
        payload, st = small.decode(pcm)
        assert (st["status"] == 0).all() and (payload == sent).all()
        ns2 = np.full(21, 95200, np.int32)
        ns2[5] = 40000                      # one ragged window in the second chunk's neighbourhood
        payload, st = small.decode(pcm, n_samples=ns2)
        assert st["status"][5] != 0 and (np.delete(st["status"], 5) == 0).all() and (np.delete(payload, 5, 0) == np.delete(sent, 5, 0)).all()
    finally:
        small.close()


def test_device_memory_interface(rx, oracle):
    import torch
    import modem_b200 as M
    pcm, ns, sent = oracle.encode_batch(8, seed0=900, channels=2)
    d = torch.from_numpy(pcm).cuda()
    out = torch.empty((8, 5380), dtype=torch.uint8, device="cuda")
    st = torch.empty((8, 112), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rx.decode_raw(d.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, 8, pcm.shape[1] // 2, None, 0, out.data_ptr(), st.data_ptr(), stream)
    torch.cuda.synchronize()
    assert (out.cpu().numpy() == sent).all()
    f = (d.to(torch.float32) / 32767.0).contiguous()   # float2 windows from device memory
    rx.decode_raw(f.data_ptr(), M.MEM_DEVICE, M.FMT_F32_IQ, 8, pcm.shape[1] // 2, None, 0, out.data_ptr(), st.data_ptr(), stream)
    torch.cuda.synchronize()
    assert (out.cpu().numpy() == sent).all()
    ms, n = rx.stage_times()
    assert n == 8 and ms["polar_scl"] > 0


def test_decode_cli_matches_reference_contract(oracle, tmp_path):
    """`modem_b200/decode OUTPUT INPUT [SKIP]` (C++ host driver over the C-ABI) against the oracle's decode_ref on the
    README quick-start recipe: same bytes, exit 0, the reference's stderr lines; failure still writes 5380 bytes."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dec, ref, enc = os.path.join(root, "modem_b200", "decode"), os.path.join(root, "oracle", "build", "decode_ref"), os.path.join(root, "oracle", "build", "encode_ref")
    data = [os.urandom(5380) for _ in range(2)]
    for i, d in enumerate(data):
        (tmp_path / ("in%d.dat" % i)).write_bytes(d)
    wav = str(tmp_path / "enc.wav")
    for ch in ("1", "2"):
        subprocess.run([enc, wav, "8000", "16", ch, "2000", "6", "CALLSIGN", str(tmp_path / "in0.dat"), str(tmp_path / "in1.dat")], check=True)
        for skip in (0, 1):
            r = subprocess.run([dec, str(tmp_path / "gpu.dat"), wav, str(skip)], capture_output=True)
            q = subprocess.run([ref, str(tmp_path / "cpu.dat"), wav, str(skip)], capture_output=True)
            assert r.returncode == 0 and q.returncode == 0
            assert (tmp_path / "gpu.dat").read_bytes() == (tmp_path / "cpu.dat").read_bytes() == data[skip]
            for line in (b"symbol pos:", b"coarse cfo:", b"oper mode: 6", b"call sign:  CALLSIGN", b"bit flips: 0"):
                assert line in r.stderr, line
    # another operation mode through the same command line (mode 13: QPSK, 256 carriers, 126 rows, second code table)
    subprocess.run([enc, wav, "8000", "16", "1", "1500", "13", "N0CALL", str(tmp_path / "in1.dat")], check=True)
    r = subprocess.run([dec, str(tmp_path / "gpu.dat"), wav], capture_output=True)
    q = subprocess.run([ref, str(tmp_path / "cpu.dat"), wav], capture_output=True)
    assert r.returncode == 0 and (tmp_path / "gpu.dat").read_bytes() == (tmp_path / "cpu.dat").read_bytes() == data[1]
    assert b"oper mode: 13" in r.stderr and b"call sign:    N0CALL" in r.stderr and b"bit flips: 0" in r.stderr
    # another sample rate through the same command line (48 kHz, 2-channel: symbol length 7680)
    subprocess.run([enc, wav, "48000", "16", "2", "3000", "7", "CQ", str(tmp_path / "in0.dat")], check=True)
    r = subprocess.run([dec, str(tmp_path / "gpu.dat"), wav], capture_output=True)
    q = subprocess.run([ref, str(tmp_path / "cpu.dat"), wav], capture_output=True)
    assert r.returncode == 0 and (tmp_path / "gpu.dat").read_bytes() == (tmp_path / "cpu.dat").read_bytes() == data[0]
    assert b"oper mode: 7" in r.stderr and b"coarse cfo: 3000" in r.stderr and b"bit flips: 0" in r.stderr
    r = subprocess.run([dec], capture_output=True)
    assert r.returncode == 1 and b"usage:" in r.stderr
    raw = bytearray(open(wav, "rb").read())
    raw[44:] = bytes(len(raw) - 44)
    open(wav, "wb").write(raw)
    r = subprocess.run([dec, str(tmp_path / "gpu.dat"), wav], capture_output=True)
    q = subprocess.run([ref, str(tmp_path / "cpu.dat"), wav], capture_output=True)
    assert r.returncode == 0 and (tmp_path / "gpu.dat").read_bytes() == (tmp_path / "cpu.dat").read_bytes()
    assert len((tmp_path / "gpu.dat").read_bytes()) == 5380


def test_against_committed_golden_vectors(rx, oracle):
    """tests/golden/oracle_vectors.npz (made by tests/golden/make_golden.py from the CPU oracle; the CPU suite checks that the
    oracle still reproduces it): the CUDA path against the committed numbers themselves, clean and through the README chain."""
    import os
    import modem_b200 as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
    pl = g["payload"]
    payload, st = rx.decode(oracle.encode(pl).reshape(1, -1), channels=1)
    s = st[0]
    assert s["status"] == 0 and (payload[0] == pl).all() and s["flips"] == 0
    assert s["sc_pos"] == int(g["sc_pos"]) and abs(s["cfo_rad"] - float(g["cfo_rad"])) < 1e-5
    dsoft = np.abs(rx.taps(M.TAP_SOFT, 0, 1)[0][:255].astype(int) - g["soft"].astype(int))
    assert dsoft.max() <= 1 and (dsoft != 0).sum() <= 8
    prec = rx.taps(M.TAP_TS, 0, 1)[0][:, 2]
    assert (np.abs(prec - g["precision"]) / g["precision"]).max() < TOL_PRECISION
    head = g["llr_head"]
    assert np.abs(rx.taps(M.TAP_LLR, 0, 1)[0][:256] - head).max() / np.abs(head).mean() < TOL_LLR
    imp = oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=9)
    payload, st = rx.decode(oracle.encode(pl, channels=2, imp=imp).reshape(1, -1), channels=2)
    s = st[0]
    assert s["status"] == 0 and (payload[0] == pl).all()
    assert s["sc_pos"] == int(g["imp_sc_pos"]) and abs(s["cfo_rad"] - float(g["imp_cfo_rad"])) < 1e-5
    slope = rx.taps(M.TAP_TS, 0, 1)[0][:, 0]
    assert np.abs(slope - g["imp_slope"]).max() <= TOL_SLOPE_REL * np.abs(g["imp_slope"]).max() + 1e-7

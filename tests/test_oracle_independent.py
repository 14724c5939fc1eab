"""Independent pins of the oracle's recalled third-party primitives (aicodix/dsp, aicodix/code are not on the box): each is
checked against a SEPARATE implementation — numpy / scipy in float64, or a textbook restatement written here — that shares
no code with oracle/ref_*.hh.  What this pins: the arithmetic of the oracle as it is stated.  What it cannot pin: whether the
statement itself (delay conventions, the 0 / 1000 initial path metrics, tie orders) is what aicodix's sources do — DESIGN.md §1."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [640, 1280, 2560, 3528, 3840, 7056, 7680])
def test_fft_against_numpy(oracle, n):
    """DSP::FastFourierTransform<N, cmplx, -1 / +1> (decode.cc:43-44,191): unnormalised mixed-radix transforms of every
    symbol and half-symbol length of the four sample rates."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for sign in (-1, 1):
        out = np.zeros(n, np.complex64)
        oracle.lib().ref_fft(n, sign, _p(x), _p(out))
        ref = np.fft.fft(x.astype(np.complex128)) if sign < 0 else np.fft.ifft(x.astype(np.complex128)) * n
        assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-6, (n, sign)


def _hilbert_taps(taps, a=2.0):
    """closed form: Kaiser(a) windowed ideal Hilbert transformer, odd offsets k: 2 / (pi k) * w[k + (taps-1)/2]"""
    n = np.arange(taps)
    w = np.i0(np.pi * a * np.sqrt(1.0 - (2.0 * n / (taps - 1) - 1.0) ** 2)) / np.i0(np.pi * a)
    mid = (taps - 1) // 2
    h = np.zeros(taps)
    for k in range(1, mid + 1, 2):
        h[mid + k] = 2.0 / (np.pi * k) * w[mid + k]
        h[mid - k] = -h[mid + k]
    return w[mid], h


@pytest.mark.parametrize("rate", [8000, 16000, 44100, 48000])
def test_front_end_against_scipy(oracle, rate):
    """next_sample() for one channel (decode.cc:294-301): BlockDC as the IIR b (x - x[-1]) + a y[-1] with a = (s-1)/s,
    b = (1+a)/2, s = 2 (symbol + guard) (scipy.signal.lfilter in float64), then the Kaiser(2)-windowed Hilbert FIR of
    ((21 rate / 8000) & ~3) | 1 taps from its closed form, real branch = centre tap of the delay line — against the oracle's
    per-step stream taps; plus a property no restatement can fake: the result is analytic (image rejection mid-band)."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(rate)
    n = 6000
    x16 = (rng.standard_normal(n) * 6000).clip(-32767, 32767).astype(np.int16)
    x16[:50] += 3000   # a DC step for the blocker
    iq, timing = oracle.front_taps(x16, channels=1, rate=rate)
    x = np.concatenate([x16.astype(np.float64) / 32767.0, [0.0]])   # the stream takes one step past the end (zeros)
    sym = 1280 * rate // 8000
    s = 2 * (sym + sym // 8)
    a = np.float32(s - 1) / np.float32(s)
    b = (np.float32(1) + a) / np.float32(2)
    y = lfilter([float(b), -float(b)], [1.0, -float(a)], x)
    taps = (((21 * rate) // 8000) & ~3) | 1
    reco, h = _hilbert_taps(taps)
    mid = (taps - 1) // 2
    # the filter output at step t is formed BEFORE y[t] enters the delay line: it sees y[t-taps .. t-1], centre y[t-1-mid]
    yp = np.concatenate([np.zeros(taps), y])
    re = reco * yp[taps - 1 - mid: taps - 1 - mid + n + 1]
    # correlation of the antisymmetric taps with the line (oldest first): sum_k h[mid+k] (line[mid-k] - line[mid+k]) for odd k > 0
    line = np.lib.stride_tricks.sliding_window_view(yp, taps)[: n + 1]      # line[t] = y[t-taps .. t-1]
    im = np.zeros(n + 1)
    for k in range(1, mid + 1, 2):
        im += h[mid + k] * (line[:, mid - k] - line[:, mid + k])
    ref = re + 1j * im
    assert np.abs(iq - ref).max() < 5e-6, np.abs(iq - ref).max()
    # analytic signal: a tone at +f0 in the middle of the band keeps its energy, its image at -f0 is suppressed
    t = np.arange(n)
    tone = (np.cos(2 * np.pi * 0.25 * t) * 12000).astype(np.int16)
    iqt, _ = oracle.front_taps(tone, channels=1, rate=rate)
    spec = np.abs(np.fft.fft(iqt[200:n].astype(np.complex128) * np.hanning(n - 200)))
    k0 = int(round(0.25 * (n - 200)))
    assert spec[k0 - 2:k0 + 3].max() > 100 * spec[-k0 - 2:-k0 + 3].max()      # > 40 dB image rejection at fs / 4


def test_timing_metric_against_brute_force(oracle):
    """SchmidlCox::operator() metric part (decode.cc:86-90) on a real frame: P over 640 lag products at the buffer taps
    search_pos + 640 / + 1280, R = half the energy of 1280 samples floored at 0.064, box-161 sum of |P|^2 / R^2 — cumulative
    sums in float64 against the oracle's per-step taps (which use the recalled SMA4 sliding sums in fp32)."""
    pcm = oracle.encode(oracle.make_payload(5), channels=2, imp=oracle.impair(awgn_db=-25, seed=3))
    iq, timing = oracle.front_taps(pcm, channels=2)
    a = iq.astype(np.complex128)
    n = a.size
    # newest sample is buffer tap 8639; taps 2880 + 640 and 2880 + 1280 are 5119 and 4479 steps old
    old = np.concatenate([np.zeros(5119, complex), a])[:n]
    cur = np.concatenate([np.zeros(4479, complex), a])[:n]
    c, e = old * np.conj(cur), np.abs(cur) ** 2
    def box(v, w):
        cs = np.concatenate([[0], np.cumsum(v)])
        i = np.arange(1, v.size + 1)
        return cs[i] - cs[np.maximum(i - w, 0)]
    P, R = box(c, 640), np.maximum(0.5 * box(e, 1280), 0.0001 * 640)
    ref = box(np.abs(P) ** 2 / R ** 2, 161)
    assert np.abs(timing - ref).max() < 1e-3 and ref.max() > 100.0


def test_theil_sen_against_numpy(oracle):
    """DSP::TheilSenEstimator (decode.cc:488): upper medians (element count / 2 after nth_element) of all pairwise fp32
    slopes and of the fp32 intercepts."""
    rng = np.random.default_rng(2)
    for n in (432, 256, 37):
        x = (np.arange(n) - n // 2).astype(np.float32)
        y = (0.003 * x + 0.2 + 0.05 * rng.standard_normal(n)).astype(np.float32)
        y[rng.integers(0, n, 5)] += 1.5
        slope, yint = np.zeros(1, np.float32), np.zeros(1, np.float32)
        oracle.lib().ref_theil_sen(_p(x), _p(y), n, _p(slope), _p(yint))
        i, j = np.triu_indices(n, 1)
        q = ((y[j] - y[i]).astype(np.float32) / (x[j] - x[i]).astype(np.float32)).astype(np.float32)
        s = np.partition(q, q.size // 2)[q.size // 2]
        z = (y - (s * x).astype(np.float32)).astype(np.float32)
        assert slope[0] == s and yint[0] == np.partition(z, n // 2)[n // 2]


def _scl_textbook(llr, frozen, L=8):
    """Textbook LLR-domain successive-cancellation list decoder (min-sum f, g = b +- a, penalty |llr| for a decision against
    the sign), natural bit order, recursive over (sub-tree, list); state per path: its partial sums.  fp32 like the reference
    (SIMD<float, 8>).  Conventions taken from the oracle's header: metrics start at 0, 1000, 1000, ...; the 2L forks are ranked
    by a stable sort on the metric in fork order 2 k + bit."""
    f32 = np.float32
    n = llr.size
    paths = [dict(metric=f32(0.0 if k == 0 else 1000.0), alpha=[llr.astype(f32)], u=[]) for k in range(L)]

    def fnode(a, b):
        return (np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))).astype(f32)

    def rec(paths, size, index):
        # every path holds alpha[-1] = the LLRs of this node; returns the paths extended by the node's partial sums beta
        if size == 1:
            if frozen[index]:
                for p in paths:
                    a = p["alpha"][-1][0]
                    if a < 0:
                        p["metric"] = f32(p["metric"] - a)
                    p["beta"] = np.zeros(1, np.uint8)
                return paths
            forks = []
            for k, p in enumerate(paths):
                a = p["alpha"][-1][0]
                m0 = f32(p["metric"] - a) if a < 0 else p["metric"]
                m1 = p["metric"] if a < 0 else f32(p["metric"] + a)
                forks += [(m0, 2 * k, k, 0), (m1, 2 * k + 1, k, 1)]
            forks.sort(key=lambda t: (t[0], t[1]))
            out = []
            for m, _, k, bit in forks[:L]:
                q = dict(metric=m, alpha=list(paths[k]["alpha"]), u=list(paths[k]["u"]), beta=np.array([bit], np.uint8))
                q["saved"] = dict(paths[k].get("saved", {}))
                out.append(q)
            return out
        h = size // 2
        for p in paths:
            a = p["alpha"][-1]
            p.setdefault("saved", {})
            p["alpha"] = p["alpha"] + [fnode(a[:h], a[h:])]
        paths = rec(paths, h, index)
        for p in paths:
            p["alpha"] = p["alpha"][:-1]
            a = p["alpha"][-1]
            bl = p["beta"]
            p["saved"] = dict(p["saved"]); p["saved"][(size, index)] = bl
            p["alpha"] = p["alpha"] + [np.where(bl == 1, a[h:] - a[:h], a[h:] + a[:h]).astype(f32)]
        paths = rec(paths, h, index + h)
        for p in paths:
            p["alpha"] = p["alpha"][:-1]
            bl = p["saved"][(size, index)]
            p["beta"] = np.concatenate([bl ^ p["beta"], p["beta"]])
        return paths

    paths = rec(paths, n, 0)
    order = sorted(range(L), key=lambda k: (paths[k]["metric"], k))
    return np.stack([paths[k]["beta"] for k in order]), np.array([paths[k]["metric"] for k in order], f32)


@pytest.mark.parametrize("order,seed", [(5, 1), (6, 2), (7, 3), (8, 4)])
def test_list_decoder_against_textbook_scl(oracle, order, seed):
    """CODE::PolarListDecoder (decode.cc:201,530) as the oracle states it, against a recursive textbook SCL written here,
    on small random codes with noisy LLRs: all 8 survivors (re-encoded codewords) and their fp32 metrics — bit for bit with
    the penalties of frozen leaves added leaf by leaf (r0_max = 1), and with the same survivors and metrics to rounding when
    the oracle sums a whole all-frozen sub-tree at its root (the order the GPU kernel uses too; equal in exact arithmetic)."""
    rng = np.random.default_rng(seed)
    n = 1 << order
    for trial in range(6):
        rel = rng.permutation(n)
        frozen = np.zeros(n, np.uint8)
        frozen[rel[: n // 2 + rng.integers(-n // 8, n // 8)]] = 1
        frozen[0] = 1
        words = np.packbits(frozen, bitorder="little").view(np.uint32).copy()
        llr = (rng.standard_normal(n) * 3 + 1.0).astype(np.float32)
        if trial == 5:
            llr[::5] = 0.0
        lanes, met = np.zeros((8, n), np.uint8), np.zeros(8, np.float32)
        oracle.lib().ref_polar_decode_any(order, _p(words), _p(llr), 1, _p(lanes), _p(met))
        tl, tm = _scl_textbook(llr, frozen)
        assert (lanes == tl).all() and (met == tm).all(), (order, trial)
        oracle.lib().ref_polar_decode_any(order, _p(words), _p(llr), 1 << 16, _p(lanes), _p(met))
        if len(set(tm.tolist())) == 8 and np.diff(tm).min() > 1e-3:      # (no near-ties that rounding could reorder)
            assert (lanes == tl).all() and np.allclose(met, tm, rtol=1e-5, atol=1e-5), (order, trial)

"""GPU parity tests (-m gpu) of the streaming front end and of the boundary additions of round 2, through the C-ABI:
stage taps K0 (ReadWAV scaling, BlockDC, Hilbert: TAP_IQ) and K1a (Schmidl-Cox timing metric: TAP_TIMING) against the
oracle's per-step taps, float sample formats (8 / 24 / 32-bit WAVs at the reference's precision, decode.cc:576), and SKIP
walks past sixteen detections (decode.cc:390-448 is unbounded)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_IQ = 4e-6        # |iq_gpu - iq_ref|, samples are <= 1: the DC blocker's recurrence runs as a scan (measured max 2.1e-6; I/Q input: 0)
TOL_TIMING = 1e-3    # timing metric (0 .. 161): prefix-sum differences over 3.5 k-sample tiles vs exact sliding sums (measured max 1.7e-4)


def _front(rx, oracle, pcm, channels, M):
    n = pcm.size // channels
    payload, st = rx.decode(pcm.reshape(1, -1), channels=channels)
    iq = rx.taps(M.TAP_IQ, 0, 1)[0].view(np.complex64)[: n + 1]
    tm = rx.taps(M.TAP_TIMING, 0, 1)[0][: n + 1]
    oiq, otm = oracle.front_taps(pcm, channels=channels)
    return payload, st, iq, tm, oiq, otm


def test_front_end_stage_taps(oracle):
    """a2/a3 (BlockDC + Hilbert) and a5 (P, R, box-161 metric) sample by sample over whole windows: a one-sample delay or a
    sign convention that the fine synchronisation would absorb shows up here."""
    import modem_b200 as M
    rx = M.Receiver(max_frames=1, keep_taps=True)
    try:
        cases = [("clean mono", oracle.encode(oracle.make_payload(71)), 1),
                 ("README chain IQ", oracle.encode(oracle.make_payload(72), channels=2,
                                                    imp=oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=9)), 2),
                 ("AWGN -16 mono", oracle.encode(oracle.make_payload(73), channels=1, imp=oracle.impair(awgn_db=-16, seed=10)), 1)]
        for name, pcm, ch in cases:
            pcm = np.ascontiguousarray(pcm).reshape(-1)
            n = pcm.size // ch
            if n > 95200:   # (the SFO stretches the stream by a few samples: cut to the handle's window)
                pcm = pcm[: 95200 * ch]
            payload, st, iq, tm, oiq, otm = _front(rx, oracle, pcm, ch, M)
            e_iq, e_tm = np.abs(iq - oiq).max(), np.abs(tm - otm).max()
            assert e_iq < TOL_IQ, (name, e_iq)
            assert e_tm < TOL_TIMING, (name, e_tm)
            # the trigger's decisions follow from the metric: same firing step, same arg-max bookkeeping
            ost, opay, tp = oracle.decode(pcm, channels=ch)
            assert st["status"][0] == ost and st["t_fire"][0] == tp.t_fire and st["sc_pos"][0] == tp.sc_pos, name
            if ch == 1 and "clean" in name:
                assert st["index_max"][0] == tp.index_max and abs(st["timing_max"][0] - tp.timing_max) < TOL_TIMING
    finally:
        rx.close()


@pytest.mark.parametrize("bits,channels", [(8, 1), (24, 1), (8, 2), (24, 2)])
def test_float_sample_formats(oracle, bits, channels):
    """8- and 24-bit recordings (Makefile:14 tests the reference with 8 bits) enter as floats scaled like DSP::ReadWAV<float>
    (v / (2^(bits-1) - 1), decode.cc:576) instead of being re-quantised to 16 bits: payload, sync and LLRs follow the oracle
    fed with the same floats."""
    import modem_b200 as M
    from test_gpu_parity import TOL_CONS, llr_close
    fac = float((1 << (bits - 1)) - 1)
    imp = oracle.impair(awgn_db=-22, seed=bits) if channels == 2 else None
    p16 = oracle.encode(oracle.make_payload(80 + bits), channels=channels, imp=imp).astype(np.float64) / 32767.0
    q = np.rint(np.clip(p16, -1, 1) * fac)                       # what `encode OUT 8000 <bits> ...` writes
    samples = (q.astype(np.float32) / np.float32(fac)).astype(np.float32).reshape(-1)
    rx = M.Receiver(max_frames=1, keep_taps=True)
    try:
        payload, st = rx.decode(samples.reshape(1, -1), channels=channels)
        ost, opay, tp = oracle.decode_f32(samples, channels=channels)
        assert st["status"][0] == ost == 0 and (payload[0] == opay).all()
        assert (st["sc_pos"][0], st["shift"][0], st["mode"][0]) == (tp.sc_pos, tp.shift, tp.mode)
        iq = rx.taps(M.TAP_IQ, 0, 1)[0].view(np.complex64)[: samples.size // channels + 1]
        oiq, otm = oracle.front_taps(samples, channels=channels)
        assert np.abs(iq - oiq).max() < TOL_IQ
        ollr = oracle.taps_np(tp, "llr")
        llr_close(rx.taps(M.TAP_LLR, 0, 1)[0], ollr)
        assert np.abs(rx.taps(M.TAP_CONS_RAW, 0, 1, 6)[0] - oracle.taps_np(tp, "cons_raw")).max() < TOL_CONS
    finally:
        rx.close()


def test_skip_walks_past_sixteen_detections(oracle):
    """24 frames in one recording, SKIP = 20 (and 23, and one past the end): the detection list is sized from the window
    length, not capped at 16 as in round 1; status, detection count and payload equal the oracle's."""
    import modem_b200 as M
    pls = np.stack([oracle.make_payload(900 + i) for i in range(24)])
    pcm = oracle.encode(pls)
    rx = M.Receiver(max_frames=1, max_samples=pcm.shape[0])
    try:
        for skip in (0, 15, 16, 20, 23, 24):
            payload, st = rx.decode(pcm.reshape(1, -1), skip=skip)
            ost, opay, tp = oracle.decode(pcm, skip=skip, want_taps=True)
            assert st["status"][0] == ost and st["detections"][0] == tp.detections and st["det_overflow"][0] == 0, skip
            assert (payload[0] == opay).all(), skip
            if skip < 24:
                assert ost == 0 and (payload[0] == pls[skip]).all(), skip
    finally:
        rx.close()

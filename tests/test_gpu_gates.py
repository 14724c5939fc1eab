"""The north-star gates of BASELINE.json, sized for the driver's `-m gpu` run (the full-size versions are tools/clean100k_check.py,
tools/ber_sweep.py and bench.py's config5; their results are under profiles/):
  * clean channel: 20 000 DISTINCT mode-6 frames generated on the device (include/ofdmtx.h) decode with 0 payload bit errors;
  * AWGN sweep across the waterfall (-15.75 ... -14.0 dB, 8 points x 200 windows): every window ends like the CPU oracle's
    decode of the same samples, so frame and bit error rates per point are identical (the +-0.1 dB gate holds with margin 0);
  * mixed impairments: windows with different impairment classes in ONE batch decode to the sent bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gate_clean_20k_distinct_frames(oracle):
    import torch
    import modem_b200 as M
    n, chunk = 20000, 5000
    tx = M.Transmitter(max_windows=chunk)
    rx = M.Receiver(max_frames=chunk)
    stride = tx.window_samples(6)
    cs = int(M.load().ofdmtx_call_sign(b"CALLSIGN"))
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(20000)
    errors = flips = not_ok = 0
    pcm = torch.zeros((chunk, stride), dtype=torch.int16, device="cuda")
    pay = torch.empty((chunk, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
    st = torch.empty((chunk, 112), dtype=torch.uint8, device="cuda")
    try:
        for c in range(n // chunk):
            sent = torch.randint(0, 256, (chunk, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda", generator=g)
            tx.encode_raw(sent.data_ptr(), M.MEM_DEVICE, chunk, 6, cs, 2000, None, pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_MONO, stride, None, stream)
            rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_MONO, chunk, stride, None, 0, pay.data_ptr(), st.data_ptr(), stream)
            torch.cuda.synchronize()
            stat = st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
            not_ok += int((stat["status"] != 0).sum())
            flips += int(stat["flips"].clip(0).sum())
            x = (pay ^ sent).cpu().numpy()
            errors += int(np.unpackbits(x).sum())
            if c == 0:   # the device-generated windows are the reference transmitter's: the CPU oracle decodes them to the same bytes
                host = pcm[:4].cpu().numpy()
                for i in range(4):
                    ost, opay, _ = oracle.decode(host[i], want_taps=False)
                    assert ost == 0 and (opay == sent[i].cpu().numpy()).all()
    finally:
        tx.close(); rx.close()
    assert (errors, not_ok, flips) == (0, 0, 0)


def test_gate_awgn_sweep_equals_oracle(oracle):
    import modem_b200 as M
    points = [-15.75, -15.5, -15.25, -15.0, -14.75, -14.5, -14.25, -14.0]
    per = 200
    rx = M.Receiver(max_frames=per)
    fer_gpu, fer_cpu = [], []
    try:
        for k, db in enumerate(points):
            pcm, ns, sent = oracle.encode_batch(per, seed0=70000 + 1000 * k, channels=2, imp=oracle.impair(awgn_db=db, seed=900 + k))
            payload, st = rx.decode(pcm, channels=2)
            ost, opay = oracle.decode_batch(pcm, channels=2)
            assert (st["status"] == ost).all(), (db, np.nonzero(st["status"] != ost)[0][:8])
            assert (payload == opay).all(), db
            okg = st["status"] == 0
            assert (payload[okg] == sent[okg]).all()           # a CRC-32 match that is not the sent payload: not seen
            fer_gpu.append(1.0 - okg.mean()); fer_cpu.append(1.0 - (ost == 0).mean())
    finally:
        rx.close()
    assert fer_gpu == fer_cpu
    assert fer_gpu[0] < 0.05 and fer_gpu[-1] > 0.5, fer_gpu      # the sweep does straddle the waterfall (noise LEVEL in dB: more is worse)


def test_gate_mixed_impairments_in_one_batch(oracle):
    """BASELINE configs[4] in miniature: clean, AWGN, CFO, multipath, SFO and the full README chain side by side in one
    decode call (windows of 95 200 + slack samples, analytic int16): every payload equals the sent bytes and the oracle's."""
    import modem_b200 as M
    classes = [None, dict(awgn_db=-22.0), dict(cfo_hz=-180.5), dict(multipath=True), dict(sfo_ppm=-120.0), dict(sfo_ppm=147.0),
               dict(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0), dict(multipath=True, cfo_hz=-31.25, awgn_db=-24.0)]
    per, stride = 6, 95200 + 64
    tx = M.Transmitter(max_windows=per)
    rx = M.Receiver(max_frames=per * len(classes), max_samples=stride)
    try:
        rng = np.random.default_rng(44)
        sent = rng.integers(0, 256, (per * len(classes), M.PAYLOAD_BYTES), dtype=np.uint8)
        pcm = np.zeros((per * len(classes), 2 * stride), np.int16)
        ns = np.zeros(per * len(classes), np.int32)
        for c, kw in enumerate(classes):
            imp = M.impairments(seed=500 + c, **kw) if kw else None
            p, n1 = tx.encode(sent[c * per:(c + 1) * per], channels=2, imp=imp, stride=stride)
            pcm[c * per:(c + 1) * per], ns[c * per:(c + 1) * per] = p, n1
        payload, st = rx.decode(pcm, channels=2, n_samples=ns)
        assert (st["status"] == 0).all(), st["status"]
        assert (payload == sent).all()
        for i in range(0, per * len(classes), per):     # one window per class through the CPU oracle as well
            ost, opay, _ = oracle.decode(pcm[i, :2 * ns[i]], channels=2, want_taps=False)
            assert ost == 0 and (opay == sent[i]).all()
    finally:
        tx.close(); rx.close()

"""Device-side stimulus generator (include/ofdmtx.h) on the B200, through the C-ABI: transmitted code bits
bit-exact against the oracle, sample streams within one LSB of the oracle's transmitter + impairment chain, noise
statistics, and the loop-back gate — windows generated on the device decode on the device to the payloads that went in,
with the CPU oracle agreeing on the same windows."""
import os
import subprocess

import numpy as np
import pytest

# First hardware run: round 2 (profiles/r2a_stimulus_first_run.md): memcheck clean, device loop-back 2048 / 2048.
pytestmark = pytest.mark.gpu


def close_pcm(got, ref):
    got, ref = got.reshape(-1).astype(int), ref.reshape(-1).astype(int)
    assert got.size == ref.size
    d = np.abs(got - ref)
    assert d.max() <= 1, d.max()
    assert (d != 0).mean() < 5e-3


@pytest.fixture(scope="module")
def tx():
    import modem_b200 as M
    t = M.Transmitter(max_windows=64)
    yield t
    t.close()


def test_code_bits_and_clean_windows(tx, oracle):
    for mode in (6, 9, 10, 13):
        pls = np.stack([oracle.make_payload(100 * mode + i) for i in range(3)])
        for ch in (1, 2):
            pcm, ns = tx.encode(pls, mode=mode, channels=ch)
            assert (ns == oracle.frame_samples(mode)).all()
            for i in range(3):
                close_pcm(pcm[i], oracle.encode(pls[i], mode=mode, channels=ch))
        code = tx.code_bits(0, 3)
        nb = 64800 if mode < 10 else 64512
        for i in range(3):
            ref = np.zeros(65536, np.uint8)
            oracle.lib().ref_payload_to_code(pls[i].ctypes.data, mode, ref.ctypes.data)
            assert (np.unpackbits(code[i].view(np.uint8), bitorder="little")[:nb] == ref[:nb]).all()


@pytest.mark.parametrize("kw", [dict(multipath=True), dict(cfo_hz=234.567), dict(sfo_ppm=147.0), dict(sfo_ppm=-80.0),
                                dict(multipath=True, cfo_hz=-31.25, sfo_ppm=147.0)])
def test_deterministic_impairments(tx, oracle, kw):
    import modem_b200 as M
    pl = oracle.make_payload(9)
    pcm, ns = tx.encode(pl, channels=2, imp=M.impairments(**kw))
    ref = oracle.encode(pl, channels=2, imp=oracle.impair(**kw))
    assert ns[0] == ref.shape[0]
    close_pcm(pcm[0, :2 * ns[0]], ref)
    assert not pcm[0, 2 * ns[0]:].any()


@pytest.mark.parametrize("rate", [16000, 44100, 48000])
def test_other_sample_rates(oracle, rate):
    import modem_b200 as M
    t = M.Transmitter(max_windows=2, rate=rate)
    pls = np.stack([oracle.make_payload(rate + i) for i in range(2)])
    pcm, ns = t.encode(pls, channels=2)
    for i in range(2):
        close_pcm(pcm[i], oracle.encode(pls[i], rate=rate, channels=2))
    t.close()


def test_frames_back_to_back(oracle):
    import modem_b200 as M
    t = M.Transmitter(max_windows=2, frames_per_window=3)
    pls = np.stack([oracle.make_payload(40 + i) for i in range(6)])
    pcm, ns = t.encode(pls.reshape(2, -1))
    close_pcm(pcm[0], oracle.encode(pls[:3]))
    close_pcm(pcm[1], oracle.encode(pls[3:]))
    t.close()


def test_noise_statistics(tx, oracle):
    import modem_b200 as M
    pls = np.stack([oracle.make_payload(5)] * 2)
    clean, _ = tx.encode(pls, fmt=M.FMT_F32_IQ)
    imp = M.impairments(awgn_db=-20.0, seed=11)
    noisy, _ = tx.encode(pls, fmt=M.FMT_F32_IQ, imp=imp)
    z = noisy - clean
    for v in (z[0].real, z[0].imag, z[1].real):
        assert abs(v.var() / 0.005 - 1) < 0.03 and abs(v.mean()) < 1e-3
    assert abs(np.corrcoef(z[0].real, z[0].imag)[0, 1]) < 0.02
    assert abs(np.corrcoef(z[0].real, z[1].real)[0, 1]) < 0.02
    again, _ = tx.encode(pls, fmt=M.FMT_F32_IQ, imp=imp)
    assert (again == noisy).all()
    # chunking must not change which stream a window draws from
    t2 = M.Transmitter(max_windows=1)
    chunked, _ = t2.encode(pls, fmt=M.FMT_F32_IQ, imp=imp)
    t2.close()
    assert (chunked == noisy).all()


def test_device_loopback_and_oracle_agreement(tx, oracle):
    """generate on the device -> decode on the device (device pointers, no host copy of the samples) == the payloads;
    the CPU oracle decodes a sample of the same windows to the same bytes"""
    import torch
    import modem_b200 as M
    n = 64
    pls = np.stack([oracle.make_payload(9000 + i) for i in range(n)])
    imp = M.impairments(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=77)   # README.md:49
    stride = tx.window_samples(6)
    d_pcm = torch.zeros((n, stride * 2), dtype=torch.int16, device="cuda")
    ns = np.zeros(n, np.int32)
    cs = int(M.load().ofdmtx_call_sign(b"CALLSIGN"))
    s = torch.cuda.current_stream().cuda_stream
    tx.encode_raw(pls.ctypes.data, M.MEM_HOST, n, 6, cs, 2000, imp, d_pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, stride, ns, s)
    rx = M.Receiver(max_frames=n, max_samples=stride)
    d_pay = torch.zeros((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros((n, M.STATUS_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    rx.decode_raw(d_pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, n, stride, ns, 0, d_pay.data_ptr(), d_st.data_ptr(), s)
    torch.cuda.synchronize()
    pay = d_pay.cpu().numpy()
    st = d_st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
    assert (st["status"] == 0).all() and (pay == pls).all()
    pcm = d_pcm.cpu().numpy()
    for i in (0, 31, 63):
        ost, opay, _ = oracle.decode(pcm[i, :2 * ns[i]].reshape(-1, 2), channels=2, want_taps=False)
        assert ost == 0 and (opay == pay[i]).all()
    rx.close()


def test_encode_cli_matches_reference_contract(oracle, tmp_path):
    """`encode OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT..` (encode.cc:337-446) next to the oracle's CLI"""
    import modem_b200 as M
    from modem_b200 import build as B
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "build", "encode_ref")
    names = []
    for i in range(2):
        (tmp_path / ("in%d.dat" % i)).write_bytes(oracle.make_payload(300 + i).tobytes())
        names.append(str(tmp_path / ("in%d.dat" % i)))
    for bits, ch in ((16, 1), (16, 2), (8, 1)):
        args = ["8000", str(bits), str(ch), "2000", "6", "CALLSIGN"] + names
        subprocess.run([B.ENCODE, str(tmp_path / "gpu.wav")] + args, check=True)
        subprocess.run([ref, str(tmp_path / "cpu.wav")] + args, check=True)
        g, c = (tmp_path / "gpu.wav").read_bytes(), (tmp_path / "cpu.wav").read_bytes()
        assert len(g) == len(c) and g[:44] == c[:44]
        if bits == 16:
            close_pcm(np.frombuffer(g[44:], "<i2"), np.frombuffer(c[44:], "<i2"))
        else:
            close_pcm(np.frombuffer(g[44:], np.uint8), np.frombuffer(c[44:], np.uint8))
    r = subprocess.run([B.ENCODE, "-", "8000", "16", "1", "2025", "6", "CALLSIGN", names[0]], capture_output=True)
    assert r.returncode == 1 and b"divisible by 50" in r.stderr
    r = subprocess.run([B.ENCODE, "-", "8000", "16", "1", "2000", "5", "CALLSIGN", names[0]], capture_output=True)
    assert r.returncode == 1 and b"Unsupported operation mode." in r.stderr
    # and the device decoder reads what the device encoder wrote
    subprocess.run([B.ENCODE, str(tmp_path / "gpu.wav"), "8000", "16", "1", "2000", "6", "CALLSIGN"] + names, check=True)
    subprocess.run([B.DECODE, str(tmp_path / "out.dat"), str(tmp_path / "gpu.wav"), "1"], check=True)
    assert (tmp_path / "out.dat").read_bytes() == oracle.make_payload(301).tobytes()

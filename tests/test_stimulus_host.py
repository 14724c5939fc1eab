"""The stimulus generator's arithmetic (modem_b200/csrc/stimulus.cuh — the routines the k_tx_* kernels call) compiled for
the host by tests/stimulus_host.cu and checked against the oracle's transmitter and impairment chain.  No GPU needed: this
is the CPU-side proof that the device code computes the right thing; tests/test_gpu_stimulus.py repeats it on the B200."""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def stim(oracle):
    from modem_b200 import build as B
    B.build()
    L = C.CDLL(B.STIMHOST)
    L.stimulus_host.restype = C.c_longlong
    L.stimulus_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p, C.c_longlong]
    L.stimulus_host_code.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    cs = oracle.lib().ref_base37(b"CALLSIGN")

    def run(payloads, mode=6, rate=8000, fmt=0, imp=None, fpw=1, window=0, freq_off=2000):
        payloads = np.ascontiguousarray(payloads, np.uint8)
        stride = 2 * rate + (2 + fpw * (3 + oracle.MODE_ROWS[mode])) * (1440 * rate // 8000) + 64
        out = np.zeros(stride, np.complex64) if fmt == 2 else np.zeros(stride * (1 if fmt == 0 else 2), np.int16)
        n = L.stimulus_host(rate, mode, freq_off, cs, payloads.ctypes.data, fpw, C.byref(imp) if imp is not None else None,
                            window, fmt, out.ctypes.data, stride)
        assert n > 0, n
        return out[: n * (2 if fmt == 1 else 1)]
    run.lib = L
    return run


def close_pcm(got, ref):
    """equal up to the rounding of a few samples: the two sides run different FFT factorizations in fp32"""
    got, ref = got.reshape(-1).astype(int), ref.reshape(-1).astype(int)
    assert got.size == ref.size
    d = np.abs(got - ref)
    assert d.max() <= 1, d.max()
    assert (d != 0).mean() < 5e-3


@pytest.mark.parametrize("mode,table", [(6, 0), (9, 0), (10, 1), (13, 1)])
def test_code_bits_are_exact(stim, oracle, mode, table):
    nb = 64800 if table == 0 else 64512
    for seed in (1, 2):
        pl = oracle.make_payload(1000 * mode + seed)
        code = np.zeros(2048, np.uint32)
        stim.lib.stimulus_host_code(pl.ctypes.data, table, code.ctypes.data)
        ref = np.zeros(65536, np.uint8)
        oracle.lib().ref_payload_to_code(pl.ctypes.data, mode, ref.ctypes.data)
        assert (np.unpackbits(code.view(np.uint8), bitorder="little")[:nb] == ref[:nb]).all()


@pytest.mark.parametrize("mode", [6, 8, 11, 13])
def test_clean_windows_match_the_oracle_encoder(stim, oracle, mode):
    pl = oracle.make_payload(mode)
    close_pcm(stim(pl, mode=mode, fmt=0), oracle.encode(pl, mode=mode, channels=1))
    close_pcm(stim(pl, mode=mode, fmt=1), oracle.encode(pl, mode=mode, channels=2))


@pytest.mark.parametrize("rate", [16000, 44100, 48000])
def test_other_sample_rates(stim, oracle, rate):
    # the oracle runs the literal 4N-point transforms of encode.cc:80-100; the product's polyphase form must agree
    pl = oracle.make_payload(rate)
    close_pcm(stim(pl, rate=rate, fmt=1), oracle.encode(pl, rate=rate, channels=2))


def test_frames_back_to_back_and_other_offsets(stim, oracle):
    pls = np.stack([oracle.make_payload(40 + i) for i in range(3)])
    close_pcm(stim(pls, fpw=3), oracle.encode(pls))                      # encode.cc:289: one stream, three frames
    pl = oracle.make_payload(77)
    close_pcm(stim(pl, fmt=1, freq_off=-450), oracle.encode(pl, channels=2, freq_off=-450))
    close_pcm(stim(pl, fmt=0, freq_off=1350), oracle.encode(pl, channels=1, freq_off=1350))


@pytest.mark.parametrize("kw", [dict(multipath=True), dict(cfo_hz=234.567), dict(sfo_ppm=147.0), dict(sfo_ppm=-80.0),
                                dict(multipath=True, cfo_hz=-31.25, sfo_ppm=147.0)])
def test_deterministic_impairments_match_the_oracle(stim, oracle, kw):
    imp = oracle.impair(**kw)
    pl = oracle.make_payload(9)
    close_pcm(stim(pl, fmt=1, imp=imp), oracle.encode(pl, channels=2, imp=imp))


def test_noise_statistics(stim, oracle):
    pl = oracle.make_payload(5)
    clean = stim(pl, fmt=2)
    imp = oracle.impair(awgn_db=-20.0, seed=11)
    z0, z1 = stim(pl, fmt=2, imp=imp, window=0) - clean, stim(pl, fmt=2, imp=imp, window=1) - clean
    for z in (z0.real, z0.imag, z1.real):
        assert abs(z.var() / 0.005 - 1) < 0.03 and abs(z.mean()) < 1e-3
    assert abs(np.corrcoef(z0.real, z0.imag)[0, 1]) < 0.02
    assert abs(np.corrcoef(z0.real, z1.real)[0, 1]) < 0.02          # windows draw from different streams
    assert abs(np.corrcoef(z0.real[:-1], z0.real[1:])[0, 1]) < 0.02
    assert (stim(pl, fmt=2, imp=imp, window=0) - clean == z0).all()   # and the stream is reproducible
    k = ((z0.real / np.sqrt(0.005)) ** 4).mean()
    assert abs(k - 3.0) < 0.1


def test_generated_windows_decode(stim, oracle):
    imp = oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=3)   # README.md:49
    for mode in (6, 12):
        pl = oracle.make_payload(mode)
        pcm = stim(pl, mode=mode, fmt=1, imp=imp)
        st, pay, _ = oracle.decode(pcm.reshape(-1, 2), channels=2, want_taps=False)
        assert st == 0 and (pay == pl).all()


def test_argument_rules(stim, oracle):
    L = stim.lib
    out = np.zeros(200000, np.int16)
    pl = oracle.make_payload(1)
    cs = oracle.lib().ref_base37(b"CALLSIGN")
    def rc(rate=8000, mode=6, off=2000, call=cs, fmt=0):
        return L.stimulus_host(rate, mode, off, call, pl.ctypes.data, 1, None, 0, fmt, out.ctypes.data, 100000)
    assert rc() == 95200
    assert rc(mode=5) == -22 and rc(mode=14) == -22          # encode.cc:353-356
    assert rc(call=0) == -22 and rc(call=37 ** 9) == -22      # encode.cc:357-361
    assert rc(off=2025) == -22                                # encode.cc:394-397
    assert rc(off=1300) == -22                                # real output: offset below half the band width (encode.cc:389)
    assert rc(off=2700) == -22                                # above rate/2 - bw/2
    assert rc(off=-1000, fmt=1) == 95200                      # analytic output may sit at negative offsets
    assert rc(rate=22050) == -22


def test_no_barrier_is_missing(stim):
    """the CTA-cooperative routines (payload -> code bits, one OFDM symbol, the shared Stockham FFT) on six host threads with
    OFDMRX_CTA_SYNC() as a real barrier, under ThreadSanitizer (tests/stimulus_cta_tsan.cu): no data race, and the same bits
    as the one-thread run.  (Dropping any one barrier from stimulus.cuh makes this fail — tried.)"""
    import subprocess
    from modem_b200 import build as B
    r = subprocess.run([B.STIMTSAN, "6"], capture_output=True, text=True)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr and "0 mismatching" in r.stdout, (r.stdout, r.stderr[-2000:])

"""ctypes binding of the CPU oracle (oracle/build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DATA_BYTES = 5380
FRAME_SAMPLES = 95200  # 1 s silence + 55 symbols x 1440 + 1 s silence at 8 kHz (encode.cc:288,311-313,423,441)


class Impair(C.Structure):
    _fields_ = [("multipath", C.c_int32), ("cfo_hz", C.c_float), ("sfo_ppm", C.c_float),
                ("awgn", C.c_int32), ("awgn_db", C.c_float), ("seed", C.c_uint64)]


class Taps(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("status", "detections", "t_fire", "symbol_pos", "sc_pos", "index_max", "shift",
                                         "pos_err", "osd_unique", "mode", "best_lane", "flips", "rows", "cols")] + \
               [("timing_max", C.c_float), ("frac_cfo", C.c_float), ("cfo_rad", C.c_float),
                ("md", C.c_int64), ("forks", C.c_int64), ("osd_visited", C.c_int64),
                ("soft", C.c_int8 * 256), ("hdr", C.c_uint8 * 32), ("call_sign", C.c_char * 12),
                ("metrics", C.c_float * 8),
                ("slope", C.c_float * 128), ("yint", C.c_float * 128), ("precision", C.c_float * 128),
                ("cons_raw", C.c_float * (2 * 32400)), ("cons", C.c_float * (2 * 32400)),
                ("llr", C.c_float * 65536)]


def build(fast=False):
    """(Re)build the oracle with its Makefile; building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_libs = {}


def lib(fast=False):
    key = "fast" if fast else "safe"
    if key in _libs:
        return _libs[key]
    path = os.path.join(ORACLE_DIR, "build", "liboracle_fast.so" if fast else "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.ref_encode_pcm16.restype = C.c_int64
    L.ref_encode_pcm16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    L.ref_make_payload.argtypes = [C.c_uint64, C.c_void_p]
    L.ref_encode_batch_pcm16.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    L.ref_decode_pcm16.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_decode_batch_pcm16.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int]
    L.ref_front_taps.restype = C.c_int64
    L.ref_front_taps.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_decode_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_polar_decode_any.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_mls.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.ref_crc16_u64.restype = C.c_uint32
    L.ref_crc16_u64.argtypes = [C.c_uint64]
    L.ref_crc32_bytes.restype = C.c_uint32
    L.ref_crc32_bytes.argtypes = [C.c_void_p, C.c_int]
    L.ref_crc32_bits.restype = C.c_uint32
    L.ref_crc32_bits.argtypes = [C.c_void_p, C.c_int]
    L.ref_xorshift.argtypes = [C.c_int, C.c_void_p]
    L.ref_base37.restype = C.c_int64
    L.ref_base37.argtypes = [C.c_char_p]
    L.ref_frozen_table.argtypes = [C.c_int, C.c_void_p]
    L.ref_bch_generator.argtypes = [C.c_void_p]
    L.ref_bch_genmat.argtypes = [C.c_void_p]
    L.ref_bch_encode.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_osd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_fft.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_theil_sen.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_payload_to_code.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_polar_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.ref_taps_size() == C.sizeof(Taps), (L.ref_taps_size(), C.sizeof(Taps))
    _libs[key] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_payload(seed):
    out = np.zeros(DATA_BYTES, np.uint8)
    lib().ref_make_payload(seed, _p(out))
    return out


def impair(multipath=False, cfo_hz=0.0, sfo_ppm=0.0, awgn_db=None, seed=1):
    return Impair(int(multipath), cfo_hz, sfo_ppm, int(awgn_db is not None), awgn_db if awgn_db is not None else 0.0, seed)


def encode(payloads, rate=8000, channels=1, freq_off=2000, call_sign=b"CALLSIGN", mode=6, imp=None):
    payloads = np.ascontiguousarray(payloads, np.uint8).reshape(-1, DATA_BYTES)
    count = payloads.shape[0]
    cap = 2 * rate + (4 + 130 * count) * (1440 * rate // 8000) + 64  # up to 126 rows + 3 per frame (mode 13)
    out = np.zeros(cap * channels, np.int16)
    n = lib().ref_encode_pcm16(_p(payloads), count, rate, channels, freq_off, call_sign, mode,
                               C.byref(imp) if imp is not None else None, _p(out), cap)
    if n < 0:
        raise ValueError("encode rejected the arguments (%d)" % n)
    return out[: n * channels].reshape(n, channels) if channels == 2 else out[:n]


# payload symbols per mode (decode.cc:302-374): a single-frame stream is 1 s + (5 + rows) symbols + 1 s (encode.cc:288-313,423,441)
MODE_ROWS = {6: 50, 7: 54, 8: 81, 9: 90, 10: 42, 11: 56, 12: 84, 13: 126}


def frame_samples(mode=6, rate=8000):
    return 2 * rate + (5 + MODE_ROWS[mode]) * (1440 * rate // 8000)


def encode_batch(n, seed0=0, rate=8000, channels=1, freq_off=2000, call_sign=b"CALLSIGN", mode=6, imp=None,
                 stride=None, nthreads=None):
    """n single-frame windows of `stride` sample frames; payload i = make_payload(seed0 + i)."""
    nthreads = nthreads or os.cpu_count() or 1
    stride = stride or frame_samples(mode, rate)
    pcm = np.zeros((n, stride * channels), np.int16)
    ns = np.zeros(n, np.int32)
    pay = np.zeros((n, DATA_BYTES), np.uint8)
    r = lib().ref_encode_batch_pcm16(n, seed0, rate, channels, freq_off, call_sign, mode,
                                     C.byref(imp) if imp is not None else None, _p(pcm), stride, _p(ns), _p(pay), nthreads)
    if r != 0:
        raise ValueError("batch encode failed")
    return pcm, ns, pay


def decode(pcm, channels=1, rate=8000, skip=0, list_size=8, r0_max=1 << 16, osd_literal=False, want_taps=True, fast=False):
    pcm = np.ascontiguousarray(pcm, np.int16)
    n = pcm.size // channels
    out = np.zeros(DATA_BYTES, np.uint8)
    taps = Taps() if want_taps else None
    st = lib(fast).ref_decode_pcm16(_p(pcm), n, channels, rate, skip, list_size, r0_max, int(osd_literal), _p(out),
                                    C.byref(taps) if taps is not None else None)
    return st, out, taps


def decode_f32(samples, channels=1, rate=8000, skip=0):
    """float samples as DSP::ReadWAV<float> delivers them (any bit depth) -> (status, payload, taps)"""
    samples = np.ascontiguousarray(samples, np.float32)
    out = np.zeros(DATA_BYTES, np.uint8)
    taps = Taps()
    st = lib().ref_decode_f32(_p(samples), samples.size // channels, channels, rate, skip, _p(out), C.byref(taps))
    return st, out, taps


def front_taps(pcm, channels=1, rate=8000):
    """Stream taps of the front end for one window (int16 or float32 samples): (iq complex64 [n + 1], timing float32 [n + 1]),
    one entry per stream step — what next_sample() pushed (decode.cc:294-301) and the timing metric of decode.cc:90."""
    pcm = np.asarray(pcm)
    is_float = pcm.dtype.kind == "f"
    pcm = np.ascontiguousarray(pcm, np.float32 if is_float else np.int16)
    n = pcm.size // channels
    iq = np.zeros(n + 1, np.complex64)
    timing = np.zeros(n + 1, np.float32)
    got = lib().ref_front_taps(None if is_float else _p(pcm), _p(pcm) if is_float else None, n, channels, rate, _p(iq), _p(timing))
    assert got == n + 1, (got, n)
    return iq, timing


def decode_batch(pcm, n_samples=None, channels=1, rate=8000, skip=0, list_size=8, nthreads=None, fast=False):
    pcm = np.ascontiguousarray(pcm, np.int16)
    n = pcm.shape[0]
    stride = pcm.shape[1] // channels
    nthreads = nthreads or os.cpu_count() or 1
    out = np.zeros((n, DATA_BYTES), np.uint8)
    st = np.zeros(n, np.int32)
    ns = np.ascontiguousarray(n_samples, np.int32) if n_samples is not None else None
    lib(fast).ref_decode_batch_pcm16(_p(pcm), n, stride, _p(ns), channels, rate, skip, list_size, _p(out), _p(st), nthreads)
    return st, out


def polar_decode(llr, table=0, list_size=8, r0_max=1 << 16):
    llr = np.ascontiguousarray(llr, np.float32)
    lanes = np.zeros((list_size, 65536), np.uint8)
    metrics = np.zeros(list_size, np.float32)
    payload = np.zeros(DATA_BYTES, np.uint8)
    flips = C.c_int32(0)
    best = lib().ref_polar_decode(_p(llr), table, list_size, r0_max, _p(lanes), _p(metrics), _p(payload), C.byref(flips))
    return best, lanes, metrics, payload, flips.value


def taps_np(t, name):
    a = np.ctypeslib.as_array(getattr(t, name))
    if name in ("cons", "cons_raw"):
        n = t.rows * t.cols
        return a[: 2 * n].view(np.complex64).reshape(t.rows, t.cols).copy()
    if name in ("slope", "yint", "precision"):
        return a[: t.rows].copy()
    return a.copy()

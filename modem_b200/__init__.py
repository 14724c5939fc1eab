"""modem_b200 — Python host mirror of the aicodix/modem receive path on B200 (sm_100a).

The reference has no Python API; its interface for this path is `decode OUTPUT INPUT [SKIP]`
(/root/reference/decode.cc:559-620).  This module is a thin ctypes layer over the C-ABI in include/ofdmrx.h
(libofdmrx.so, hand-written CUDA) that mirrors that contract for batches of windows:

    rx = Receiver(max_frames=4096)
    payload, status = rx.decode(pcm_int16, channels=1, skip=0)        # == n x `decode out.dat window_i.wav [SKIP]`
    payload, status = decode_wav("recorded.wav", skip=0)               # one WAV file, like the reference CLI

There is NO CPU fallback: importing works anywhere, but creating a Receiver without the built extension or
without a B200-class GPU raises.
"""
import ctypes as C
import os
import struct

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFDMRX_LIB") or os.path.join(PKG, "libofdmrx.so")  # OFDMRX_LIB: another build of the same library (A/B runs)

PAYLOAD_BYTES = 5380
CODE_LEN = 65536
FRAME_SAMPLES = 95200
FMT_S16_MONO, FMT_S16_IQ, FMT_F32_IQ, FMT_F32_MONO = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
ST_OK, ST_NO_SYNC, ST_OSD_FAIL, ST_HDR_CRC, ST_BAD_MODE, ST_BAD_CALL, ST_PAYLOAD_CRC, ST_UNSUPPORTED_MODE = range(8)
STATUS_TEXT = {  # the reference's stderr strings (decode.cc:419,430,435,440,543)
    ST_OK: "ok", ST_NO_SYNC: "no sync", ST_OSD_FAIL: "OSD error.", ST_HDR_CRC: "header CRC error.",
    ST_BAD_MODE: "operation mode unsupported.", ST_BAD_CALL: "call sign unsupported.",
    ST_PAYLOAD_CRC: "payload decoding error.", ST_UNSUPPORTED_MODE: "operation mode not built.",
}
TAP_IQ, TAP_TIMING, TAP_SOFT, TAP_CONS_RAW, TAP_CONS, TAP_TS, TAP_LLR, TAP_PHASE = range(8)
_TAP_DTYPE = {TAP_IQ: np.complex64, TAP_TIMING: np.float32, TAP_SOFT: np.int8, TAP_CONS_RAW: np.complex64,
              TAP_CONS: np.complex64, TAP_TS: np.float32, TAP_LLR: np.float32, TAP_PHASE: np.float32}

STATUS_DTYPE = np.dtype([
    ("status", "<i4"), ("detections", "<i4"), ("t_fire", "<i4"), ("symbol_pos", "<i4"), ("sc_pos", "<i4"),
    ("index_max", "<i4"), ("shift", "<i4"), ("pos_err", "<i4"), ("timing_max", "<f4"), ("frac_cfo", "<f4"),
    ("cfo_rad", "<f4"), ("osd_unique", "<i4"), ("mode", "<i4"), ("md_lo", "<u4"), ("md_hi", "<u4"),
    ("best_lane", "<i4"), ("flips", "<i4"), ("metrics", "<f4", (8,)), ("osd_visited", "<i4"), ("ts_sweeps", "<i4"), ("det_overflow", "<i4"),
])
assert STATUS_DTYPE.itemsize == 112

# (rows, carriers) of the payload symbols per operation mode (decode.cc:302-374)
MODE_GEOMETRY = {6: (50, 432), 7: (54, 400), 8: (81, 400), 9: (90, 360), 10: (42, 512), 11: (56, 384), 12: (84, 384), 13: (126, 256)}
EXPORTS = ["ofdmrx_create", "ofdmrx_destroy", "ofdmrx_set_option", "ofdmrx_decode_batch", "ofdmrx_polar_decode",
           "ofdmrx_get_taps", "ofdmrx_tap_elems", "ofdmrx_last_launches", "ofdmrx_stage_times", "ofdmrx_get_table",
           "ofdmrx_version", "ofdmrx_theil_sen", "ofdmrx_measure_fp32"]
TX_EXPORTS = ["ofdmtx_create", "ofdmtx_destroy", "ofdmtx_call_sign", "ofdmtx_window_samples", "ofdmtx_encode_batch",
              "ofdmtx_get_code", "ofdmtx_last_launches"]
STAGES = ["frontend", "sync_metric", "sync_detect", "acquire", "demod", "compact_init", "polar_scl", "demod_fft", "theil_sen", "soft_demap"]

_lib = None


class OfdmrxError(RuntimeError):
    pass


def load():
    """dlopen libofdmrx.so (built in-tree by modem_b200/build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OfdmrxError("libofdmrx.so is not built (run `python -m modem_b200.build`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.ofdmrx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.ofdmrx_destroy.argtypes = [C.c_void_p]
    L.ofdmrx_destroy.restype = None
    L.ofdmrx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.ofdmrx_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    L.ofdmrx_polar_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ofdmrx_theil_sen.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ofdmrx_get_taps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    L.ofdmrx_tap_elems.argtypes = [C.c_void_p, C.c_int]
    L.ofdmrx_tap_elems.restype = C.c_int64
    L.ofdmrx_last_launches.argtypes = [C.c_void_p]
    L.ofdmrx_measure_fp32.argtypes = [C.c_void_p, C.c_void_p]
    L.ofdmrx_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ofdmrx_get_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.ofdmrx_version.restype = C.c_char_p
    L.ofdmtx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.ofdmtx_destroy.argtypes = [C.c_void_p]
    L.ofdmtx_destroy.restype = None
    L.ofdmtx_call_sign.argtypes = [C.c_char_p]
    L.ofdmtx_call_sign.restype = C.c_int64
    L.ofdmtx_window_samples.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ofdmtx_window_samples.restype = C.c_int64
    L.ofdmtx_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.ofdmtx_get_code.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ofdmtx_last_launches.argtypes = [C.c_void_p]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise OfdmrxError("%s failed with code %d" % (what, rc))


def base37_decode(val, length=9):
    """decode.cc:155-159"""
    tab = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    out = []
    for _ in range(length):
        out.append(tab[val % 37])
        val //= 37
    return "".join(reversed(out))


def call_sign(status_row):
    md = (int(status_row["md_hi"]) << 32) | int(status_row["md_lo"])
    return base37_decode(md >> 8)


class Receiver:
    """Batched stand-in for the reference's Decoder<float, Complex<float>, RATE> (decode.cc:161-557, instantiated at :590-606)."""

    def __init__(self, device=0, max_frames=1024, max_samples=FRAME_SAMPLES, rate=8000, keep_taps=False, scl_ctas_per_sm=None):
        self._lib = load()
        self._h = C.c_void_p()
        self.device, self.max_frames, self.max_samples = device, max_frames, max_samples
        _check(self._lib.ofdmrx_create(C.byref(self._h), device, rate, max_frames, max_samples), "ofdmrx_create")
        if scl_ctas_per_sm:
            _check(self._lib.ofdmrx_set_option(self._h, b"scl_ctas_per_sm", int(scl_ctas_per_sm)), "set_option")
        if keep_taps:
            _check(self._lib.ofdmrx_set_option(self._h, b"keep_taps", 1), "set_option")

    def set_option(self, key, value):
        """ofdmrx_set_option: "keep_taps", "polar_table", "scl_ctas_per_sm", "sub_chunks" (0 = by batch size, 1 = no overlap of the
        list decoder with the next sub-chunk's front stages: needed for stage_times() on large batches)"""
        _check(self._lib.ofdmrx_set_option(self._h, key.encode() if isinstance(key, str) else key, int(value)), "set_option")

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.ofdmrx_destroy(h)
            h.value = None

    __del__ = close

    # ---- host buffers (numpy; pass pinned memory for asynchronous copies) ------------------------------------
    def decode(self, pcm, channels=1, n_samples=None, skip=0):
        """pcm: [n_windows, stride*channels] host array — int16 (16-bit WAV samples) or float32 (samples already scaled to
        [-1, 1] as DSP::ReadWAV<float> delivers them for 8 / 24 / 32-bit files).  Returns (payload uint8 [n,5380], status)."""
        pcm = np.asarray(pcm)
        is_float = pcm.dtype.kind == "f"
        pcm = np.ascontiguousarray(pcm, dtype=np.float32 if is_float else np.int16)
        if pcm.ndim == 1:
            pcm = pcm.reshape(1, -1)
        n, stride = pcm.shape[0], pcm.shape[1] // channels
        payload = np.empty((n, PAYLOAD_BYTES), np.uint8)
        status = np.zeros(n, STATUS_DTYPE)
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, np.int32)
        fmt = (FMT_F32_MONO if channels == 1 else FMT_F32_IQ) if is_float else (FMT_S16_MONO if channels == 1 else FMT_S16_IQ)
        self.decode_raw(pcm.ctypes.data, MEM_HOST, fmt, n, stride, ns, skip, payload.ctypes.data, status.ctypes.data, None)
        return payload, status

    # ---- raw pointers (device memory from torch: tensor.data_ptr(); stream: torch.cuda.current_stream().cuda_stream)
    def decode_raw(self, samples_ptr, mem_kind, fmt, n_frames, stride, n_samples, skip, payload_ptr, status_ptr, stream):
        ns_ptr = n_samples.ctypes.data if n_samples is not None else None
        _check(self._lib.ofdmrx_decode_batch(self._h, samples_ptr, mem_kind, fmt, n_frames, stride, ns_ptr, skip,
                                             payload_ptr, status_ptr, stream), "ofdmrx_decode_batch")

    def polar_decode(self, llr, want_xbits=False, table=0):
        """llr: float32 [n, 65536] after lengthen(); table 0 = frozen set of modes 6..9, 1 = modes 10..13.
        Returns (payload, status[, xbits uint32 [n,8,2048]])."""
        _check(self._lib.ofdmrx_set_option(self._h, b"polar_table", int(table)), "set_option")
        llr = np.ascontiguousarray(llr, np.float32).reshape(-1, CODE_LEN)
        n = llr.shape[0]
        payload = np.empty((n, PAYLOAD_BYTES), np.uint8)
        status = np.zeros(n, STATUS_DTYPE)
        xb = np.zeros((n, 8, 2048), np.uint32) if want_xbits else None
        _check(self._lib.ofdmrx_polar_decode(self._h, llr.ctypes.data, n, payload.ctypes.data, status.ctypes.data,
                                             xb.ctypes.data if xb is not None else None), "ofdmrx_polar_decode")
        return (payload, status, xb) if want_xbits else (payload, status)

    def theil_sen(self, y):
        """DSP::TheilSenEstimator over rows of phase values at x = i - cols/2 (decode.cc:452,484,488):
        y [n, cols], cols <= 512 -> (slope [n], yint [n])."""
        y = np.ascontiguousarray(y, np.float32)
        assert y.ndim == 2
        n, cols = y.shape
        out = np.zeros((n, 3), np.float32)
        _check(self._lib.ofdmrx_theil_sen(self._h, y.ctypes.data, n, cols, out.ctypes.data), "ofdmrx_theil_sen")
        self.last_sweeps = out[:, 2].astype(int)   # pair sweeps per row (>= 100: the bisection fallback ran)
        return out[:, 0].copy(), out[:, 1].copy()

    def taps(self, stage, first=0, count=1, mode=6):
        per = int(self._lib.ofdmrx_tap_elems(self._h, stage))
        out = np.empty((count, per), _TAP_DTYPE[stage])
        _check(self._lib.ofdmrx_get_taps(self._h, stage, first, count, out.ctypes.data, out.nbytes), "ofdmrx_get_taps")
        if stage in (TAP_CONS_RAW, TAP_CONS, TAP_PHASE):
            rows, cols = MODE_GEOMETRY[mode]   # rows x cols values at the front of the window's 32400 slots
            return out[:, :rows * cols].reshape(count, rows, cols)
        if stage == TAP_TS:
            return out.reshape(count, 126, 3)[:, :MODE_GEOMETRY[mode][0]]
        return out

    def table(self, which):
        n = 2048 if which in (0, 2) else 16384   # 0/2: frozen sets of modes 6..9 / 10..13, 1/3: their SCL schedules
        buf = np.zeros(n, np.uint32)
        got = self._lib.ofdmrx_get_table(self._h, which, buf.ctypes.data, buf.nbytes)
        if got < 0:
            raise OfdmrxError("ofdmrx_get_table failed %d" % got)
        return buf[:got]

    def stage_times(self):
        """CUDA-event durations (ms) of the stages of the last chunk -> (dict, windows in that chunk)."""
        ms = np.zeros(10, np.float32)
        n = self._lib.ofdmrx_stage_times(self._h, ms.ctypes.data, 10)
        if n < 0:
            raise OfdmrxError("ofdmrx_stage_times failed %d" % n)
        return dict(zip(STAGES, [float(x) for x in ms])), int(n)

    def measure_fp32(self):
        """measured FP32 FMA throughput of the device in TFLOP/s (roofline denominator of the list decoder)"""
        v = C.c_float(0)
        _check(self._lib.ofdmrx_measure_fp32(self._h, C.byref(v)), "ofdmrx_measure_fp32")
        return float(v.value)

    @property
    def last_launches(self):
        return int(self._lib.ofdmrx_last_launches(self._h))


class Impairments(C.Structure):
    """ofdmtx_impairments (include/ofdmtx.h): the README.md:49 chain as the oracle re-specifies it."""
    _fields_ = [("multipath", C.c_int32), ("cfo_hz", C.c_float), ("sfo_ppm", C.c_float), ("awgn", C.c_int32),
                ("awgn_db", C.c_float), ("seed", C.c_uint64)]


def impairments(multipath=False, cfo_hz=0.0, sfo_ppm=0.0, awgn_db=None, seed=1):
    return Impairments(int(multipath), cfo_hz, sfo_ppm, int(awgn_db is not None), awgn_db if awgn_db is not None else 0.0, seed)


class Transmitter:
    """Batched stand-in for the reference's Encoder<float, Complex<float>, RATE> (encode.cc:27-318) plus the impairment
    chain of README.md:49 — the device-side stimulus generator."""

    def __init__(self, device=0, max_windows=1024, rate=8000, frames_per_window=1):
        self._lib = load()
        self._h = C.c_void_p()
        self.rate, self.frames_per_window = rate, frames_per_window
        _check(self._lib.ofdmtx_create(C.byref(self._h), device, rate, max_windows, frames_per_window), "ofdmtx_create")

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.ofdmtx_destroy(h)
            h.value = None

    __del__ = close

    def window_samples(self, mode=6):
        return int(self._lib.ofdmtx_window_samples(self.rate, mode, self.frames_per_window))

    def encode(self, payloads, mode=6, call_sign=b"CALLSIGN", freq_off=2000, channels=1, imp=None, stride=None, fmt=None):
        """payloads: uint8 [n_windows (x frames_per_window), 5380] (host).  Returns (samples, n_samples): int16
        [n, stride * channels] like `encode - RATE 16 CHANNELS ...` would write them (fmt=FMT_F32_IQ: complex64 [n, stride])."""
        payloads = np.ascontiguousarray(payloads, np.uint8).reshape(-1, self.frames_per_window * PAYLOAD_BYTES)
        n = payloads.shape[0]
        fmt = (FMT_S16_MONO if channels == 1 else FMT_S16_IQ) if fmt is None else fmt
        if not stride:   # a negative SFO stretches the window: len / (1 + ppm 1e-6) sample frames
            stride = self.window_samples(mode)
            if imp is not None and imp.sfo_ppm < 0:
                stride = int(stride / (1.0 + imp.sfo_ppm * 1e-6)) + 2
        out = np.zeros((n, stride), np.complex64) if fmt == FMT_F32_IQ else np.zeros((n, stride * (1 if fmt == FMT_S16_MONO else 2)), np.int16)
        ns = np.zeros(n, np.int32)
        cs = int(self._lib.ofdmtx_call_sign(call_sign))
        self.encode_raw(payloads.ctypes.data, MEM_HOST, n, mode, cs, freq_off, imp, out.ctypes.data, MEM_HOST, fmt, stride, ns, None)
        return out, ns

    def encode_raw(self, payload_ptr, payload_mem, n_windows, mode, call_sign, freq_off, imp, samples_ptr, mem_kind, fmt, stride,
                   n_samples, stream):
        _check(self._lib.ofdmtx_encode_batch(self._h, payload_ptr, payload_mem, n_windows, mode, call_sign, freq_off,
                                             C.byref(imp) if imp is not None else None, samples_ptr, mem_kind, fmt, stride,
                                             n_samples.ctypes.data if n_samples is not None else None, stream), "ofdmtx_encode_batch")

    def code_bits(self, first=0, count=1):
        """transmitted code words of the last chunk: uint32 [count, 2048]"""
        out = np.zeros((count, 2048), np.uint32)
        _check(self._lib.ofdmtx_get_code(self._h, first, count, out.ctypes.data), "ofdmtx_get_code")
        return out

    @property
    def last_launches(self):
        return int(self._lib.ofdmtx_last_launches(self._h))


def read_wav(path_or_bytes):
    """Minimal RIFF/WAVE PCM reader (what DSP::ReadWAV<float> delivers, decode.cc:576): returns (rate, channels, samples [n, ch]) —
    int16 for 16-bit files (the device scales them by 1/32767), float32 = v / (2^(bits-1) - 1) for 8 / 24 / 32 bits."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    o, fmt = 12, None
    while o + 8 <= len(data):
        cid, sz = data[o:o + 4], struct.unpack("<I", data[o + 4:o + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", data[o + 8:o + 24])
        elif cid == b"data":
            if fmt is None or fmt[0] != 1:
                raise ValueError("PCM WAV expected")
            ch, rate, bits = fmt[1], fmt[2], fmt[5]
            raw = data[o + 8:] if sz in (0, 0xFFFFFFFF) else data[o + 8:o + 8 + sz]
            nb = bits // 8
            if nb < 1 or nb > 4:
                raise ValueError("8, 16, 24 or 32-bit PCM expected")
            raw = raw[:len(raw) // (nb * ch) * nb * ch]
            if bits == 16:
                a = np.frombuffer(raw, "<i2").astype(np.int16)
            else:
                b = np.frombuffer(raw, np.uint8).reshape(-1, nb).astype(np.int64)
                v = sum(b[:, k] << (8 * k) for k in range(nb))
                v = v - 128 if nb == 1 else np.where(v >= 1 << (bits - 1), v - (1 << bits), v)
                a = (v.astype(np.float32) / np.float32((1 << (bits - 1)) - 1)).astype(np.float32)
            return rate, ch, a.reshape(-1, ch)
        o += 8 + sz + (sz & 1)
    raise ValueError("no data chunk")


def decode_wav(path, skip=0, device=0):
    """`decode OUTPUT INPUT [SKIP]` for one file: returns (5380 payload bytes, status record)."""
    rate, ch, pcm = read_wav(path)
    if rate not in (8000, 16000, 44100, 48000):
        raise OfdmrxError("Unsupported sample rate.")  # decode.cc:590-606
    if ch < 1 or ch > 2:
        raise OfdmrxError("Only real or analytic signal (one or two channels) supported.")  # decode.cc:578-581
    rx = Receiver(device=device, max_frames=1, max_samples=max(pcm.shape[0], 1), rate=rate)
    try:
        payload, status = rx.decode(pcm.reshape(1, -1), channels=ch, skip=skip)
    finally:
        rx.close()
    return payload[0].tobytes(), status[0]

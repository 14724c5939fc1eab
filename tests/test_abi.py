"""The C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/ofdmrx.h declares, and
refuses to run without a B200 (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from modem_b200 import build
    build.build()
    import modem_b200 as M
    return M.load()


def test_exports_match_header(lib):
    import modem_b200 as M
    hdr = open(os.path.join(ROOT, "include", "ofdmrx.h")).read()
    declared = set(re.findall(r"\b(ofdmrx_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(M.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ofdmrx_version()
    hdr = open(os.path.join(ROOT, "include", "ofdmtx.h")).read()
    declared = set(re.findall(r"\b(ofdmtx_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(M.TX_EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_status_struct_layout():
    import modem_b200 as M
    hdr = open(os.path.join(ROOT, "include", "ofdmrx.h")).read()
    body = hdr[hdr.index("typedef struct ofdmrx_frame_status {"):hdr.index("} ofdmrx_frame_status;")]
    fields = re.findall(r"\b(?:int32_t|uint32_t|float)\s+([^;]+);", body)
    names = []
    for f in fields:
        for part in f.split(","):
            names.append(re.sub(r"\[.*\]", "", part).strip())
    assert names == list(M.STATUS_DTYPE.names)
    assert M.STATUS_DTYPE.itemsize == 112


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import modem_b200 as M
    h = C.c_void_p()
    assert lib.ofdmrx_create(C.byref(h), 0, 8000, 4, 95200) == -19
    with pytest.raises(M.OfdmrxError):
        M.Receiver(max_frames=4)
    assert lib.ofdmrx_create(C.byref(h), 0, 22050, 4, 95200) == -22    # only the reference's four rates exist (decode.cc:590-606)
    assert lib.ofdmtx_create(C.byref(h), 0, 8000, 4, 1) == -19
    assert lib.ofdmtx_create(C.byref(h), 0, 22050, 4, 1) == -22
    with pytest.raises(M.OfdmrxError):
        M.Transmitter(max_windows=4)
    assert lib.ofdmtx_call_sign(b"CALLSIGN") == 1263905687425 and lib.ofdmtx_call_sign(b"no!") == -1
    assert lib.ofdmtx_window_samples(8000, 6, 1) == 95200


def test_sass_is_sm100_only():
    import subprocess
    import modem_b200 as M
    out = subprocess.run(["cuobjdump", "-lelf", M.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_read_wav(tmp_path, oracle):
    import modem_b200 as M
    import subprocess
    data = os.urandom(5380)
    (tmp_path / "in.dat").write_bytes(data)
    wav = str(tmp_path / "e.wav")
    enc = os.path.join(ROOT, "oracle", "build", "encode_ref")
    for ch in (1, 2):
        subprocess.run([enc, wav, "8000", "16", str(ch), "2000", "6", "CALLSIGN", str(tmp_path / "in.dat")], check=True)
        rate, c, pcm = M.read_wav(wav)
        assert rate == 8000 and c == ch and pcm.shape == (95200, ch) and pcm.dtype == np.int16
    assert M.base37_decode(1263905687425) == " CALLSIGN"

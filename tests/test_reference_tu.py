"""The reference's OWN translation units (decode.cc, encode.cc, freezer.cc compiled where they lie under /root/reference,
oracle/Makefile target `ref`) against the oracle's restatement of them.

The third-party headers those files include (aicodix/dsp, aicodix/code) are absent; oracle/shim/ stands in for them with the
oracle's restated primitives wrapped one to one.  So these tests pin what the reference repository itself holds — control
flow, index arithmetic, constants, bit orders, argv and stderr contracts of decode.cc / encode.cc / freezer.cc — and do NOT
pin the third-party arithmetic.  They need /root/reference and therefore only run in the build container;
tests/golden/reference_tu.json carries their outputs to boxes without it (test_oracle_matches_reference_tu_golden)."""
import hashlib
import json
import os
import subprocess
import wave

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
ORA = os.path.join(ROOT, "oracle", "build")
REFSRC = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REFSRC, "decode.cc")), reason="reference sources are not on this box")


@pytest.fixture(scope="module")
def ref_bins(oracle):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    return REF


def _stderr_lines(b):
    return b.decode().splitlines()


def write_wav(path, pcm, rate, channels):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(pcm, "<i2").tobytes())


def both_decoders(tmp_path, wav, skip=None, lanes=8):
    args = [str(wav)] + ([str(skip)] if skip is not None else [])
    r = subprocess.run([os.path.join(REF, "decode" if lanes == 8 else "decode_l4"), str(tmp_path / "r.dat")] + args, capture_output=True)
    o = subprocess.run([os.path.join(ORA, "decode_ref"), str(tmp_path / "o.dat")] + args, capture_output=True,
                       env=dict(os.environ, REF_LIST=str(lanes)))
    assert r.returncode == o.returncode == 0, (r.stderr, o.stderr)
    rd, od = (tmp_path / "r.dat").read_bytes(), (tmp_path / "o.dat").read_bytes()
    assert len(rd) == len(od) == 5380
    if b"bit flips:" not in r.stderr:
        # a failed decode writes the reference's UNINITIALISED buffer (decode.cc:588, `new uint8_t[data_len]`); the oracle
        # writes the de-scrambled zero buffer instead (deviations) — only the diagnostics are comparable
        rd = od
    return rd, od, _stderr_lines(r.stderr), _stderr_lines(o.stderr)


@needs_reference
def test_freezer_prints_polar_tables(ref_bins):
    """freezer.cc's main() through the restated PolarCodeConst0 prints polar_tables.hh byte for byte"""
    out = subprocess.run([os.path.join(REF, "freezer")], capture_output=True).stdout
    assert out == open(os.path.join(REFSRC, "polar_tables.hh"), "rb").read()


@needs_reference
@pytest.mark.parametrize("rate,bits,ch,off,mode,call,n_in", [
    (8000, 16, 1, 2000, 6, "CALLSIGN", 1),      # README.md:15
    (8000, 8, 1, 2000, 6, "ANONYMOUS", 1),      # Makefile:14
    (8000, 16, 2, 2000, 6, "CALLSIGN", 2),      # README.md:49, two frames in one stream (encode.cc:289)
    (8000, 16, 2, -450, 9, "DL1ABC", 1),
    (8000, 16, 1, 1500, 13, "N0CALL", 1),
    (16000, 16, 1, 3000, 10, "A", 1),
    (44100, 16, 2, 5000, 7, "ZZZZZZZZZ", 1),
    (48000, 24, 1, 1600, 12, "Q 1", 1),
])
def test_encoder_streams_are_identical(ref_bins, oracle, tmp_path, rate, bits, ch, off, mode, call, n_in):
    names = []
    for i in range(n_in):
        (tmp_path / ("in%d.dat" % i)).write_bytes(oracle.make_payload(rate + 10 * mode + i).tobytes())
        names.append(str(tmp_path / ("in%d.dat" % i)))
    args = [str(rate), str(bits), str(ch), str(off), str(mode), call] + names
    r = subprocess.run([os.path.join(REF, "encode"), str(tmp_path / "r.wav")] + args, capture_output=True)
    o = subprocess.run([os.path.join(ORA, "encode_ref"), str(tmp_path / "o.wav")] + args, capture_output=True)
    assert r.returncode == o.returncode == 0, (r.stderr, o.stderr)
    assert (tmp_path / "r.wav").read_bytes() == (tmp_path / "o.wav").read_bytes()


@needs_reference
def test_encoder_argument_errors_match(ref_bins, tmp_path):
    (tmp_path / "in.dat").write_bytes(bytes(5380))
    for args in (["8000", "16", "1", "2000", "5", "CALLSIGN"], ["8000", "16", "1", "2000", "6", "call-sign"], ["8000", "16", "1", "2000", "6", "TENLETTERS"],
                 ["8000", "16", "1", "1300", "6", "CALLSIGN"], ["8000", "16", "2", "2700", "6", "CALLSIGN"], ["8000", "16", "1", "2025", "6", "CALLSIGN"],
                 ["22050", "16", "1", "2000", "6", "CALLSIGN"]):
        r = subprocess.run([os.path.join(REF, "encode"), str(tmp_path / "x.wav")] + args + [str(tmp_path / "in.dat")], capture_output=True)
        o = subprocess.run([os.path.join(ORA, "encode_ref"), str(tmp_path / "y.wav")] + args + [str(tmp_path / "in.dat")], capture_output=True)
        assert r.returncode == o.returncode == 1 and r.stderr == o.stderr, (args, r.stderr, o.stderr)
    r = subprocess.run([os.path.join(REF, "encode")], capture_output=True)
    assert r.returncode == 1 and b"usage:" in r.stderr


@needs_reference
@pytest.mark.parametrize("lanes", [8, 4])
def test_decoder_clean_impaired_and_skip(ref_bins, oracle, tmp_path, lanes):
    pls = np.stack([oracle.make_payload(500 + i) for i in range(3)])
    cases = [("clean mono", oracle.encode(pls[0]), 1, None, pls[0]),
             ("three frames, skip 2", oracle.encode(pls), 1, 2, pls[2]),
             ("three frames, skip 5", oracle.encode(pls), 1, 5, None),
             ("readme chain", oracle.encode(pls[1], channels=2, imp=oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=5)), 2, None, pls[1]),
             ("awgn near threshold", oracle.encode(pls[0], channels=2, imp=oracle.impair(awgn_db=-14.75, seed=8)), 2, None, None),
             ("awgn below threshold", oracle.encode(pls[0], channels=2, imp=oracle.impair(awgn_db=-12.0, seed=9)), 2, None, None),
             ("cut inside the payload", oracle.encode(pls[0])[:40000], 1, None, None),
             ("silence", np.zeros(30000, np.int16), 1, None, None)]
    for name, pcm, ch, skip, want in cases:
        wav = tmp_path / "c.wav"
        write_wav(wav, pcm, 8000, ch)
        rd, od, re_, oe = both_decoders(tmp_path, wav, skip, lanes)
        assert rd == od, name
        assert re_ == oe, (name, re_, oe)
        if want is not None:
            assert rd == want.tobytes(), name


@needs_reference
@pytest.mark.parametrize("rate,mode", [(8000, 7), (8000, 10), (8000, 13), (16000, 8), (44100, 11), (48000, 6)])
def test_decoder_other_modes_and_rates(ref_bins, oracle, tmp_path, rate, mode):
    pl = oracle.make_payload(rate + mode)
    imp = oracle.impair(cfo_hz=-17.5, awgn_db=-28.0, seed=mode)
    wav = tmp_path / "m.wav"
    write_wav(wav, oracle.encode(pl, rate=rate, mode=mode, channels=2, imp=imp), rate, 2)
    rd, od, re_, oe = both_decoders(tmp_path, wav)
    assert rd == od == pl.tobytes() and re_ == oe, (re_, oe)


@needs_reference
def test_decoder_header_failures(ref_bins, oracle, tmp_path):
    """a damaged metadata symbol: both print the same header diagnostics and go on to the next frame (decode.cc:417-442)"""
    pls = np.stack([oracle.make_payload(900 + i) for i in range(2)])
    pcm = oracle.encode(pls).copy()
    pitch = 1440
    start = 8000 + 2 * pitch          # leading pilot, Schmidl-Cox, then the first frame's metadata symbol
    rng = np.random.default_rng(3)
    pcm[start:start + pitch] = rng.integers(-3000, 3000, pitch)
    wav = tmp_path / "h.wav"
    write_wav(wav, pcm, 8000, 1)
    for skip in (None, 1):
        rd, od, re_, oe = both_decoders(tmp_path, wav, skip)
        assert rd == od and re_ == oe, (skip, re_, oe)
    rd, od, re_, oe = both_decoders(tmp_path, wav, 1)
    assert rd == pls[1].tobytes()     # the failed header consumed one SKIP count


@needs_reference
def test_decoder_usage_and_format_errors(ref_bins, tmp_path):
    r = subprocess.run([os.path.join(REF, "decode")], capture_output=True)
    o = subprocess.run([os.path.join(ORA, "decode_ref")], capture_output=True)
    assert r.returncode == o.returncode == 1 and b"usage:" in r.stderr and b"usage:" in o.stderr
    write_wav(tmp_path / "r.wav", np.zeros(1000, np.int16), 22050, 1)
    r = subprocess.run([os.path.join(REF, "decode"), str(tmp_path / "x"), str(tmp_path / "r.wav")], capture_output=True)
    o = subprocess.run([os.path.join(ORA, "decode_ref"), str(tmp_path / "y"), str(tmp_path / "r.wav")], capture_output=True)
    assert r.returncode == o.returncode == 1 and r.stderr == o.stderr == b"Unsupported sample rate.\n"
    write_wav(tmp_path / "c.wav", np.zeros(3000, np.int16), 8000, 3)
    r = subprocess.run([os.path.join(REF, "decode"), str(tmp_path / "x"), str(tmp_path / "c.wav")], capture_output=True)
    o = subprocess.run([os.path.join(ORA, "decode_ref"), str(tmp_path / "y"), str(tmp_path / "c.wav")], capture_output=True)
    assert r.returncode == o.returncode == 1 and r.stderr == o.stderr


def descramble_stream():
    """CODE::Xorshift32 low bytes (decode.cc:613-615): what a zero buffer turns into"""
    y, out = 2463534242, np.zeros(5380, np.uint8)
    for i in range(5380):
        y ^= (y << 13) & 0xFFFFFFFF
        y ^= y >> 17
        y ^= (y << 5) & 0xFFFFFFFF
        out[i] = y & 255
    return out


# ---- the same evidence, portable: outputs of the reference translation units committed by tests/golden/make_reference_tu.py
def golden_cases(oracle):
    """(name, encode kwargs, impairment kwargs, skip) — shared with tests/golden/make_reference_tu.py"""
    return [("clean_mono", dict(seeds=[1]), None, None),
            ("clean_iq_two_frames_skip1", dict(seeds=[2, 3], channels=2), None, 1),
            ("readme_chain", dict(seeds=[4], channels=2), dict(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=4), None),
            ("mode13", dict(seeds=[5], mode=13, freq_off=1500), None, None),
            ("mode9_awgn", dict(seeds=[6], mode=9, channels=2), dict(awgn_db=-24.0, seed=6), None),
            ("rate48k_mode10", dict(seeds=[7], rate=48000, mode=10, freq_off=3000), None, None),
            ("awgn_fail", dict(seeds=[8], channels=2), dict(awgn_db=-12.0, seed=8), None)]


def golden_stimulus(oracle, kw, imp):
    kw = dict(kw)
    pls = np.stack([oracle.make_payload(7000 + s) for s in kw.pop("seeds")])
    pcm = oracle.encode(pls, imp=oracle.impair(**imp) if imp else None, **kw)
    return pls, pcm, kw.get("rate", 8000), kw.get("channels", 1)


def test_oracle_matches_reference_tu_golden(oracle, tmp_path):
    g = json.load(open(os.path.join(HERE, "golden", "reference_tu.json")))
    for name, kw, imp, skip in golden_cases(oracle):
        pls, pcm, rate, ch = golden_stimulus(oracle, kw, imp)
        assert hashlib.sha256(np.ascontiguousarray(pcm, "<i2").tobytes()).hexdigest() == g[name]["pcm_sha256"], name
        wav = tmp_path / "g.wav"
        write_wav(wav, pcm, rate, ch)
        o = subprocess.run([os.path.join(ORA, "decode_ref"), str(tmp_path / "o.dat"), str(wav)] + ([str(skip)] if skip is not None else []), capture_output=True)
        assert o.returncode == 0
        assert hashlib.sha256((tmp_path / "o.dat").read_bytes()).hexdigest() == g[name]["payload_sha256"], name
        assert _stderr_lines(o.stderr) == g[name]["stderr"], name

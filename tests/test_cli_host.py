"""The product's `decode` host driver (modem_b200/csrc/host/decode_main.cc) on the CPU: linked against tests/mock_ofdmrx.cc — a
stand-in for libofdmrx.so's C-ABI built on the oracle — and diffed against the reference's own main() (oracle/_ref/decode, the
reference decode.cc over oracle/shim/).  What is checked is the HOST side of the drop-in: argv rules, WAV parsing (8/16/24-bit,
1/2 channels), the SKIP walk, every stderr line, exit codes, the 5380 output bytes.  The device side of the same binary is
covered by tests/test_gpu_parity.py::test_decode_cli_matches_reference_contract."""
import os
import subprocess

import numpy as np
import pytest

import test_reference_tu as T

ROOT = T.ROOT
needs_reference = T.needs_reference


@pytest.fixture(scope="module")
def cli(oracle):
    from modem_b200 import build as B
    d = os.path.join(B.OBJ, "mock")
    os.makedirs(d, exist_ok=True)
    lib, exe = os.path.join(d, "libofdmrx.so"), os.path.join(d, "decode")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", inc, os.path.join(ROOT, "tests", "mock_ofdmrx.cc"), "-o", lib], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", inc, os.path.join(B.CSRC, "host", "decode_main.cc"), "-o", exe, "-L", d, "-lofdmrx", "-Wl,-rpath," + d], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", inc, os.path.join(B.CSRC, "host", "encode_main.cc"), "-o", os.path.join(d, "encode"), "-L", d, "-lofdmrx", "-Wl,-rpath," + d], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    return exe


def run_both(cli, tmp_path, wav, skip=None):
    args = [str(wav)] + ([str(skip)] if skip is not None else [])
    p = subprocess.run([cli, str(tmp_path / "p.dat")] + args, capture_output=True)
    r = subprocess.run([os.path.join(T.REF, "decode"), str(tmp_path / "r.dat")] + args, capture_output=True)
    assert p.returncode == r.returncode == 0, (p.stderr, r.stderr)
    assert p.stderr.decode().splitlines() == r.stderr.decode().splitlines()      # every line, Es/N0 and sfo/cfo estimates included
    pd, rd = (tmp_path / "p.dat").read_bytes(), (tmp_path / "r.dat").read_bytes()
    assert len(pd) == len(rd) == 5380
    if b"bit flips:" in r.stderr:      # a failed reference decode writes an uninitialised buffer (decode.cc:588)
        assert pd == rd
    return pd, r.stderr


@needs_reference
def test_quick_start_skip_walk_and_failures(cli, oracle, tmp_path):
    pls = np.stack([oracle.make_payload(600 + i) for i in range(3)])
    three = oracle.encode(pls)
    damaged = three.copy()
    damaged[8000 + 2 * 1440:8000 + 3 * 1440] = np.random.default_rng(3).integers(-3000, 3000, 1440)   # first frame's metadata symbol
    cases = [(oracle.encode(pls[0]), 1, None, pls[0]), (three, 1, 0, pls[0]), (three, 1, 2, pls[2]), (three, 1, 3, None), (three, 1, 40, None),
             (damaged, 1, None, None), (damaged, 1, 1, pls[1]), (damaged, 1, 2, pls[2]),
             (oracle.encode(pls[1], channels=2, imp=oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=5)), 2, None, pls[1]),
             (oracle.encode(pls[0], channels=2, imp=oracle.impair(awgn_db=-12.0, seed=9)), 2, None, None),
             (oracle.encode(pls[0])[:40000], 1, None, None), (np.zeros(30000, np.int16), 1, None, None)]
    for pcm, ch, skip, want in cases:
        wav = tmp_path / "c.wav"
        T.write_wav(wav, pcm, 8000, ch)
        out, err = run_both(cli, tmp_path, wav, skip)
        if want is not None:
            assert out == want.tobytes()


@needs_reference
@pytest.mark.parametrize("rate,mode,bits", [(8000, 13, 16), (16000, 9, 16), (48000, 10, 16), (8000, 6, 8), (8000, 7, 24), (8000, 6, 32)])
def test_modes_rates_and_sample_widths(cli, oracle, tmp_path, rate, mode, bits):
    """8- and 24-bit files go through the reference's own encoder (Makefile:14 uses 8 bits).  The driver hands 16-bit samples to
    the library as they are and every other depth as floats scaled like DSP::ReadWAV<float> (v / (2^(bits-1) - 1), decode.cc:576),
    so payload AND every stderr float equal the reference's own main() at all depths."""
    pl = oracle.make_payload(rate + mode + bits)
    (tmp_path / "in.dat").write_bytes(pl.tobytes())
    wav = tmp_path / "e.wav"
    subprocess.run([os.path.join(T.REF, "encode"), str(wav), str(rate), str(bits), "1", "2000", str(mode), "CALLSIGN", str(tmp_path / "in.dat")], check=True, capture_output=True)
    out, err = run_both(cli, tmp_path, wav)
    assert out == pl.tobytes() and b"bit flips:" in err and ("oper mode: %d" % mode).encode() in err


@needs_reference
def test_usage_and_format_errors(cli, tmp_path):
    ref = os.path.join(T.REF, "decode")
    for args in ([], ["a"], ["a", "b", "1", "2"]):
        p, r = subprocess.run([cli] + args, capture_output=True), subprocess.run([ref] + args, capture_output=True)
        assert p.returncode == r.returncode == 1 and b"usage:" in p.stderr and b"OUTPUT INPUT [SKIP]" in r.stderr
    T.write_wav(tmp_path / "r.wav", np.zeros(1000, np.int16), 22050, 1)
    T.write_wav(tmp_path / "c.wav", np.zeros(3000, np.int16), 8000, 3)
    for wav in ("r.wav", "c.wav"):
        p = subprocess.run([cli, str(tmp_path / "x"), str(tmp_path / wav)], capture_output=True)
        r = subprocess.run([ref, str(tmp_path / "y"), str(tmp_path / wav)], capture_output=True)
        assert p.returncode == r.returncode == 1 and p.stderr == r.stderr
    p = subprocess.run([cli, str(tmp_path / "x"), str(tmp_path / "missing.wav")], capture_output=True)
    assert p.returncode == 1


def test_batch_extension(cli, oracle, tmp_path):
    """--batch[=STRIDE]: N back-to-back windows in one file -> N x 5380 bytes (no reference counterpart)"""
    pcm, ns, sent = oracle.encode_batch(3, seed0=50)
    T.write_wav(tmp_path / "b.wav", pcm.reshape(-1), 8000, 1)
    p = subprocess.run([cli, "--batch", str(tmp_path / "b.dat"), str(tmp_path / "b.wav")], capture_output=True)
    assert p.returncode == 0 and (tmp_path / "b.dat").read_bytes() == sent.tobytes()
    assert p.stderr.count(b"bit flips: 0") == 3 and b"window 2:" in p.stderr


@needs_reference
@pytest.mark.parametrize("args", [["8000", "16", "1", "2000", "6", "CALLSIGN"], ["8000", "8", "1", "2000", "6", "ANONYMOUS"], ["8000", "16", "2", "-450", "9", "dl1abc"],
                                  ["16000", "24", "2", "3000", "10", "A"], ["48000", "32", "1", "1600", "12", "Q 1"]])
def test_encode_driver_writes_the_reference_bytes(cli, oracle, tmp_path, args):
    """modem_b200/csrc/host/encode_main.cc over the mock (the oracle's float stream behind include/ofdmtx.h): header, sample
    widths, channel selection and quantisation of the WAV writer against the reference's own encode main()"""
    names = []
    for i in range(2):
        (tmp_path / ("in%d.dat" % i)).write_bytes(oracle.make_payload(70 + i).tobytes())
        names.append(str(tmp_path / ("in%d.dat" % i)))
    enc = os.path.join(os.path.dirname(cli), "encode")
    p = subprocess.run([enc, str(tmp_path / "p.wav")] + args + names, capture_output=True)
    r = subprocess.run([os.path.join(T.REF, "encode"), str(tmp_path / "r.wav")] + args + names, capture_output=True)
    assert p.returncode == r.returncode == 0, (p.stderr, r.stderr)
    pw, rw = (tmp_path / "p.wav").read_bytes(), (tmp_path / "r.wav").read_bytes()
    assert len(pw) == len(rw) and pw[:44] == rw[:44]
    if args[1] != "32":
        assert pw == rw
    else:   # 2^31 - 1 is not a float: the last bits of a 32-bit sample depend on where the product is rounded
        d = np.abs(np.frombuffer(pw[44:], "<i4").astype(np.int64) - np.frombuffer(rw[44:], "<i4").astype(np.int64))
        assert d.max() <= 256


@needs_reference
def test_encode_driver_argument_errors(cli, tmp_path):
    (tmp_path / "in.dat").write_bytes(bytes(5380))
    enc = os.path.join(os.path.dirname(cli), "encode")
    for args in (["8000", "16", "1", "2000", "5", "CALLSIGN"], ["8000", "16", "1", "2000", "6", "call-sign"], ["8000", "16", "1", "2000", "6", "TENLETTERS"],
                 ["8000", "16", "1", "1300", "6", "CALLSIGN"], ["8000", "16", "2", "2700", "6", "CALLSIGN"], ["8000", "16", "1", "2025", "6", "CALLSIGN"],
                 ["22050", "16", "1", "2000", "6", "CALLSIGN"]):
        p = subprocess.run([enc, str(tmp_path / "x.wav")] + args + [str(tmp_path / "in.dat")], capture_output=True)
        r = subprocess.run([os.path.join(T.REF, "encode"), str(tmp_path / "y.wav")] + args + [str(tmp_path / "in.dat")], capture_output=True)
        assert p.returncode == r.returncode == 1 and p.stderr == r.stderr, (args, p.stderr, r.stderr)
    p = subprocess.run([enc], capture_output=True)
    assert p.returncode == 1 and b"usage:" in p.stderr


def test_python_mirror_plumbing_through_the_mock(cli, oracle, tmp_path, monkeypatch):
    """modem_b200's ctypes layer (argument order and types of every call it makes) against the mock library: Transmitter.encode,
    Receiver.decode and decode_wav return what the oracle returns.  The product itself never loads anything but libofdmrx.so."""
    import modem_b200 as M
    monkeypatch.setattr(M, "LIB_PATH", os.path.join(os.path.dirname(cli), "libofdmrx.so"))
    monkeypatch.setattr(M, "_lib", None)
    try:
        pls = np.stack([oracle.make_payload(880 + i) for i in range(2)])
        tx = M.Transmitter(max_windows=2)
        assert tx.window_samples(6) == 95200 and tx.window_samples(13) == oracle.frame_samples(13)
        for ch in (1, 2):
            pcm, ns = tx.encode(pls, channels=ch)
            assert (ns == 95200).all()
            for i in range(2):
                assert (pcm[i].reshape(-1) == oracle.encode(pls[i], channels=ch).reshape(-1)).all()
        iq, ns = tx.encode(pls[:1], fmt=M.FMT_F32_IQ)
        assert iq.dtype == np.complex64 and iq.shape == (1, 95200) and abs(np.abs(iq[0, 8000:87200]) ** 2).mean() > 0.05
        tx.close()
        rx = M.Receiver(max_frames=2)
        payload, st = rx.decode(pcm, channels=2)
        assert (payload == pls).all() and (st["status"] == 0).all() and (st["mode"] == 6).all() and M.call_sign(st[0]) == " CALLSIGN"
        rx.close()
        T.write_wav(tmp_path / "w.wav", oracle.encode(pls[1]), 8000, 1)
        data, s = M.decode_wav(str(tmp_path / "w.wav"))
        assert data == pls[1].tobytes() and s["flips"] == 0
        # the probes' window source (tools/_stimulus.py) with STIM=device: same payloads and impairment definitions as the CPU path
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import _stimulus
        kw = dict(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=21)
        monkeypatch.setenv("STIM", "cpu")
        c_pcm, c_ns, c_sent = _stimulus.windows(2, 300, channels=2, imp=kw)
        monkeypatch.setenv("STIM", "device")
        d_pcm, d_ns, d_sent = _stimulus.windows(2, 300, channels=2, imp=kw)
        assert (c_sent == d_sent).all() and (c_ns == d_ns).all() and (c_pcm == d_pcm).all()
    finally:
        M._lib = None   # later tests must bind the real library again


@needs_reference
def test_integration_binding_compiles_into_the_reference_main(cli, oracle, tmp_path):
    """INTEGRATION.md §2 for real: tests/integration/decode_cc_binding.inc spliced into a scratch copy of the reference's
    decode.cc (nothing of the reference is kept in the repository), compiled over oracle/shim/ and linked to the mock of the
    C-ABI — the patched binary returns the payload through ofdmrx_decode_batch and skips the reference's own Decoder."""
    src = open(os.path.join(T.REFSRC, "decode.cc")).read().split("\n")
    inc = open(os.path.join(ROOT, "tests", "integration", "decode_cc_binding.inc")).read()
    out, state = [], 0
    for ln in src:
        if state == 0 and ln.startswith("#include \"polar_list_decoder.hh\""):
            out += [ln, "#include <vector>", "#include \"ofdmrx.h\""]
            continue
        if ln.strip() == "switch (input_file.rate()) {":
            out += [inc, "\tif (!ofdmrx_done)"]          # the reference's CPU path stays as the fallback
            state = 1
        if ln.strip() == "CODE::Xorshift32 scrambler;":
            out.append("\tif (!ofdmrx_done) {")          # the library already de-scrambled
            state = 2
        out.append(ln)
        if state == 2 and "output_data[i] ^= scrambler();" in ln:
            out.append("\t}")
            state = 3
    assert state == 3
    (tmp_path / "decode_patched.cc").write_text("\n".join(out))
    d = os.path.dirname(cli)
    exe = str(tmp_path / "decode_patched")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-strict-aliasing", "-D__AVX2__=1", "-I", os.path.join(ROOT, "oracle", "shim"),
                    "-I", T.REFSRC, "-I", os.path.join(ROOT, "include"), str(tmp_path / "decode_patched.cc"), "-o", exe, "-L", d, "-lofdmrx", "-Wl,-rpath," + d], check=True)
    pls = np.stack([oracle.make_payload(990 + i) for i in range(2)])
    T.write_wav(tmp_path / "two.wav", oracle.encode(pls, channels=2, imp=oracle.impair(cfo_hz=12.5, awgn_db=-26.0, seed=2)), 8000, 2)
    for skip in (0, 1):
        r = subprocess.run([exe, str(tmp_path / "o.dat"), str(tmp_path / "two.wav"), str(skip)], capture_output=True)
        assert r.returncode == 0 and (tmp_path / "o.dat").read_bytes() == pls[skip].tobytes()
        assert b"demod" not in r.stderr          # the reference's own Decoder did not run

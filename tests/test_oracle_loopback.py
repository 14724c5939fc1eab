"""Loop-back and CLI-contract checks of the CPU oracle (README.md:4-50 recipes, Makefile:13-15 smoke run)."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(os.path.dirname(HERE), "oracle", "build")


def test_clean_loopback_mono_and_iq(oracle):
    pl = oracle.make_payload(11)
    for ch in (1, 2):
        pcm = oracle.encode(pl, channels=ch)
        assert pcm.shape[0] == 95200                                   # 1 s + 55 symbols + 1 s (encode.cc:288,311-313,423,441)
        st, out, tp = oracle.decode(pcm, channels=ch)
        assert st == 0 and (out == pl).all() and tp.flips == 0 and tp.mode == 6
        assert tp.call_sign.decode().strip() == "CALLSIGN"
        assert tp.shift == 160 and abs(tp.cfo_rad * 8000 / (2 * np.pi) - 2000) < 1e-2
        assert tp.sc_pos == (9611 if ch == 1 else 9600)                # mono: +11 samples of Hilbert group delay


def test_readme_impairment_chain(oracle):
    pl = oracle.make_payload(12)
    imp = oracle.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=2)
    st, out, tp = oracle.decode(oracle.encode(pl, channels=2, imp=imp), channels=2)
    assert st == 0 and (out == pl).all()
    assert abs(tp.cfo_rad * 8000 / (2 * np.pi) - 2234.567) < 1.0


def test_golden_regression(oracle):
    g = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    pcm = oracle.encode(g["payload"])
    assert (pcm[9500:9800] == g["pcm_head"]).all()
    st, out, tp = oracle.decode(pcm)
    assert st == 0 and tp.sc_pos == int(g["sc_pos"]) and (oracle.taps_np(tp, "soft")[:255] == g["soft"]).all()
    assert np.allclose(oracle.taps_np(tp, "precision"), g["precision"], rtol=1e-4)


def test_list_size_and_rate0_variants(oracle):
    pl = oracle.make_payload(13)
    pcm = oracle.encode(pl, channels=2, imp=oracle.impair(awgn_db=-22, seed=4))
    st8, out8, t8 = oracle.decode(pcm, channels=2)
    st1, out1, t1 = oracle.decode(pcm, channels=2, r0_max=1)           # leaf-by-leaf rate-0 accumulation
    st4, out4, t4 = oracle.decode(pcm, channels=2, list_size=4)        # non-AVX2 reference build
    assert st8 == st1 and (out8 == out1).all()
    assert np.allclose(oracle.taps_np(t8, "metrics"), oracle.taps_np(t1, "metrics"), rtol=1e-5)
    assert st4 in (0, 6) and (st4 != 0 or (out4 == pl).all())


def test_multiframe_skip_semantics(oracle):
    pls = np.stack([oracle.make_payload(20 + i) for i in range(3)])
    pcm = oracle.encode(pls)                                           # one WAV, three frames (encode.cc:289)
    for skip in range(3):
        st, out, tp = oracle.decode(pcm, skip=skip)
        assert st == 0 and (out == pls[skip]).all() and tp.detections >= skip + 1
    st, out, tp = oracle.decode(pcm, skip=3)                            # runs out of detections
    assert st != 0


def test_failure_paths(oracle):
    st, out, tp = oracle.decode(np.zeros(20000, np.int16))
    assert st == 1                                                      # no sync
    pl = oracle.make_payload(30)
    pcm = oracle.encode(pl).copy()
    st, out, tp = oracle.decode(pcm[:50000])                            # truncated mid-frame: header ok, payload garbage
    assert st in (6, 0) and not (st == 0 and (out == pl).all())
    rng = np.random.default_rng(0)
    st, out, tp = oracle.decode((rng.standard_normal(60000) * 3000).astype(np.int16))
    assert st != 0


def test_other_modes_loopback(oracle):
    pl = oracle.make_payload(31)
    for mode, off in ((7, 2000), (9, 1500), (10, 2000), (13, 1000)):
        st, out, tp = oracle.decode(oracle.encode(pl, mode=mode, freq_off=off))
        assert st == 0 and (out == pl).all() and tp.mode == mode


def test_cli_contract(oracle, tmp_path):
    enc, dec = os.path.join(BUILD, "encode_ref"), os.path.join(BUILD, "decode_ref")
    r = subprocess.run([dec], capture_output=True)
    assert r.returncode == 1 and b"usage:" in r.stderr                  # decode.cc:561-564
    r = subprocess.run([enc, "x.wav", "8000", "16", "1", "2000", "5", "CALLSIGN", "/dev/null"], capture_output=True)
    assert r.returncode == 1 and b"Unsupported operation mode." in r.stderr
    r = subprocess.run([enc, "x.wav", "8000", "16", "1", "2025", "6", "CALLSIGN", "/dev/null"], capture_output=True)
    assert r.returncode == 1 and b"divisible by 50" in r.stderr
    data = os.urandom(5380)
    (tmp_path / "in.dat").write_bytes(data)
    wav, out = str(tmp_path / "e.wav"), str(tmp_path / "out.dat")
    assert subprocess.run([enc, wav, "8000", "16", "1", "2000", "6", "CALLSIGN", str(tmp_path / "in.dat")]).returncode == 0
    assert os.path.getsize(wav) == 190444
    r = subprocess.run([dec, out, wav], capture_output=True)
    assert r.returncode == 0 and open(out, "rb").read() == data
    for line in (b"symbol pos:", b"coarse cfo:", b"oper mode: 6", b"call sign:  CALLSIGN", b"Es/N0 (dB):", b"bit flips: 0"):
        assert line in r.stderr
    # Makefile:13-15 smoke run: 8-bit mono
    assert subprocess.run([enc, wav, "8000", "8", "1", "2000", "6", "ANONYMOUS", str(tmp_path / "in.dat")]).returncode == 0
    r = subprocess.run([dec, out, wav], capture_output=True)
    assert r.returncode == 0 and len(open(out, "rb").read()) == 5380
    # a failed decode still writes 5380 bytes and exits 0 (decode.cc:608-619)
    import modem_b200
    silent = str(tmp_path / "silent.wav")
    subprocess.run([enc, silent, "8000", "16", "1", "2000", "6", "CALLSIGN", str(tmp_path / "in.dat")])
    raw = bytearray(open(silent, "rb").read())
    raw[44:] = bytes(len(raw) - 44)
    open(silent, "wb").write(raw)
    r = subprocess.run([dec, out, silent], capture_output=True)
    assert r.returncode == 0 and len(open(out, "rb").read()) == 5380

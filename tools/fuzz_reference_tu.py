"""Developer probe (build container, needs /root/reference): random modes, rates, channel counts, carrier offsets, frame counts,
SKIP values and impairments through oracle/_ref/decode (the reference's own decode.cc over oracle/shim/) and the oracle's
decode_ref — payload bytes and the complete stderr must be identical.  `python tools/fuzz_reference_tu.py [cases]`."""
import os, sys, subprocess, numpy as np, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as O, test_reference_tu as T
rng = np.random.default_rng(2026)
bad = 0
with tempfile.TemporaryDirectory() as d:
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
        mode = int(rng.integers(6, 14)); rate = int(rng.choice([8000, 8000, 8000, 16000, 44100, 48000])); ch = int(rng.integers(1, 3))
        bw = {6: 2700, 7: 2500, 8: 2500, 9: 2250, 10: 3200, 11: 2400, 12: 2400, 13: 1600}[mode]
        lo = bw // 2 if ch == 1 else bw // 2 - rate // 2
        hi = rate // 2 - bw // 2
        off = int(rng.integers(-(-lo // 50), hi // 50 + 1)) * 50
        n_fr = int(rng.integers(1, 3)); skip = int(rng.integers(0, 3))
        kw = {}
        if ch == 2 and rng.random() < 0.7:
            kw = dict(multipath=bool(rng.integers(0, 2)), cfo_hz=float(rng.uniform(-40, 40)), sfo_ppm=float(rng.uniform(-100, 100)) if rate == 8000 else 0.0,
                      awgn_db=float(rng.uniform(-32, -13)), seed=int(rng.integers(1, 1000)))
        pls = np.stack([O.make_payload(int(rng.integers(0, 10 ** 6))) for _ in range(n_fr)])
        pcm = O.encode(pls, rate=rate, channels=ch, freq_off=off, mode=mode, imp=O.impair(**kw) if kw else None)
        wav = os.path.join(d, "f.wav"); T.write_wav(wav, pcm, rate, ch)
        r = subprocess.run([os.path.join(T.REF, "decode"), os.path.join(d, "r.dat"), wav, str(skip)], capture_output=True)
        o = subprocess.run([os.path.join(T.ORA, "decode_ref"), os.path.join(d, "o.dat"), wav, str(skip)], capture_output=True)
        same_err = r.stderr == o.stderr
        ok = b"bit flips" in r.stderr
        same_out = (not ok) or open(os.path.join(d, "r.dat"), "rb").read() == open(os.path.join(d, "o.dat"), "rb").read()
        print(it, mode, rate, ch, off, n_fr, skip, "ok" if ok else "fail", same_err, same_out, flush=True)
        bad += not (same_err and same_out)
print("mismatches:", bad)

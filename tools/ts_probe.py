"""Developer probe (gpurun): how many pair sweeps the Theil-Sen bracket search needs on real phase-error rows."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O

rx = M.Receiver(max_frames=64, keep_taps=True)
for name, imp, ch in (("clean mono", None, 1), ("awgn -30 + multipath/cfo/sfo", O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=3), 2),
                      ("awgn -20", O.impair(awgn_db=-20, seed=4), 2), ("awgn -14.5", O.impair(awgn_db=-14.5, seed=5), 2)):
    pcm, ns, sent = O.encode_batch(32, seed0=900, channels=ch, imp=imp) if imp is not None else O.encode_batch(32, seed0=900, channels=ch)
    payload, st = rx.decode(pcm, channels=ch)
    ok = np.nonzero(st["status"] != 99)[0]
    y = np.concatenate([rx.taps(M.TAP_PHASE, int(f), 1)[0] for f in ok if st["status"][f] in (0, 6)])
    slope, yint = rx.theil_sen(y)
    sw = rx.last_sweeps
    print("%-32s rows %5d  sweeps histogram %s  sigma(y) median %.4f" % (name, len(sw), np.bincount(np.minimum(sw, 20)).tolist(), np.median(y.std(axis=1))), flush=True)

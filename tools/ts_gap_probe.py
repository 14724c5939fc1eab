"""Developer probe (gpurun): how far the exact Theil-Sen slope lies from the OLS pilot, in units of the first bracket."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O
rx = M.Receiver(max_frames=64, keep_taps=True)
x = np.arange(432) - 216 + 0.5
for name, imp, ch in (("clean mono", None, 1), ("readme chain", O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=3), 2),
                      ("multipath + awgn -20", O.impair(multipath=True, awgn_db=-20, seed=4), 2), ("awgn -16", O.impair(awgn_db=-16, seed=5), 2)):
    pcm, ns, sent = O.encode_batch(16, seed0=900, channels=ch, imp=imp) if imp is not None else O.encode_batch(16, seed0=900, channels=ch)
    payload, st = rx.decode(pcm, channels=ch)
    y = np.concatenate([rx.taps(M.TAP_PHASE, int(f), 1)[0] for f in range(16) if st["status"][f] in (0, 6)]).astype(np.float64)
    slope, yint = rx.theil_sen(y.astype(np.float32))
    sw = rx.last_sweeps
    c0 = (y * x).sum(1) / (432 * (432 ** 2 - 1) / 12)
    res = y - y.mean(1, keepdims=True) - c0[:, None] * x
    s_res = np.sqrt((res ** 2).sum(1) / 430)
    s_d = np.sqrt((np.diff(y, axis=1) ** 2).sum(1) / (2 * 431))
    gap_res = np.abs(slope - c0) / (1.35e-4 * s_res)
    gap_d = np.abs(slope - c0) / (1.35e-4 * s_d)
    print("%-22s rows %4d sweeps %s | s_res/s_d median %.2f p90 %.2f | gap/bracket(s_res): median %.1f p90 %.1f max %.1f | gap/bracket(s_d): median %.1f p90 %.1f max %.1f"
          % (name, len(sw), np.bincount(np.minimum(sw, 9)).tolist(), np.median(s_res / s_d), np.percentile(s_res / s_d, 90),
             np.median(gap_res), np.percentile(gap_res, 90), gap_res.max(), np.median(gap_d), np.percentile(gap_d, 90), gap_d.max()), flush=True)

"""Developer probe (gpurun): per-row differences of the Theil-Sen line between the B200 path and the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O
rx = M.Receiver(max_frames=8, keep_taps=True)
for seed in (77, 31):
    imp = O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=seed)
    pcm, ns, sent = O.encode_batch(4, seed0=2000, channels=2, imp=imp)
    payload, st = rx.decode(pcm, channels=2)
    for i in range(2):
        ost, opay, tp = O.decode(pcm[i], channels=2)
        ts = rx.taps(M.TAP_TS, i, 1, 6)[0]
        dy = np.abs(ts[:, 1] - O.taps_np(tp, "yint")); ds = np.abs(ts[:, 0] - O.taps_np(tp, "slope")) * 216
        y = rx.taps(M.TAP_PHASE, i, 1, 6)[0]
        print("seed", seed, "win", i, "yint diff: median %.2e max %.2e | slope*216 diff: median %.2e max %.2e | rows>1e-4: %d | sigma(y) %.3f" % (np.median(dy), dy.max(), np.median(ds), ds.max(), int(((dy > 1e-4) | (ds > 1e-4)).sum()), y.std(axis=1).mean()))
        print("   dy sorted head", np.sort(dy)[::-1][:6], "ds", np.sort(ds)[::-1][:6])

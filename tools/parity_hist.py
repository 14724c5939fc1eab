"""Developer probe (gpurun): distribution of the stage-tap differences between the B200 path and the CPU oracle over many
windows — the evidence behind the tolerances in tests/test_gpu_parity.py / test_gpu_frontend.py.  Writes a markdown table."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O

N = int(os.environ.get("N", "48"))
rx = M.Receiver(max_frames=N, keep_taps=True)
rows = []
for name, imp, ch in (("clean mono", None, 1),
                      ("README chain (multipath, CFO 234.567 Hz, SFO 147 ppm, AWGN -30)", dict(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=31), 2),
                      ("AWGN -20", dict(awgn_db=-20, seed=32), 2), ("AWGN -15", dict(awgn_db=-15, seed=33), 2)):
    pcm, ns, sent = O.encode_batch(N, seed0=5000, channels=ch, imp=O.impair(**imp) if imp else None)
    payload, st = rx.decode(pcm, channels=ch)
    e = {k: [] for k in ("iq", "timing", "cons_raw", "cons", "slope_rel", "yint", "precision_rel", "llr_norm", "llr_p999")}
    ok = 0
    for i in range(N):
        ost, opay, tp = O.decode(pcm[i], channels=ch)
        oiq, otm = O.front_taps(pcm[i], channels=ch)
        n1 = oiq.size
        e["iq"].append(np.abs(rx.taps(M.TAP_IQ, i, 1)[0].view(np.complex64)[:n1] - oiq).max())
        e["timing"].append(np.abs(rx.taps(M.TAP_TIMING, i, 1)[0][:n1] - otm).max())
        if ost not in (0, 6) or st["status"][i] != ost:
            continue
        ok += 1
        e["cons_raw"].append(np.abs(rx.taps(M.TAP_CONS_RAW, i, 1, 6)[0] - O.taps_np(tp, "cons_raw")).max())
        oc = O.taps_np(tp, "cons")
        e["cons"].append((np.abs(rx.taps(M.TAP_CONS, i, 1, 6)[0] - oc) / np.maximum(1.0, np.abs(oc))).max())
        ts = rx.taps(M.TAP_TS, i, 1, 6)[0]
        osl = O.taps_np(tp, "slope")
        e["slope_rel"].append(np.abs(ts[:, 0] - osl).max() / max(np.abs(osl).max(), 1e-12))
        e["yint"].append(np.abs(ts[:, 1] - O.taps_np(tp, "yint")).max())
        e["precision_rel"].append((np.abs(ts[:, 2] - O.taps_np(tp, "precision")) / O.taps_np(tp, "precision")).max())
        ollr = O.taps_np(tp, "llr")
        dl = np.abs(rx.taps(M.TAP_LLR, i, 1)[0] - ollr) / np.abs(ollr[:64512]).mean()
        e["llr_norm"].append(dl.max())
        e["llr_p999"].append(np.quantile(dl, 0.999))
    rows.append((name, ok, {k: (np.median(v), np.max(v)) if len(v) else (float("nan"), float("nan")) for k, v in e.items()}))
out = ["# Stage-tap differences B200 vs CPU oracle, %d windows per channel condition (max over each window; median / max over windows)" % N, "",
       "| condition | windows compared | iq | timing | cons_raw | cons (rel.) | slope (rel. to max) | yint | precision (rel.) | LLR / mean abs LLR (max) | LLR (99.9th percentile) |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for name, ok, d in rows:
    out.append("| %s | %d | " % (name, ok) + " | ".join("%.2e / %.2e" % d[k] for k in ("iq", "timing", "cons_raw", "cons", "slope_rel", "yint", "precision_rel", "llr_norm", "llr_p999")) + " |")
txt = "\n".join(out) + "\n"
print(txt)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "parity_hist.md"), "w").write(txt)

"""BASELINE configs[3]: AWGN sweep, 1k windows per point, payload BER / FER of the B200 path vs the CPU oracle on the
SAME noisy windows (gpurun).  Writes profiles/ber_sweep_<tag>.json and .md.

AWGN level L dB = complex noise of total variance 10^(L/10) added to the analytic signal (signal power ~ -9.1 dBFS),
see oracle/ref_modem.hh apply_impairments (the absent `disorders/awgn` re-specified)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M  # noqa: E402
import oracle_lib as O  # noqa: E402
from _stimulus import windows  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    n = int(os.environ.get("N_PER_POINT", "1000"))
    mode, rate = int(os.environ.get("MODE", "6")), int(os.environ.get("RATE", "8000"))
    lo, hi, step = float(os.environ.get("DB_LO", "-40")), float(os.environ.get("DB_HI", "-10")), float(os.environ.get("DB_STEP", "1"))
    levels = list(np.arange(lo, hi + 1e-9, step))
    if os.environ.get("FINE", "1") != "0":
        fine = [x for x in np.arange(-17.0, -12.99, 0.25) if x not in levels]   # resolve the waterfall (mode 6)
        levels = sorted(set(levels) | set(fine))
    stride = O.frame_samples(mode, rate)
    rx = M.Receiver(max_frames=n, max_samples=stride, rate=rate)
    cores = os.cpu_count() or 1
    rows = []
    for db in levels:
        t0 = time.time()
        pcm, ns, sent = windows(n, int(1e6 + 1000 * (db + 100)), channels=2, rate=rate, mode=mode, stride=stride,
                                imp=dict(awgn_db=float(db), seed=int(7e5 + 100 * (db + 100))))   # STIM=device: generated on the GPU
        gp, gs = rx.decode(pcm, channels=2)
        ost, op = O.decode_batch(pcm, channels=2, rate=rate, nthreads=cores)
        bits = n * 43040
        g_err = int(np.unpackbits(gp ^ sent, axis=1).sum())
        o_err = int(np.unpackbits(op ^ sent, axis=1).sum())
        g_fail, o_fail = int((gs["status"] != 0).sum()), int((ost != 0).sum())
        both_ok = (gs["status"] == 0) & (ost == 0)
        mism_ok = int((gp[both_ok] != op[both_ok]).any(axis=1).sum())
        only_g, only_o = int(((gs["status"] == 0) & (ost != 0)).sum()), int(((gs["status"] != 0) & (ost == 0)).sum())
        rows.append({"awgn_db": float(db), "frames": n, "gpu_fer": g_fail / n, "cpu_fer": o_fail / n, "gpu_ber": g_err / bits, "cpu_ber": o_err / bits,
                     "payload_mismatch_when_both_decode": mism_ok, "decoded_only_by_gpu": only_g, "decoded_only_by_cpu": only_o,
                     "status_equal": int((gs["status"] == ost).sum()), "seconds": time.time() - t0})
        print(rows[-1], flush=True)
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    json.dump({"n_per_point": n, "stimulus": os.environ.get("STIM", "cpu"), "rows": rows}, open(os.path.join(out, "ber_sweep_%s.json" % tag), "w"), indent=1)
    with open(os.path.join(out, "ber_sweep_%s.md" % tag), "w") as f:
        f.write("# AWGN sweep (BASELINE configs[3]): B200 path vs CPU oracle on identical windows, mode %d at %d Hz, %d windows/point\n\n" % (mode, rate, n))
        f.write("| AWGN dB | GPU FER | CPU FER | GPU BER | CPU BER | both decode but differ | only GPU | only CPU |\n|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %.2f | %.4f | %.4f | %.3e | %.3e | %d | %d | %d |\n" % (r["awgn_db"], r["gpu_fer"], r["cpu_fer"], r["gpu_ber"], r["cpu_ber"],
                    r["payload_mismatch_when_both_decode"], r["decoded_only_by_gpu"], r["decoded_only_by_cpu"]))
    rx.close()


if __name__ == "__main__":
    main()

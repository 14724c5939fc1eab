"""Developer diagnostic for a GPU box (run through gpurun): stage-by-stage comparison against the CPU oracle
with verbose output.  Not part of the product or of the test-suite (tests/ holds the gated versions)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M  # noqa: E402
import oracle_lib as O  # noqa: E402


def section(name):
    print("\n==== " + name, flush=True)


def noisy_llr(seed, sigma):
    import ctypes as C
    rng = np.random.default_rng(seed)
    pl = O.make_payload(1000 + seed)
    code = np.zeros(64800, np.uint8)
    O.lib().ref_payload_to_code(pl.ctypes.data_as(C.c_void_p), 6, code.ctypes.data_as(C.c_void_p))
    x = 1.0 - 2.0 * code.astype(np.float32)
    y = x + sigma * rng.standard_normal(64800).astype(np.float32)
    llr = np.concatenate([(2 * y / max(sigma, 0.3) ** 2).astype(np.float32), np.full(736, 9000, np.float32)])
    return pl, llr


def check_polar(rx):
    section("polar SCL vs oracle")
    sig = [0.0, 0.5, 0.7, 0.75, 0.8, 0.9, 0.6, 0.78]
    pls, llrs = zip(*[noisy_llr(i, s) for i, s in enumerate(sig)])
    llr = np.stack(llrs)
    t = time.time()
    payload, st, xb = rx.polar_decode(llr, want_xbits=True)
    print("gpu polar_decode %d cw: %.3f s" % (len(sig), time.time() - t))
    ok = True
    for i, s in enumerate(sig):
        best, lanes, met, opay, flips = O.polar_decode(llr[i])
        gl = np.unpackbits(xb[i].view(np.uint8), bitorder="little").reshape(8, 65536)
        lanes_eq = bool((gl == lanes).all())
        met_eq = bool((st["metrics"][i] == met).all())
        pay_eq = bool((payload[i] == opay).all())
        print("sigma %.2f: oracle best %d gpu best %d status %d flips %d/%d  lanes_eq %s metrics_eq %s payload_eq %s  m0 %.4f/%.4f"
              % (s, best, st["best_lane"][i], st["status"][i], st["flips"][i], flips, lanes_eq, met_eq, pay_eq,
                 st["metrics"][i][0], met[0]))
        ok &= lanes_eq and met_eq and pay_eq and best == st["best_lane"][i]
    print("POLAR", "PASS" if ok else "FAIL")
    return ok


def check_pipeline(rx, channels, imp, name, n=4):
    section("pipeline " + name)
    pcm, ns, pay = O.encode_batch(n, seed0=7, channels=channels, imp=imp)
    t = time.time()
    payload, st = rx.decode(pcm, channels=channels)
    print("gpu decode %d windows: %.3f s, launches %d" % (n, time.time() - t, rx.last_launches))
    allok = True
    for i in range(n):
        ost, opay, tp = O.decode(pcm[i], channels=channels)
        s = st[i]
        print("frame %d: status gpu %d oracle %d | sc_pos %d/%d t_fire %d/%d index_max %d/%d shift %d/%d pos_err %d/%d cfo %.6f/%.6f tmax %.3f/%.3f"
              % (i, s["status"], ost, s["sc_pos"], tp.sc_pos, s["t_fire"], tp.t_fire, s["index_max"], tp.index_max, s["shift"], tp.shift,
                 s["pos_err"], tp.pos_err, s["cfo_rad"], tp.cfo_rad, s["timing_max"], tp.timing_max))
        soft = rx.taps(M.TAP_SOFT, i, 1)[0][:255]
        osoft = O.taps_np(tp, "soft")[:255]
        print("   soft maxdiff %d, md %x/%x mode %d" % (np.abs(soft.astype(int) - osoft.astype(int)).max(),
              (int(s["md_hi"]) << 32) | int(s["md_lo"]), tp.md, s["mode"]))
        if ost == 0 or ost == 6:
            cr = rx.taps(M.TAP_CONS_RAW, i, 1)[0]
            co = rx.taps(M.TAP_CONS, i, 1)[0]
            ts = rx.taps(M.TAP_TS, i, 1)[0]
            llr = rx.taps(M.TAP_LLR, i, 1)[0]
            ocr, oco = O.taps_np(tp, "cons_raw"), O.taps_np(tp, "cons")
            print("   cons_raw maxabs %.3e  cons maxabs %.3e" % (np.abs(cr - ocr).max(), np.abs(co - oco).max()))
            print("   slope maxabs %.3e (rel %.3e) yint maxabs %.3e precision maxrel %.3e"
                  % (np.abs(ts[:, 0] - O.taps_np(tp, "slope")).max(),
                     np.abs(ts[:, 0] - O.taps_np(tp, "slope")).max() / (np.abs(O.taps_np(tp, "slope")).max() + 1e-30),
                     np.abs(ts[:, 1] - O.taps_np(tp, "yint")).max(),
                     (np.abs(ts[:, 2] - O.taps_np(tp, "precision")) / O.taps_np(tp, "precision")).max()))
            ollr = O.taps_np(tp, "llr")
            scale = np.abs(ollr[:64800]).mean()
            print("   llr max|diff|/mean|llr| %.3e  sign mismatches %d" % (np.abs(llr - ollr).max() / scale,
                  int(((llr < 0) != (ollr < 0)).sum())))
            print("   metrics gpu %s\n           ora %s  flips %d/%d best %d/%d" % (s["metrics"][:4], O.taps_np(tp, "metrics")[:4],
                  s["flips"], tp.flips, s["best_lane"], tp.best_lane))
        pe = bool((payload[i] == opay).all())
        pin = bool((payload[i] == pay[i]).all())
        print("   payload == oracle %s, == sent %s" % (pe, pin))
        allok &= pe and s["status"] == ost
    print("PIPELINE", name, "PASS" if allok else "FAIL")
    return allok


def check_speed(rx, n):
    section("speed, %d windows (clean mono, %d unique)" % (n, 8))
    pcm8, ns, pay8 = O.encode_batch(8, seed0=100)
    pcm = np.tile(pcm8, (n // 8, 1))
    for rep in range(2):
        t = time.time()
        payload, st = rx.decode(pcm)
        dt = time.time() - t
        good = int((st["status"] == 0).sum())
        match = int((payload == np.tile(pay8, (n // 8, 1))).all(axis=1).sum())
        print("rep %d: %.3f s -> %.1f frames/s, ok %d, payload match %d" % (rep, dt, n / dt, good, match))


def main():
    print(M.load().ofdmrx_version().decode())
    nspeed = int(os.environ.get("N_SPEED", "512"))
    rx = M.Receiver(max_frames=max(nspeed, 16), keep_taps=True)
    import ctypes as C
    fr = rx.table(0)
    ofr = np.zeros(2048, np.uint32)
    O.lib().ref_frozen_table(0, ofr.ctypes.data_as(C.c_void_p))
    print("frozen table on device == oracle:", bool((fr == ofr).all()))
    results = {}
    for name, fn in [("polar", lambda: check_polar(rx)),
                     ("clean_mono", lambda: check_pipeline(rx, 1, None, "clean mono")),
                     ("clean_iq", lambda: check_pipeline(rx, 2, None, "clean iq")),
                     ("impaired_iq", lambda: check_pipeline(rx, 2, O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=5), "README chain iq")),
                     ("speed", lambda: check_speed(rx, nspeed))]:
        try:
            results[name] = fn()
        except Exception:
            traceback.print_exc()
            results[name] = "EXC"
    print("\nSUMMARY", results)


if __name__ == "__main__":
    main()

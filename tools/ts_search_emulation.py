"""Developer tool (CPU, numpy): the bracket search of k_theil_sen (modem_b200/csrc/demod.cu: ts_slope) restated on exact sorted
quotients, with the per-lane sub-queues modelled pair by pair — how many sweeps a row needs, and why.  It mirrors the search's
control flow and constants, not its arithmetic (the pilot runs in float64), so it answers "does the search converge on this
row", not "what does the kernel return".

    python tools/ts_search_emulation.py rows.npy [huber_steps] [early_stop] [old|new]
rows.npy: float32 [rows, carriers] phase errors (tap PHASE; tools/ts_dump_fallback_rows.py writes the windows that needed the
bisection fallback).  `old` = the overflow zoom before the fix of round 2 (hw = width * cap / (6 nq))."""
import sys
import numpy as np

F32 = np.float32
K_CAP, LANE_CAP, CAND_CAP = 2048, 64, 1280   # kTsCap, kTsLaneCap, kTsCandCap


def pilot(y, steps=4, early_stop=0.0, half_k=0.9e-4):
    """least squares start + Huber M-estimate (k = 1.345, proposal-2 scale) -> (slope, scale, steps used)"""
    n = y.size
    x = np.arange(n) - n // 2 + 0.5
    y = y.astype(np.float64)
    c0, icpt = (x * y).sum() / (n * (n * n - 1.0) / 12.0), y.mean()
    e = y - icpt - c0 * x
    srob = np.sqrt((e * e).sum() / (n - 2))
    kh, beta, used = 1.345, 0.71016, 0
    for _ in range(steps):
        e = y - icpt - c0 * x
        srob = np.sqrt(np.minimum(e * e, (kh * srob) ** 2).sum() / (n * beta))
        ae = np.abs(e)
        w = np.where(ae <= kh * srob, 1.0, kh * srob / np.maximum(ae, 1e-300))
        sw, swx, swy, swxx, swxy = w.sum(), (w * x).sum(), (w * y).sum(), (w * x * x).sum(), (w * x * y).sum()
        den = sw * swxx - swx * swx
        used += 1
        if not den > 0 or not srob > 0:
            break
        c1 = (sw * swxy - swx * swy) / den
        moved, c0 = abs(c1 - c0), c1
        icpt = (swy - c0 * swx) / sw
        if moved < early_stop * half_k * srob:
            break
    return c0, srob, used


def search(y, steps=4, early_stop=0.0, half_k=0.9e-4, new_zoom=True):
    """-> (sweeps, +100 if the bisection fallback would run; trace of (blo - pilot, bhi - pilot, #below - rank, in bracket, fullest lane))"""
    n = y.size
    I, J = np.triu_indices(n, 1)
    q = ((y[J] - y[I]).astype(F32) / (J - I).astype(F32)).astype(F32)
    order = np.argsort(q, kind="stable")
    qs = q[order]
    rank, pairs = q.size // 2, q.size
    c0, srob, _ = pilot(y, steps, early_stop, half_k)
    half = max(half_k * srob * (432.0 / n) ** 1.5, abs(c0) * 4e-6, 1e-12)
    blo, bhi = c0 - half, c0 + half
    L, U, cL, cU = -np.inf, np.inf, 0, pairs
    yabs = float(np.abs(y).max())
    x = np.arange(n) - n // 2
    y64 = y.astype(np.float64)
    trace = []
    for attempt in range(16):
        width = bhi - blo
        eps = 6e-7 * (yabs + 432.0 * max(abs(blo), abs(bhi))) + 1e-30
        u = y64 - blo * x
        d = u[J] - u[I]
        queued = (d >= -eps) & (d < width * (J - I) + eps)     # what the sweep puts into the lanes' sub-queues
        lanes = np.bincount(I[queued] % 32, minlength=32)
        nq, nqmax, cd = int(queued.sum()), int(lanes.max()), int((d < -eps).sum())
        cb = int(np.searchsorted(qs, F32(blo), "left"))
        nin = int(np.searchsorted(qs, F32(bhi), "left")) - cb
        trace.append((blo - c0, bhi - c0, cb - rank, nin, nqmax))
        if nqmax > LANE_CAP:
            kd = rank - cd
            if kd < 0:
                U, cU = blo, cd
            elif kd >= nq:
                L, cL = bhi, cd + nq
            else:
                centre = blo + width * ((kd + 0.5) / nq)
                hw = 0.5 * width * min(K_CAP / (3.0 * nq), 0.75 * LANE_CAP / nqmax) if new_zoom else width * (K_CAP / (6.0 * nq))
                blo, bhi = max(centre - hw, blo), min(centre + hw, bhi)
                if not blo < bhi:
                    break
                continue
        else:
            kk = rank - cb
            if 0 <= kk < nin:
                if nin <= CAND_CAP:
                    return attempt + 1, trace
                L, cL, U, cU = blo, cb, bhi, cb + nin
                centre, hw = blo + width * ((kk + 0.5) / nin), width * (CAND_CAP / (4.0 * nin))
                blo, bhi = max(centre - hw, L), min(centre + hw, U)
                if not blo < bhi:
                    break
                continue
            if kk < 0:
                U, cU = blo, cb
            else:
                L, cL = bhi, cb + nin
            if nin >= 64 and not (L > -np.inf and U < np.inf):
                per = width / nin
                centre = blo + (kk - 0.5) * per if kk < 0 else bhi + ((kk - nin) + 0.5) * per
                hw = 0.5 * width * min(1.5, (K_CAP // 3) / nin)
                blo, bhi = max(centre - hw, L), min(centre + hw, U)
                if not blo < bhi:
                    break
                continue
        if L > -np.inf and U < np.inf:
            span, dens = U - L, float(max(cU - cL, 1))
            centre = L + span * ((rank - cL + 0.5) / dens)
            hw = max(span * (K_CAP / (8.0 * dens)), 0.25 * half)
            blo, bhi = max(centre - hw, L), min(centre + hw, U)
        elif U < np.inf:
            bhi, blo = U, U - 2.0 * width
        else:
            blo, bhi = L, L + 2.0 * width
        if not blo < bhi:
            break
    return 100 + attempt + 1, trace


if __name__ == "__main__":
    rows = np.load(sys.argv[1])
    rows = rows["phase"].reshape(-1, rows["phase"].shape[-1]) if hasattr(rows, "files") else rows
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    early = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    new_zoom = (sys.argv[4] if len(sys.argv) > 4 else "new") != "old"
    lo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    hi = int(sys.argv[6]) if len(sys.argv) > 6 else rows.shape[0]
    hist = {}
    for r in range(lo, hi):
        a, tr = search(rows[r], steps, early, new_zoom=new_zoom)
        hist[a] = hist.get(a, 0) + 1
        if a >= 3:
            print("row", r, "sweeps", a)
            for t in tr:
                print("   bracket %+.3e .. %+.3e around the pilot, below - rank %d, inside %d, fullest sub-queue %d" % t)
    print(sorted(hist.items()))

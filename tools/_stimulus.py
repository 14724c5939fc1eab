"""Shared by the gpurun probes: n single-frame windows either from the CPU oracle's encoder (the default everything in
profiles/ up to round 1 was made with) or, with STIM=device, from the device-side generator (include/ofdmtx.h) — same
payloads (oracle_lib.make_payload(seed0 + i)), same impairment definitions, Philox instead of mt19937 noise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M  # noqa: E402
import oracle_lib as O  # noqa: E402

_tx = {}


def windows(n, seed0, channels=2, rate=8000, mode=6, stride=None, imp=None):
    """imp: dict(multipath=, cfo_hz=, sfo_ppm=, awgn_db=, seed=) or None -> (pcm int16 [n, stride*channels], n_samples, payloads)"""
    imp = imp or {}
    if os.environ.get("STIM", "cpu") != "device":
        return O.encode_batch(n, seed0=seed0, channels=channels, rate=rate, mode=mode, stride=stride,
                              imp=O.impair(**imp) if imp else None)
    key = (rate, n)
    if key not in _tx:
        _tx[key] = M.Transmitter(max_windows=min(n, 4096), rate=rate)
    sent = np.stack([O.make_payload(seed0 + i) for i in range(n)])
    pcm, ns = _tx[key].encode(sent, mode=mode, channels=channels, imp=M.impairments(**imp) if imp else None, stride=stride)
    return pcm, ns, sent

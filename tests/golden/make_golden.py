"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference for the polar table digest).

 * polar_tables.sha256 — digests of the two frozen-bit tables parsed out of /root/reference/polar_tables.hh
   (the only golden vector the reference itself holds for this path; the table text is NOT copied).
 * oracle_vectors.npz — small regression vectors produced by the CPU oracle (oracle/), clearly NOT reference output:
   the reference cannot be built here (aicodix/dsp + aicodix/code are absent), see DESIGN.md.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def reference_tables():
    txt = open("/root/reference/polar_tables.hh").read()
    out = {}
    for m in re.finditer(r"static const uint32_t (\w+)\[(\d+)\] = \{([^}]*)\}", txt):
        vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", m.group(3))]
        assert len(vals) == int(m.group(2))
        out[m.group(1)] = np.array(vals, np.uint32)
    return out


def main():
    tabs = reference_tables()
    dig = {k: {"sha256": hashlib.sha256(v.astype("<u4").tobytes()).hexdigest(), "words": int(v.size),
               "frozen": int(sum(bin(int(x)).count("1") for x in v))} for k, v in tabs.items()}
    json.dump(dig, open(os.path.join(HERE, "polar_tables.sha256"), "w"), indent=1, sort_keys=True)
    import oracle_lib as O
    pl = O.make_payload(42)
    pcm = O.encode(pl)
    st, out, tp = O.decode(pcm)
    assert st == 0 and (out == pl).all()
    imp = O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=9)
    pcm2 = O.encode(pl, channels=2, imp=imp)
    st2, out2, tp2 = O.decode(pcm2, channels=2)
    assert st2 == 0 and (out2 == pl).all()
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"),
                        payload=pl, pcm_sha=np.frombuffer(hashlib.sha256(pcm.tobytes()).digest(), np.uint8),
                        pcm_head=pcm[9500:9800].copy(), sc_pos=np.int32(tp.sc_pos), cfo_rad=np.float32(tp.cfo_rad),
                        soft=O.taps_np(tp, "soft")[:255], precision=O.taps_np(tp, "precision"),
                        llr_head=O.taps_np(tp, "llr")[:256],
                        imp_sc_pos=np.int32(tp2.sc_pos), imp_cfo_rad=np.float32(tp2.cfo_rad), imp_metrics=O.taps_np(tp2, "metrics"),
                        imp_slope=O.taps_np(tp2, "slope"), imp_flips=np.int32(tp2.flips))
    print(json.dumps(dig, indent=1))


if __name__ == "__main__":
    main()

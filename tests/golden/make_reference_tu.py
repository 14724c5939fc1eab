"""Regenerates tests/golden/reference_tu.json.  Run in the build container (needs /root/reference):

    python tests/golden/make_reference_tu.py

For each case of tests/test_reference_tu.py::golden_cases the stimulus comes from the oracle's encoder (its clean streams are
byte-identical to the reference encode.cc's, test_encoder_streams_are_identical) and the recorded outputs — payload digest
and stderr — come from oracle/_ref/decode: the reference's own decode.cc compiled against oracle/shim/ (restated
third-party primitives).  NOT output of a stock reference build: that cannot exist here."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import test_reference_tu as T  # noqa: E402


def main():
    O.build()
    subprocess.run(["make", "-s", "-C", os.path.join(T.ROOT, "oracle"), "ref"], check=True)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, kw, imp, skip in T.golden_cases(O):
            pls, pcm, rate, ch = T.golden_stimulus(O, kw, imp)
            wav = os.path.join(d, "g.wav")
            T.write_wav(wav, pcm, rate, ch)
            r = subprocess.run([os.path.join(T.REF, "decode"), os.path.join(d, "r.dat"), wav] + ([str(skip)] if skip is not None else []), capture_output=True)
            assert r.returncode == 0
            payload = open(os.path.join(d, "r.dat"), "rb").read()
            if b"bit flips:" not in r.stderr:   # failed decode: the reference's buffer is uninitialised (decode.cc:588); record the
                payload = bytes(np.frombuffer(bytes(5380), np.uint8) ^ T.descramble_stream())   # oracle's defined outcome instead
            out[name] = {"pcm_sha256": hashlib.sha256(np.ascontiguousarray(pcm, "<i2").tobytes()).hexdigest(),
                         "payload_sha256": hashlib.sha256(payload).hexdigest(),
                         "payload_is_sent": bool(any(payload == p.tobytes() for p in pls)),
                         "stderr": T._stderr_lines(r.stderr)}
            print(name, out[name]["payload_is_sent"], out[name]["stderr"][:4])
    json.dump(out, open(os.path.join(HERE, "reference_tu.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

"""North-star gate (gpurun): 10^5 distinct clean mode-6 frames, 8000 Hz 16-bit real — payload bit errors of the B200 path
against the sent bytes on every frame and against the CPU oracle's decode on a strided sample.  Writes profiles/clean100k_<tag>.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
total, chunk = int(os.environ.get("N_FRAMES", "100000")), 10000
cores = os.cpu_count() or 1
rx = M.Receiver(max_frames=chunk)
bit_err = ok = flips = oracle_checked = oracle_equal = 0
gpu_s = enc_s = 0.0
for c0 in range(0, total, chunk):
    n = min(chunk, total - c0)
    t = time.time()
    pcm, ns, sent = O.encode_batch(n, seed0=9_000_000 + c0, nthreads=cores)
    enc_s += time.time() - t
    t = time.time()
    payload, st = rx.decode(pcm)
    gpu_s += time.time() - t
    bit_err += int(np.unpackbits(payload ^ sent, axis=1).sum())
    ok += int((st["status"] == 0).sum())
    flips += int(st["flips"].clip(min=0).sum())
    for i in range(0, n, 200):   # oracle on a strided sample
        ost, opay, _ = O.decode(pcm[i], want_taps=False)
        oracle_checked += 1
        oracle_equal += int(ost == st["status"][i] and (opay == payload[i]).all())
    print("frames %6d: bit errors so far %d, status ok %d, oracle %d/%d" % (c0 + n, bit_err, ok, oracle_equal, oracle_checked), flush=True)
res = {"frames": total, "payload_bit_errors_vs_sent": bit_err, "frames_status_ok": ok, "bit_flips_reported": flips,
       "oracle_sample": oracle_checked, "oracle_sample_equal": oracle_equal,
       "gpu_wall_s_host_path": gpu_s, "encode_s": enc_s, "workload": "distinct clean mode-6 frames, 8000 Hz 16-bit real, seeds 9000000.."}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "profiles", "clean100k_%s.json" % tag), "w"), indent=1)
rx.close()

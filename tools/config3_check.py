"""BASELINE configs[2] at scale (gpurun): N mode-6 frames through the README impairment chain (multipath + CFO 234.567 Hz +
SFO 147 ppm + AWGN -30 dB, per-frame noise seed) — B200 path vs CPU oracle on identical windows: status and payload of
every window, plus the stage times of the batch.  Writes profiles/config3_<tag>.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import modem_b200 as M
import oracle_lib as O
from _stimulus import windows

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
n = int(os.environ.get("N_FRAMES", "10000"))
cores = os.cpu_count() or 1
t = time.time()
pcm, ns, sent = windows(n, 31337, channels=2, imp=dict(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=12345))  # STIM=device: generated on the GPU
t_enc = time.time() - t
rx = M.Receiver(max_frames=n)
rx.set_option("sub_chunks", 1)   # one list-decoder launch per chunk: stage_times() adds up
d = torch.from_numpy(pcm).cuda()
out = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
st = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ms = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    rx.decode_raw(d.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, n, pcm.shape[1] // 2, None, 0, out.data_ptr(), st.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
stages, _ = rx.stage_times()
gp = out.cpu().numpy(); gs = st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
t = time.time()
ost, op = O.decode_batch(pcm, channels=2, nthreads=cores)
t_cpu = time.time() - t
res = {"frames": n, "impairments": "multipath(4 taps) + CFO 234.567 Hz + SFO 147 ppm + AWGN -30 dB, 2-channel int16",
       "gpu_ok": int((gs["status"] == 0).sum()), "cpu_ok": int((ost == 0).sum()), "status_equal": int((gs["status"] == ost).sum()),
       "payload_equal_windows": int((gp == op).all(axis=1).sum()), "payload_bit_errors_vs_sent_gpu": int(np.unpackbits(gp ^ sent, axis=1).sum()),
       "gpu_ms_per_batch": min(ms), "gpu_frames_per_s": n / min(ms) * 1e3, "stage_ms": stages,
       "cpu_oracle_s": t_cpu, "cpu_frames_per_s": n / t_cpu, "cpu_threads": cores, "encode_s": t_enc, "stimulus": os.environ.get("STIM", "cpu")}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "profiles", "config3_%s.json" % tag), "w"), indent=1)
rx.close()

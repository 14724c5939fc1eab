"""Developer probe (gpurun): stage breakdown of a full-residency batch with device-resident inputs."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import modem_b200 as M
import oracle_lib as O

n = int(os.environ.get("N_FRAMES", "9472"))
uniq = int(os.environ.get("N_UNIQ", "592"))
ctas = int(os.environ.get("CTAS", "0")) or None
reps = int(os.environ.get("REPS", "3"))
t = time.time()
impaired = os.environ.get("IMPAIR", "0") != "0"   # README chain on 2-channel windows instead of clean mono ones
if impaired:
    pcm_u, ns, pay_u = O.encode_batch(uniq, seed0=1, channels=2, imp=O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=99))
else:
    pcm_u, ns, pay_u = O.encode_batch(uniq, seed0=1)
print("encode %d frames: %.1f s" % (uniq, time.time() - t), flush=True)
idx = np.arange(n) % uniq
pcm = torch.from_numpy(pcm_u)[torch.from_numpy(idx)].cuda()
rx = M.Receiver(max_frames=n, scl_ctas_per_sm=ctas)
rx.set_option("sub_chunks", 1)   # one list-decoder launch per chunk: stage_times() adds up
payload = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
status = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ if impaired else M.FMT_S16_MONO, n, pcm.shape[1] // (2 if impaired else 1), None, 0, payload.data_ptr(), status.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st, nch = rx.stage_times()
    print("rep %d: %.1f ms -> %.0f frames/s | " % (rep, ms, n / ms * 1e3) + " ".join("%s=%.2f" % kv for kv in st.items()), flush=True)
stat = status.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
pay = payload.cpu().numpy()
print("theil-sen sweeps per row: %.3f" % (stat["ts_sweeps"].sum() / (50.0 * n)))
print("ok frames", int((stat["status"] == 0).sum()), "payload match", int((pay == pay_u[idx]).all(axis=1).sum()), "of", n)

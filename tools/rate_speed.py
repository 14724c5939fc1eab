"""Developer probe (gpurun): stage times of a batch at another sample rate / mode (device-resident mono windows)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import modem_b200 as M
import oracle_lib as O
rate, mode = int(os.environ.get("RATE", "48000")), int(os.environ.get("MODE", "6"))
n, uniq = int(os.environ.get("N_FRAMES", "1184")), int(os.environ.get("N_UNIQ", "74"))
stride = O.frame_samples(mode, rate)
pcm_u, ns, pay_u = O.encode_batch(uniq, seed0=5, rate=rate, mode=mode, stride=stride)
idx = np.arange(n) % uniq
pcm = torch.from_numpy(pcm_u)[torch.from_numpy(idx)].cuda()
rx = M.Receiver(max_frames=n, max_samples=stride, rate=rate)
rx.set_option("sub_chunks", 1)   # one list-decoder launch per chunk: stage_times() adds up
payload = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
status = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_MONO, n, stride, None, 0, payload.data_ptr(), status.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
st, _ = rx.stage_times()
ok = int((payload.cpu().numpy() == pay_u[idx]).all(axis=1).sum())
print("rate %d mode %d: %d windows of %d samples in %.1f ms -> %.0f frames/s | %s | payload match %d" % (rate, mode, n, stride, ms, n / ms * 1e3, " ".join("%s=%.2f" % kv for kv in st.items()), ok))

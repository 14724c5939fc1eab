"""Developer probe (gpurun): throughput of the device-side stimulus generator (include/ofdmtx.h) with payloads and samples
resident in HBM, clean and with the README.md:49 impairment chain, followed by a device loop-back (generate -> decode)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import modem_b200 as M

n = int(os.environ.get("N_FRAMES", "2048"))
mode = int(os.environ.get("MODE", "6"))
rate = int(os.environ.get("RATE", "8000"))
reps = int(os.environ.get("REPS", "3"))
tx = M.Transmitter(max_windows=n, rate=rate)
stride = tx.window_samples(mode)
rng = np.random.default_rng(1)
pls = torch.from_numpy(rng.integers(0, 256, (n, M.PAYLOAD_BYTES), dtype=np.uint8)).cuda()
pcm = torch.zeros((n, stride * 2), dtype=torch.int16, device="cuda")
cs = int(M.load().ofdmtx_call_sign(b"CALLSIGN"))
stream = torch.cuda.current_stream().cuda_stream
chain = M.impairments(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=5)
for name, imp in (("clean", None), ("awgn", M.impairments(awgn_db=-20.0, seed=5)), ("readme chain", chain)):
    for rep in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tx.encode_raw(pls.data_ptr(), M.MEM_DEVICE, n, mode, cs, 2000, imp, pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, stride, None, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print("%-13s %d windows: %.1f ms -> %.0f frames/s (%d launches)" % (name, n, ms, n / ms * 1e3, tx.last_launches), flush=True)
# loop-back on the impaired windows
rx = M.Receiver(max_frames=n, max_samples=stride, rate=rate)
pay = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda")
st = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, n, stride, None, 0, pay.data_ptr(), st.data_ptr(), stream)
torch.cuda.synchronize()
stat = st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
print("loop-back: ok windows", int((stat["status"] == 0).sum()), "payload match", int((pay == pls).all(dim=1).sum().item()), "of", n)

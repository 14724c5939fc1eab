"""Developer probe (gpurun, optionally under ncu): the list decoder on 10 000 device-generated windows of one impairment class.
IMP = clean | chain | awgn25 | awgn18; SEED = noise seed of the chain"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
n = int(os.environ.get("FRAMES", "10000"))
kind = os.environ.get("IMP", "chain")
seed = int(os.environ.get("SEED", "5"))
imp = {"clean": None, "chain": M.impairments(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=seed),
       "awgn25": M.impairments(awgn_db=-25.0, seed=6), "awgn18": M.impairments(awgn_db=-18.0, seed=7)}[kind]
tx = M.Transmitter(max_windows=2048)
stride = tx.window_samples(6) + 64
rx = M.Receiver(max_frames=n, max_samples=stride)
cs = int(M.load().ofdmtx_call_sign(b"CALLSIGN"))
s = torch.cuda.current_stream().cuda_stream
pseed = os.environ.get("PSEED")   # payload seed (bench.py config3: 777)
gen = torch.Generator(device="cuda").manual_seed(int(pseed)) if pseed else None
sent = torch.randint(0, 256, (n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda", generator=gen)
pcm = torch.zeros((n, 2 * stride), dtype=torch.int16, device="cuda")
tx.encode_raw(sent.data_ptr(), M.MEM_DEVICE, n, 6, cs, 2000, imp, pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, stride, None, s)
pay = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda"); st = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
for _ in range(3):
    rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, n, stride, None, 0, pay.data_ptr(), st.data_ptr(), s)
torch.cuda.synchronize()
ms, _ = rx.stage_times()
stat = st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
ok = stat["status"] == 0
err = int(np.unpackbits((pay ^ sent).cpu().numpy()[ok]).sum())
sw = stat["ts_sweeps"][ok]
print("windows with a bisection fallback (>= 100 sweeps):", int((sw >= 100).sum()), "max sweeps in a window:", int(sw.max()), "windows with > 75 sweeps:", int((sw > 75).sum()))
print("%s: %d windows, ok %d, bit errors %d, sweeps per row %.3f, stage ms %s" % (kind, n, int(ok.sum()), err, stat["ts_sweeps"][ok].mean() / 50.0, {k: round(v, 2) for k, v in ms.items()}), flush=True)

"""Developer probe (gpurun): decode the README-chain batch of bench.py's config3 (noise seed 4242, payload seed 777) and save the
phase rows (tap PHASE) and Theil-Sen outputs of every window in which a row went through the bisection fallback (>= 100 sweeps)
to gpurun_out/ts_fallback_rows.npz — input for tools/ts_search_emulation.py."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modem_b200 as M
n = int(os.environ.get("FRAMES", "10000"))
imp = M.impairments(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=int(os.environ.get("SEED", "4242")))
tx = M.Transmitter(max_windows=2048)
stride = tx.window_samples(6) + 64
rx = M.Receiver(max_frames=n, max_samples=stride, keep_taps=True)
cs = int(M.load().ofdmtx_call_sign(b"CALLSIGN"))
s = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device="cuda").manual_seed(int(os.environ.get("PSEED", "777")))
sent = torch.randint(0, 256, (n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda", generator=gen)
pcm = torch.zeros((n, 2 * stride), dtype=torch.int16, device="cuda")
tx.encode_raw(sent.data_ptr(), M.MEM_DEVICE, n, 6, cs, 2000, imp, pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, stride, None, s)
pay = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda"); st = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
rx.decode_raw(pcm.data_ptr(), M.MEM_DEVICE, M.FMT_S16_IQ, n, stride, None, 0, pay.data_ptr(), st.data_ptr(), s)
torch.cuda.synchronize()
stat = st.cpu().numpy().view(M.STATUS_DTYPE).reshape(-1)
bad = np.nonzero((stat["status"] == 0) & (stat["ts_sweeps"] >= 100))[0]
print("windows with a fallback row:", bad.tolist(), "sweeps", stat["ts_sweeps"][bad].tolist(), flush=True)
if len(bad):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "ts_fallback_rows.npz"), windows=bad,
             phase=np.stack([rx.taps(M.TAP_PHASE, int(f), 1)[0] for f in bad]), ts=np.stack([rx.taps(M.TAP_TS, int(f), 1)[0] for f in bad]))
    print("saved", flush=True)

#!/bin/bash
# Developer probe (gpurun): sub-chunk pipeline A/B (list decoder of sub-chunk k beside the front stages of sub-chunk k+1).
for sc in ${1:-1 2 4 8}; do
python - <<PY
import os, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import modem_b200 as M, oracle_lib as O
n = int(os.environ.get("FRAMES", "10000"))
pcm, ns, sent = O.encode_batch(n, seed0=3)
dev = torch.from_numpy(pcm).cuda()
pay = torch.empty((n, M.PAYLOAD_BYTES), dtype=torch.uint8, device="cuda"); st = torch.empty((n, 112), dtype=torch.uint8, device="cuda")
rx = M.Receiver(max_frames=n)
rx.set_option("sub_chunks", $sc)
s = torch.cuda.current_stream().cuda_stream
def step(): rx.decode_raw(dev.data_ptr(), M.MEM_DEVICE, M.FMT_S16_MONO, n, 95200, None, 0, pay.data_ptr(), st.data_ptr(), s)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
err = int(np.unpackbits(pay.cpu().numpy() ^ sent).sum())
print("sub_chunks=$sc frames %d: %.2f ms per step, %.0f frames/s, bit errors %d" % (n, ms, n / ms * 1e3, err), flush=True)
PY
done

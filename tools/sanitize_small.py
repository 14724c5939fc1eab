"""Developer probe (gpurun, under compute-sanitizer): a tiny mixed-mode batch through every kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import modem_b200 as M
import oracle_lib as O

frames = [O.encode_batch(1, seed0=70 + m, mode=m) for m in (6, 13, 10, 8)]
stride = max(f[0].shape[1] for f in frames)
pcm = np.zeros((6, stride), np.int16)
ns = np.zeros(6, np.int32)
for k, f in enumerate(frames):
    pcm[k, :f[0].shape[1]] = f[0]; ns[k] = f[0].shape[1]
pcm[4, :30000] = frames[0][0][0, :30000]; ns[4] = 30000      # cut inside the payload
ns[5] = 5000                                                  # silence
sent = np.concatenate([f[2] for f in frames])
rx = M.Receiver(max_frames=6, max_samples=stride, keep_taps=True)
payload, st = rx.decode(pcm, n_samples=ns)
print("status", list(st["status"]), "modes", list(st["mode"]), "payload ok", bool((payload[:4] == sent).all()))
rng = np.random.default_rng(1)
y = (0.03 * rng.standard_normal((7, 432))).astype(np.float32)
print("theil-sen", rx.theil_sen(y)[0][:3])
rx.close()

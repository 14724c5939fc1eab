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
# noisy windows (eight distinct paths from the start / from half way: class loops, ranked forks, the general C op) through a handle
# WITHOUT stage taps (the timing metric is then stored only around detections: initcheck sees any read of an unwritten tile),
# including a float-sample window and a window cut short
imp = [O.impair(awgn_db=-16.0, seed=1), O.impair(multipath=True, cfo_hz=234.567, sfo_ppm=147, awgn_db=-30, seed=2), O.impair(awgn_db=-24.0, seed=3), None]
wins = [O.encode_batch(1, seed0=90 + k, channels=2, imp=im) for k, im in enumerate(imp)]
stride = max(w[0].shape[1] for w in wins) // 2
pcm2 = np.zeros((5, 2 * stride), np.int16); ns2 = np.zeros(5, np.int32)
for k, w in enumerate(wins):
    pcm2[k, :w[0].shape[1]] = w[0]; ns2[k] = w[0].shape[1] // 2
pcm2[4, :2 * 50000] = wins[3][0][0, :2 * 50000]; ns2[4] = 50000
rx2 = M.Receiver(max_frames=5, max_samples=stride)
p2, st2 = rx2.decode(pcm2, channels=2, n_samples=ns2)
print("noisy status", list(st2["status"]), "payload ok", [bool((p2[k] == wins[k][2][0]).all()) for k in range(4)])
pf, stf = rx2.decode((pcm2[:2].astype(np.float32) / np.float32(32767.0)), channels=2, n_samples=ns2[:2])
print("float status", list(stf["status"]), "same payload", bool((pf == p2[:2]).all()))
rx2.close()

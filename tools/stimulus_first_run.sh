#!/bin/bash
# First hardware run of the stimulus generator (it was written and CPU-checked without a GPU).  One gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/stimulus_first_run.sh'
# Order: memcheck on a tiny batch first (a wild pointer then shows up as a report, not as a dead context), then the
# parity tests, then the throughput probe.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m modem_b200.build > gpurun_out/stim_build.log 2>&1 || { tail -20 gpurun_out/stim_build.log; exit 1; }
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/stim_memcheck.log 2>&1 <<'PY'
import numpy as np, modem_b200 as M
for rate in (8000, 48000):
    tx = M.Transmitter(max_windows=2, rate=rate)
    pl = np.random.default_rng(0).integers(0, 256, (3, M.PAYLOAD_BYTES), dtype=np.uint8)     # 3 windows through 2-window chunks
    for imp in (None, M.impairments(multipath=True, cfo_hz=234.567, sfo_ppm=147.0, awgn_db=-30.0, seed=1)):
        pcm, ns = tx.encode(pl, mode=6 if rate == 8000 else 13, channels=2, imp=imp)
        print(rate, ns, int(np.abs(pcm).max()))
    tx.close()
PY
echo "memcheck exit $?"; tail -5 gpurun_out/stim_memcheck.log
timeout 600 python -m pytest tests/test_gpu_stimulus.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/stim_tests.log
timeout 300 python tools/tx_speed.py 2>&1 | tee gpurun_out/stim_speed.log

#!/bin/bash
# Developer probe (gpurun): A/B of Theil-Sen builds (modem_b200/libofdmrx*.so variants made with OFDMRX_TS_* build switches) and of
# the chains-per-window launch parameter: stage time of a 10 000-window step and the sweep histogram on noisy rows.
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "theil" 2>&1 | tail -2
probe() { for k in $1; do IMP=$k timeout 200 python tools/scl_noisy_probe.py 2>&1 | tail -1 | sed -e "s/.frontend.*.demod_fft/demod_fft/"; done; }
for lib in modem_b200/libofdmrx*.so; do
  echo "== $lib"; export OFDMRX_LIB=$PWD/$lib
  probe "clean chain awgn18"
  timeout 100 python tools/ts_probe.py 2>&1 | tail -4 | cut -c1-90
done
export OFDMRX_LIB=$PWD/modem_b200/libofdmrx.so
for c in 10 25 50; do echo "== chains $c"; OFDMRX_TS_CHAINS=$c probe "clean chain"; done

#!/bin/bash
# Developer probe (gpurun): ncu A/B of the correlator kernel with and without the bulk-copy (TMA) staging of the tile span.
for v in 0 1; do
  OFDMRX_SYNC_TMA=$v python -m modem_b200.build --force > /dev/null 2>&1
  BENCH_CONFIG3=0 BENCH_CONFIG5=0 BENCH_E2E_PIPELINE=0 timeout 600 ncu --set full --clock-control none -k regex:k_sync_metric --launch-count 1 -o gpurun_out/prof_sync_tma$v -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_sync_tma$v.log 2>&1
done
python -m modem_b200.build --force > /dev/null 2>&1

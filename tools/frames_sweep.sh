#!/bin/bash
# Developer probe (gpurun): throughput against the batch size (wave robustness of the list decoder).
for f in ${1:-5000 10000 10100 12000 20000 40000}; do
  BENCH_CONFIG3=0 BENCH_CONFIG5=0 BENCH_E2E_PIPELINE=0 timeout 900 python bench.py --steps 3 --warmup 3 --frames $f 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('| %d | %.0f | %.2f | %.2f | %.2f | %.2f | %d |' % ($f, d['frames_per_s'], d['ms_per_step'], s['polar_scl'], s['theil_sen'], s['sync_metric'], d['parity']['payload_bit_errors_vs_sent']))"
done

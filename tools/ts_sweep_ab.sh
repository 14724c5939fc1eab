#!/bin/bash
# Developer probe (gpurun): Theil-Sen first-bracket width A/B — demod stage time of a 10 000-window step and the sweep histogram.
for k in $1; do
  echo "== OFDMRX_TS_HALF=$k"
  OFDMRX_TS_HALF=$k timeout 300 python tools/ts_probe.py 2>&1 | tail -4
  OFDMRX_TS_HALF=$k BENCH_E2E_PIPELINE=0 timeout 600 python bench.py --steps 3 --warmup 3 --frames ${FRAMES:-10000} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('demod %.2f ms (clean)  %.2f ms (README chain); errors %d %d' % (d['stage_ms']['demod'], d['config3']['stage_ms']['demod'], d['parity']['payload_bit_errors_vs_sent'], d['config3']['payload_bit_errors_vs_sent']))"
done

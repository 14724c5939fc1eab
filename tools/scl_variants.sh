#!/bin/bash
# Developer probe (gpurun): rebuild the library with different list-decoder occupancy targets and print the stage times
# of a 10 000-window step (clean and README chain).   usage: bash tools/scl_variants.sh "12 16 17 21"
set -u
mkdir -p gpurun_out
for n in $1; do
  OFDMRX_SCL_CTAS=$n python -m modem_b200.build --force > gpurun_out/variant_build_$n.log 2>&1 || { tail -5 gpurun_out/variant_build_$n.log; continue; }
  grep -A2 "k_polar_scl" gpurun_out/variant_build_$n.log | grep -E "Used|spill" | tail -2
  BENCH_E2E_PIPELINE=0 timeout 600 python bench.py --steps 3 --warmup 3 --frames ${FRAMES:-10000} > gpurun_out/variant_$n.json 2> gpurun_out/variant_$n.err || tail -3 gpurun_out/variant_$n.err
  python - <<PY
import json
d = json.load(open("gpurun_out/variant_$n.json"))
print("CTAS=$n frames/s %.0f step %.2f ms scl %.2f demod %.2f | cfg3 %.0f f/s scl %.2f | errors %d %d" % (d["frames_per_s"], d["ms_per_step"], d["stage_ms"]["polar_scl"], d["stage_ms"]["demod"],
      d["config3"]["frames_per_s"], d["config3"]["stage_ms"]["polar_scl"], d["parity"]["payload_bit_errors_vs_sent"], d["config3"]["payload_bit_errors_vs_sent"]))
PY
done

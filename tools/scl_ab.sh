#!/bin/bash
# Developer probe (gpurun): A/B of list-decoder build switches.  usage: bash tools/scl_ab.sh "OFDMRX_SCL_PREFETCH=0" "OFDMRX_SCL_PREFETCH=3" ...
set -u
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v python -m modem_b200.build --force > gpurun_out/ab_build_$tag.log 2>&1 || { tail -5 gpurun_out/ab_build_$tag.log; continue; }
  BENCH_E2E_PIPELINE=0 timeout 600 python bench.py --steps 3 --warmup 3 --frames ${FRAMES:-10000} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -3 gpurun_out/ab_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$tag.json"))
print("$v: frames/s %.0f step %.2f ms scl %.2f demod %.2f | cfg3 %.0f f/s scl %.2f | errors %d %d" % (d["frames_per_s"], d["ms_per_step"], d["stage_ms"]["polar_scl"], d["stage_ms"]["demod"],
      d["config3"]["frames_per_s"], d["config3"]["stage_ms"]["polar_scl"], d["parity"]["payload_bit_errors_vs_sent"], d["config3"]["payload_bit_errors_vs_sent"]))
PY
done
python -m modem_b200.build --force > /dev/null 2>&1

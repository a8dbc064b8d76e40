#!/usr/bin/env bash
# Round-2 GPU call: panel-pool size (batches per grid) against the step time, whole grid N = 1
set -u
OUT=gpurun_out; mkdir -p $OUT
for mb in 8192 24576 49152 4096; do
  GIMIC_B200_POOL_MB=$mb timeout 400 python bench.py --steps 3 --warmup 3 --no-extras > $OUT/r02pool_$mb.json 2>/dev/null
  python - $OUT/r02pool_$mb.json $mb <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   pool %6s MB: ms/step %.2f  e2e ms %.2f  plan %.2f basis %.2f contract %.2f  launches/step %.0f  frac %.3f" % (sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], s["ms_plan"], s["ms_basis"], s["ms_contract"], r["launches_timed"]/2, r["frac"]))
PY
done

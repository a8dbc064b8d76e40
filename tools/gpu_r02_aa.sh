#!/usr/bin/env bash
# Round-2 GPU call AA: drain groups (host outputs copied out group by group when the range has few batches): parity of the partition paths,
# octant and whole-grid bench (value and e2e).
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "partition or chunked or synthetic_flake or full_size or point_order" ) 2>&1 | tail -1
for m in octant grid; do
  timeout 600 python bench.py --mode $m --steps 5 --warmup 3 --no-extras > $OUT/r02aa_$m.json 2>/dev/null
  python - $OUT/r02aa_$m.json $m <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-7s ms/step %.2f  pts/s %.3fM  e2e ms %.2f  e2e %.3fM  basis %.2f contract %.2f launches/step %.0f" % (sys.argv[2], d["ms_per_step"], d["value"]/1e6, d["e2e"].get("ms_per_step", 0), d["e2e"]["value"]/1e6, s["ms_basis"], s["ms_contract"], r["launches_timed"]/2))
PY
done
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/r02aa_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/r02aa_pytest_gpu.log | tail -1; grep -E "^FAILED|^ERROR" $OUT/r02aa_pytest_gpu.log | head -5

#!/usr/bin/env bash
# Round-2 GPU call L: cost-based gap splitting, dead-warp skip for partial tiles, slice count from the whole set: GPU suite, headline bench,
# launch list for the plan kernels.
set -u
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
if ( timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid or c4h4_integration or j_path or synthetic_flake" ) > $OUT/${TAG}_pytest_quick.log 2>&1; then
  tail -1 $OUT/${TAG}_pytest_quick.log
  ( time timeout 900 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  python - $OUT/${TAG}_bench_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f sort %.2f tiles %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
print("   integral_36x36", {k: v for k, v in d["stages"]["integral_36x36"].items() if k != "sums"})
print("   config4", d["stages"]["config4"]["points_per_s"], d["stages"]["config4"]["parity_jvec_max_scaled_err_64pts"], " parity", d["cpu_baseline"]["parity_max_scaled_err"])
PY
  GIMIC_B200_SLICES=0 timeout 300 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1_noslices.json 2>/dev/null
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n1_noslices.json')); print('   no slices: integral_36x36', {k: v for k, v in d['stages']['integral_36x36'].items() if k != 'sums'})"
  timeout 200 python tools/legacy_latency.py 2>&1 | tail -1
  timeout 300 python tools/time_driver.py --native 2>&1 | tail -8
else
  echo "quick parity FAILED or timed out"; tail -20 $OUT/${TAG}_pytest_quick.log
fi

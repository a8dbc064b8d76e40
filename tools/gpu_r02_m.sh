#!/usr/bin/env bash
# Round-2 GPU call M: A/B of the contraction kernel with the dead-warp path outside the hot loop (libgimic_b200.so) against the kernel
# of the commit before (libgimic_b200_old.so, built by hand from git history), octant step; then the headline bench.
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-10s ms/step %.2f  pts/s %.3fM  plan %.2f sort %.2f tiles %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (sys.argv[2], d["ms_per_step"], d["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
if "integral_36x36" in d.get("stages", {}): print("   integral_36x36", {k: v for k, v in d["stages"]["integral_36x36"].items() if k not in ("sums", "what")})
PY
}
if ( timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid or c4h4_integration or j_path or synthetic_flake or slice or thin_point or partition_points" ) > $OUT/${TAG}_pytest_quick.log 2>&1; then
  tail -1 $OUT/${TAG}_pytest_quick.log
  for v in new old new old; do
    if [ $v = old ]; then export GIMIC_B200_LIB=$PWD/gimic_b200/libgimic_b200_old.so; else unset GIMIC_B200_LIB; fi
    timeout 300 python bench.py --mode octant --steps 5 --warmup 3 > $OUT/${TAG}_octant_$v.json 2>/dev/null && pr $OUT/${TAG}_octant_$v.json "octant/$v"
  done
  unset GIMIC_B200_LIB
  timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err && pr $OUT/${TAG}_bench_n1.json grid/new
  GIMIC_B200_LIB=$PWD/gimic_b200/libgimic_b200_old.so timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1_old.json 2>/dev/null && pr $OUT/${TAG}_bench_n1_old.json grid/old
else
  echo "quick parity FAILED or timed out"; tail -20 $OUT/${TAG}_pytest_quick.log
fi

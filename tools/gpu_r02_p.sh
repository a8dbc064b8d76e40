#!/usr/bin/env bash
# Round-2 GPU call P: FP32 minimum-distance search in the plan / basis kernels, cost-proportional column slices, k_basis atom groups for few tiles.
set -u
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-10s ms/step %.2f  pts/s %.3fM  e2e %.3fM plan %.2f sort %.2f tiles %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (sys.argv[2], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
if "integral_36x36" in d.get("stages", {}): print("   integral_36x36", {k: v for k, v in d["stages"]["integral_36x36"].items() if k not in ("sums", "what")})
PY
}
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "random_molecules or c4h4_read_grid or c4h4_integration or j_path or synthetic_flake or slice or thin_point or partition_points or hot_path" ) > $OUT/${TAG}_pytest_quick.log 2>&1; tail -1 $OUT/${TAG}_pytest_quick.log; grep -E "^E  " $OUT/${TAG}_pytest_quick.log | head -5
timeout 120 python tools/plane_probe.py 36 5 2>&1 | tail -1
timeout 120 python tools/plane_probe.py 72 5 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err && pr $OUT/${TAG}_bench_n1.json grid
timeout 300 python bench.py --mode octant --steps 5 --warmup 3 > $OUT/${TAG}_octant.json 2>/dev/null && pr $OUT/${TAG}_octant.json octant
( time timeout 900 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2; grep -E "^FAILED" $OUT/${TAG}_pytest_gpu.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

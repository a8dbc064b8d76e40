#!/usr/bin/env bash
# Round-2 GPU call N: contraction kernel back to the full-loop version, cost-based gap splitting, 32-bit sort keys: GPU suite, headline bench,
# launch list of the plan kernels on the whole grid.
set -u
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-10s ms/step %.2f  pts/s %.3fM  e2e %.3fM plan %.2f sort %.2f tiles %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (sys.argv[2], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
if "integral_36x36" in d.get("stages", {}): print("   integral_36x36", {k: v for k, v in d["stages"]["integral_36x36"].items() if k not in ("sums", "what")})
PY
}
if ( timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid or c4h4_integration or j_path or synthetic_flake or slice or thin_point or partition_points" ) > $OUT/${TAG}_pytest_quick.log 2>&1; then
  tail -1 $OUT/${TAG}_pytest_quick.log
  timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err && pr $OUT/${TAG}_bench_n1.json grid
  timeout 300 python bench.py --mode octant --steps 5 --warmup 3 > $OUT/${TAG}_octant.json 2>/dev/null && pr $OUT/${TAG}_octant.json octant
  ( time timeout 900 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_grid.csv python bench.py --steps 1 --warmup 0 --no-extras > /dev/null 2>&1
  python - $OUT/${TAG}_launches_grid.csv <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e3 if u in ("nsecond","ns") else v*1e3 if u in ("msecond","ms") else v
    n=r[ki][:60]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,t) in agg.items(): print("   %-62s x%-4d %10.1f us" % (n,c,t))
PY
else
  echo "quick parity FAILED or timed out"; tail -20 $OUT/${TAG}_pytest_quick.log
fi

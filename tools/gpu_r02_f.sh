#!/usr/bin/env bash
# Round-2 GPU call F: default build (epilogue warpgroup, pinned loop invariants): full GPU test suite, smoke, octant + headline bench,
# launch list of the headline command, reference arm.
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (all)"
( time timeout 900 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== octant bench"
timeout 300 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_octant.json 2> $OUT/${TAG}_bench_octant.err
python - $OUT/${TAG}_bench_octant.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM contract %.2f ms  basis %.2f  tiles %.2f sort %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_contract"], s["ms_basis"], s["ms_tiles"], s["ms_sort"], r["achieved"], r["frac"]))
PY
echo "== headline bench (whole grid)"
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - $OUT/${TAG}_bench_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f (sort %.2f tiles %.2f) basis %.2f contract %.2f  TF %.2f frac %.3f useful %.3f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"], r["useful_frac"]))
print("   cpu:", d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>/dev/null; tail -c 300 $OUT/${TAG}_bench_reference.json
echo "== ncu launch list of the headline command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1; head -14 $OUT/${TAG}_launches_summary.txt

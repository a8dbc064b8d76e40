#!/usr/bin/env bash
# Round-2 GPU call C: parity after the screening-sweep rewrite, the DMMA/DFMA tap microbenchmark, octant + headline bench, and ONE
# full ncu capture (with source) of the contraction kernel on the prof_step chunk.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu: kernel parity (whole file)"
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > $OUT/${TAG}_pytest_parity.log 2>&1; tail -4 $OUT/${TAG}_pytest_parity.log
echo "== tap microbenchmark"
timeout 120 tools/dmma_tap > $OUT/${TAG}_dmma_tap.txt 2>&1; cat $OUT/${TAG}_dmma_tap.txt
echo "== octant bench"
timeout 300 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_octant.json 2> $OUT/${TAG}_bench_octant.err
python - $OUT/${TAG}_bench_octant.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   ms/step %.2f  pts/s %.3fM  contract %.2f ms  basis %.2f  tiles %.2f sort %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, s["ms_contract"], s["ms_basis"], s["ms_tiles"], s["ms_sort"], r["achieved"], r["frac"]))
PY
echo "== headline bench (whole grid)"
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - $OUT/${TAG}_bench_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f (sort %.2f tiles %.2f) basis %.2f contract %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
print("   stages:", json.dumps(d["stages"])[:900])
PY
echo "== ncu --set full with source: one k_jtensor launch on the prof_step chunk"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jtensor -s 1 -c 1 -o $OUT/${TAG}_jtensor python tools/prof_step.py --reps 2 > $OUT/${TAG}_prof_step.log 2>&1
tail -3 $OUT/${TAG}_prof_step.log; ls -la $OUT/${TAG}_jtensor.ncu-rep

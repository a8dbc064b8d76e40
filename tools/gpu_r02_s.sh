#!/usr/bin/env bash
# Round-2 GPU call S: current / |J| integrals through the J path (2 operand planes): parity (integration tests vs oracle 1e-10 and goldens),
# plane probe, scan bench, whole suite, headline bench.
set -u
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "integr or scan or batch or random_molecules or slice" ) 2>&1 | tail -1
timeout 120 python tools/plane_probe.py 36 5 2>&1 | tail -1
timeout 120 python tools/plane_probe.py 72 5 2>&1 | tail -1
timeout 200 python tools/scan_bench.py 2>&1 | tail -16 | tee $OUT/${TAG}_scan_bench.json
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/${TAG}_pytest_gpu.log | head; grep -E "^E  " $OUT/${TAG}_pytest_gpu.log | head -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - $OUT/${TAG}_bench_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   grid ms/step %.2f  pts/s %.3fM  e2e %.3fM plan %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
print("   integral_36x36", {k: v for k, v in d["stages"]["integral_36x36"].items() if k not in ("sums", "what")})
PY

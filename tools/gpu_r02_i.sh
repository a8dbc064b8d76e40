#!/usr/bin/env bash
# Round-2 GPU call I: J path with the epilogue warpgroup (parity, A/B on the prof_step chunk), new cost model, headline + octant bench.
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
echo "== quick parity of the J path kernel (guarded), then the whole GPU suite"
if ( timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "j_path or c4h4_read_grid or partition_points" ) > $OUT/${TAG}_pytest_quick.log 2>&1; then
  tail -1 $OUT/${TAG}_pytest_quick.log
  ( time timeout 900 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -4 $OUT/${TAG}_pytest_gpu.log
  echo "== J path A/B on the prof_step chunk (151552 points): GIMIC_B200_EPI=0 then 1"
  GIMIC_B200_EPI=0 timeout 200 python tools/prof_step.py --jvec --reps 3 2>&1 | tail -2
  GIMIC_B200_EPI=1 timeout 200 python tools/prof_step.py --jvec --reps 3 2>&1 | tail -2
  echo "== octant + headline"
  timeout 300 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_octant.json 2> $OUT/${TAG}_bench_octant.err
  timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  python - $OUT/${TAG}_bench_octant.json $OUT/${TAG}_bench_n1.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f)); r=d["roofline"]; s=d["stage_ms_per_step"]
    print("   %s: ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f sort %.2f tiles %.2f basis %.2f contract %.2f  TF %.2f frac %.3f" % (f.split('_bench_')[1], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_sort"], s["ms_tiles"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
PY
else
  echo "quick parity FAILED or timed out"; tail -15 $OUT/${TAG}_pytest_quick.log
fi

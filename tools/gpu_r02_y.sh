#!/usr/bin/env bash
# Round-2 GPU call Y: refreshed ring-geometry and general-P bench lines, the J path on one octant (time + ncu --set full of k_jtensor_e<GIAO,JVEC>)
set -u
OUT=gpurun_out; mkdir -p $OUT
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-10s ms/step %.2f  pts/s %.3fM  e2e %.3fM plan %.2f basis %.2f contract %.2f  TF %.2f frac %.3f  active-points/s %.1fM" % (sys.argv[2], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"], r["active_points_per_s"]/1e6))
PY
}
timeout 400 python bench.py --geometry ring --steps 5 --warmup 3 --no-extras > $OUT/r02y_bench_n1_ring.json 2>/dev/null && pr $OUT/r02y_bench_n1_ring.json ring
timeout 600 python bench.py --general-p --steps 3 --warmup 3 --no-extras > $OUT/r02y_bench_n1_general_p.json 2>/dev/null && pr $OUT/r02y_bench_n1_general_p.json general-P
timeout 300 python tools/jpath_probe.py 2>&1 | tail -3 | tee $OUT/r02y_jpath_probe.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_jtensor_e -s 1 -c 1 -o $OUT/r02y_jpath python tools/jpath_probe.py > /dev/null 2>&1
ls -la $OUT/r02y_jpath.ncu-rep 2>/dev/null | awk '{print "   ncu-rep bytes", $5}'

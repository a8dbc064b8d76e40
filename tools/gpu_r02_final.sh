#!/usr/bin/env bash
# Round-2 final check of the tree as committed: GPU suite, smoke, headline bench, launch list of one step.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/r02final_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/r02final_pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/r02final_pytest_gpu.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r02final_bench_n1.json 2> $OUT/r02final_bench_n1.err
python - $OUT/r02final_bench_n1.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   grid ms/step %.2f  pts/s %.3fM  e2e %.3fM plan %.2f basis %.2f contract %.2f  TF %.2f frac %.3f launches %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"], d.get("gpu_launches")))
print("   stages:", {k: (v.get("roofline", {}).get("frac") if isinstance(v, dict) and "roofline" in v else v.get("wall_us", v.get("points_per_s"))) for k, v in d["stages"].items()})
print("   cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], "clocks", d["clocks"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02final_launches.csv python bench.py --steps 1 --warmup 0 --no-extras > /dev/null 2>&1
python - $OUT/r02final_launches.csv <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e3 if u in ("nsecond","ns") else v*1e3 if u in ("msecond","ms") else v
    n=r[ki][:60]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
for n,(c,t) in agg.items(): print("   %-62s x%-4d %10.1f us  %5.2f %%" % (n,c,t,100*t/tot))
PY

#!/usr/bin/env bash
# Round-2 GPU call Z2: A/B of the processing order of the tiles inside a batch: exact cost descending (shipped) against cost classes of
# 2^-bits relative width with Hilbert order inside a class (GIMIC_B200_ORDER_BITS): spatial neighbours gather the same density elements.
set -u
OUT=gpurun_out; mkdir -p $OUT
pr() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
print("   %-12s ms/step %.2f  basis %.2f contract %.2f  TF %.2f frac %.3f" % (sys.argv[2], d["ms_per_step"], s["ms_basis"], s["ms_contract"], r["achieved"], r["frac"]))
PY
}
for v in base 3 1 0 base 3 1 0; do
  if [ $v = base ]; then unset GIMIC_B200_ORDER_BITS; else export GIMIC_B200_ORDER_BITS=$v; fi
  timeout 300 python bench.py --mode octant --steps 5 --warmup 3 > $OUT/r02z2_octant_$v.json 2>/dev/null && pr $OUT/r02z2_octant_$v.json "octant/$v"
done
for v in base 3 0; do
  if [ $v = base ]; then unset GIMIC_B200_ORDER_BITS; else export GIMIC_B200_ORDER_BITS=$v; fi
  timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_jtensor_e -c 2 --csv --log-file $OUT/r02z2_ncu_$v.csv python bench.py --mode octant --steps 1 --warmup 0 --no-extras > /dev/null 2>&1
  python - $OUT/r02z2_ncu_$v.csv $v <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ki=h.index("ID"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
out={}
for r in rows[1:]: out.setdefault(r[ki],{})[r[mi]]=r[vi]+" "+r[ui]
for k,v in out.items(): print("   ncu order", sys.argv[2], "launch", k, {a.split("__")[-1][:28]: b for a,b in v.items()})
PY
done
unset GIMIC_B200_ORDER_BITS
GIMIC_B200_ORDER_BITS=3 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "partition or synthetic_flake or c4h4_read" 2>&1 | tail -1

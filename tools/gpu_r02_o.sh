#!/usr/bin/env bash
# Round-2 GPU call O: where the plane integral's 2.3 ms go (launch list), ncu --set full of k_tile_split on the whole grid.
set -u
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "random_molecules" 2>&1 | tail -1
timeout 120 python tools/plane_probe.py 36 5 2>&1 | tail -1
timeout 120 python tools/plane_probe.py 72 5 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_plane.csv python tools/plane_probe.py 36 1 > /dev/null 2>&1
python - $OUT/${TAG}_launches_plane.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
out=[]
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e3 if u in ("nsecond","ns") else v*1e3 if u in ("msecond","ms") else v
    out.append((r[ki][:70], v))
# the last call of the script (profiled one) = the last 1/3 of the launches after context creation; print the tail
for n,v in out[-30:]: print("   %-72s %9.1f us" % (n,v))
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_tile_split -s 1 -c 1 -o $OUT/${TAG}_tile_split python bench.py --steps 1 --warmup 0 --no-extras > /dev/null 2>&1
ls -la $OUT/${TAG}_tile_split.ncu-rep 2>/dev/null | awk '{print "   ncu-rep bytes", $5}'

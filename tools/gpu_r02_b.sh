#!/usr/bin/env bash
# Round-2 GPU call B: remaining parity tests, then guarded A/B runs of the 16-consumer-warp tensor path and the overlapped basis batches
# (every experimental command under a SHORT timeout: a hang must not eat the budget), launch list and key counters.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu: kernel parity (whole file)"
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > $OUT/${TAG}_pytest_parity.log 2>&1; tail -6 $OUT/${TAG}_pytest_parity.log
run_variant() {   # name, env...
  local name=$1; shift
  echo "== variant $name: quick parity, then subset, then octant step"
  if ( env "$@" timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid" ) > $OUT/${TAG}_pytest_${name}_quick.log 2>&1; then
    tail -2 $OUT/${TAG}_pytest_${name}_quick.log
    ( env "$@" timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "synthetic_flake or open_shell_spin or partition_union or far_origin or uhf_total or switches or hot_path or thin_point or partition_points" ) > $OUT/${TAG}_pytest_${name}.log 2>&1; tail -3 $OUT/${TAG}_pytest_${name}.log
    env "$@" timeout 240 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_octant_${name}.json 2> $OUT/${TAG}_bench_octant_${name}.err
    python - $OUT/${TAG}_bench_octant_${name}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=d["stage_ms_per_step"]
    print("   ms/step %.2f  pts/s %.3fM  contract %.2f ms  basis %.2f  tiles %.2f  TF %.2f frac %.3f" % (d["ms_per_step"], d["value"]/1e6, s["ms_contract"], s["ms_basis"], s["ms_tiles"], r["achieved"], r["frac"]))
except Exception as e:
    print("   no result:", e)
PY
  else
    echo "   quick parity FAILED or timed out:"; tail -5 $OUT/${TAG}_pytest_${name}_quick.log
  fi
}
run_variant ncw8 GIMIC_B200_NCW=8
run_variant ncw16 GIMIC_B200_NCW=16
run_variant ncw16_ovl GIMIC_B200_NCW=16 GIMIC_B200_OVERLAP=1
run_variant ncw8_ovl GIMIC_B200_NCW=8 GIMIC_B200_OVERLAP=1
timeout 200 python tools/legacy_latency.py > $OUT/${TAG}_legacy_latency.json 2>&1; tail -c 400 $OUT/${TAG}_legacy_latency.json
echo "== ncu launch list (octant step, default build) + key counters of k_jtensor (8 and 16 warps) / k_basis / k_tile_split"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1; head -16 $OUT/${TAG}_launches_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:'k_jtensor|k_basis|k_tile_split' -s 8 -c 5 --csv --log-file $OUT/${TAG}_ncu_key.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - $OUT/${TAG}_ncu_key.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1],errors="replace")) if len(r)>5]
hdr=next(r for r in rows if "Kernel Name" in r)
ik,im,iv,iu,ii=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("Metric Unit"),hdr.index("ID")
for r in rows:
    if r is hdr: continue
    print(r[ii], r[ik][:40], r[im][-60:], r[iv], r[iu])
PY
GIMIC_B200_NCW=16 timeout 300 ncu --metrics $M --clock-control none -k regex:'k_jtensor' -s 3 -c 2 --csv --log-file $OUT/${TAG}_ncu_key_ncw16.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - $OUT/${TAG}_ncu_key_ncw16.csv <<'PY'
import csv,sys
try:
    rows=[r for r in csv.reader(open(sys.argv[1],errors="replace")) if len(r)>5]
    hdr=next(r for r in rows if "Kernel Name" in r)
    ik,im,iv,iu,ii=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("Metric Unit"),hdr.index("ID")
    for r in rows:
        if r is hdr: continue
        print(r[ii], r[ik][:40], r[im][-60:], r[iv], r[iu])
except Exception as e:
    print("no ncw16 counters:", e)
PY

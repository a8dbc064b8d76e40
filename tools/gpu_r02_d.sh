#!/usr/bin/env bash
# Round-2 GPU call D: the epilogue-warpgroup kernel (GIMIC_B200_EPI=1) against the round-1 mapping: guarded parity, octant A/B, counters.
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
run_variant() {
  local name=$1; shift
  echo "== variant $name"
  if ( env "$@" timeout 90 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid" ) > $OUT/${TAG}_pytest_${name}_quick.log 2>&1; then
    tail -1 $OUT/${TAG}_pytest_${name}_quick.log
    ( env "$@" timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > $OUT/${TAG}_pytest_${name}.log 2>&1; tail -3 $OUT/${TAG}_pytest_${name}.log
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
    echo "   quick parity FAILED or timed out:"; tail -8 $OUT/${TAG}_pytest_${name}_quick.log
  fi
}
run_variant epi1 GIMIC_B200_EPI=1
run_variant epi0 GIMIC_B200_EPI=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
GIMIC_B200_EPI=1 timeout 300 ncu --metrics $M --clock-control none -k regex:'k_jtensor' -s 3 -c 3 --csv --log-file $OUT/${TAG}_ncu_key_epi1.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - $OUT/${TAG}_ncu_key_epi1.csv <<'PY'
import csv,sys
try:
    rows=[r for r in csv.reader(open(sys.argv[1],errors="replace")) if len(r)>5]
    hdr=next(r for r in rows if "Kernel Name" in r)
    ik,im,iv,iu,ii=hdr.index("Kernel Name"),hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("Metric Unit"),hdr.index("ID")
    for r in rows:
        if r is hdr: continue
        print(r[ii], r[ik][:40], r[im][-60:], r[iv], r[iu])
except Exception as e:
    print("no counters:", e)
PY

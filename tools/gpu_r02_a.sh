#!/usr/bin/env bash
# Round-2 GPU call A: parity gate on the new device-side tile plan / partition / fused field outputs, then the headline bench
# (whole 256^3 grid through the partition API), the round-1 octant mode for comparison, launch list and key ncu counters.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh r02a'
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_nvidia_smi.txt 2>&1
echo "== pytest -m gpu (kernel parity first)"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x ) > $OUT/${TAG}_pytest_parity.log 2>&1; tail -15 $OUT/${TAG}_pytest_parity.log
( time timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_native_driver_gpu.py -q -m gpu ) > $OUT/${TAG}_pytest_drivers.log 2>&1; tail -8 $OUT/${TAG}_pytest_drivers.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1 (headline: whole grid through the partition API)"
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; tail -c 1500 $OUT/${TAG}_bench_n1.json; tail -3 $OUT/${TAG}_bench_n1.err
echo "== bench N=1 octant mode (round-1 workload, for comparison)"
timeout 600 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_n1_octant.json 2> $OUT/${TAG}_bench_n1_octant.err; tail -c 700 $OUT/${TAG}_bench_n1_octant.json
echo "== A/B: tensor path with 16 consumer warps (GIMIC_B200_NCW=16): parity subset, then the octant step"
( GIMIC_B200_NCW=16 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid or synthetic_flake or open_shell_spin or partition_union or far_origin or uhf_total or switches" ) > $OUT/${TAG}_pytest_ncw16.log 2>&1; tail -4 $OUT/${TAG}_pytest_ncw16.log
GIMIC_B200_NCW=16 timeout 600 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_n1_octant_ncw16.json 2> $OUT/${TAG}_bench_n1_octant_ncw16.err; tail -c 700 $OUT/${TAG}_bench_n1_octant_ncw16.json; tail -2 $OUT/${TAG}_bench_n1_octant_ncw16.err
echo "== A/B: + panels of batch b+1 written beside the contraction of batch b (GIMIC_B200_OVERLAP=1)"
( GIMIC_B200_NCW=16 GIMIC_B200_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c4h4_read_grid or synthetic_flake or partition_union or hot_path" ) > $OUT/${TAG}_pytest_ncw16_ovl.log 2>&1; tail -4 $OUT/${TAG}_pytest_ncw16_ovl.log
GIMIC_B200_NCW=16 GIMIC_B200_OVERLAP=1 timeout 600 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_n1_octant_ncw16_ovl.json 2> $OUT/${TAG}_bench_n1_octant_ncw16_ovl.err; tail -c 700 $OUT/${TAG}_bench_n1_octant_ncw16_ovl.json; tail -2 $OUT/${TAG}_bench_n1_octant_ncw16_ovl.err
GIMIC_B200_OVERLAP=1 timeout 600 python bench.py --mode octant --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_n1_octant_ncw8_ovl.json 2> /dev/null; tail -c 500 $OUT/${TAG}_bench_n1_octant_ncw8_ovl.json
timeout 300 python tools/legacy_latency.py > $OUT/${TAG}_legacy_latency.json 2>&1; tail -c 400 $OUT/${TAG}_legacy_latency.json
echo "== ncu launch list (octant step) + key counters of k_jtensor / k_basis / k_tile_split"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1; head -14 $OUT/${TAG}_launches_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_jtensor|k_basis|k_tile_split' -s 12 -c 6 --csv --log-file $OUT/${TAG}_ncu_key.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -60 $OUT/${TAG}_ncu_key.csv | cut -c1-260
GIMIC_B200_NCW=16 timeout 600 ncu --metrics $M --clock-control none -k regex:'k_jtensor' -s 3 -c 3 --csv --log-file $OUT/${TAG}_ncu_key_ncw16.csv \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -30 $OUT/${TAG}_ncu_key_ncw16.csv | cut -c1-260

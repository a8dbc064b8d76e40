#!/usr/bin/env bash
# One gpurun call that re-establishes the measured state of the repo on a fresh B200 (start of a round, or after a kernel change):
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/gpu_round_open.sh r02'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.  Numbers printed by commands that run under
# ncu are never bench values.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { echo "build failed"; tail -20 $OUT/${TAG}_build.log; exit 1; }

echo "== pytest -m gpu"
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

echo "== bench N=1 (headline), ring geometry, general P_b"
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; tail -c 600 $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --steps 5 --warmup 3 --geometry ring --no-cpu-baseline > $OUT/${TAG}_bench_n1_ring.json 2>/dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --general-p --no-cpu-baseline > $OUT/${TAG}_bench_n1_general_p.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>/dev/null

echo "== configs[3] (nbf=1512, ACID + jmod): tensor pass + HBM-bound field pass; drop-in CLI wall clock (Python and native)"
timeout 600 python tools/bench_fields.py 128 > $OUT/${TAG}_bench_fields_config4.json 2>&1
timeout 600 python tools/driver_e2e.py > $OUT/${TAG}_driver_e2e_config4.json 2>&1
timeout 600 python tools/time_driver.py --native > $OUT/${TAG}_driver_small_cases.txt 2>&1
timeout 300 python tools/legacy_latency.py > $OUT/${TAG}_legacy_latency.json 2>&1
timeout 600 python tools/full_grid.py > $OUT/${TAG}_full_grid_256.json 2>&1

echo "== ncu: launch list of the bench command, then ONE full capture of the contraction kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1; head -12 $OUT/${TAG}_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jtensor -s 3 -c 1 -o $OUT/${TAG}_jtensor \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py full $OUT/${TAG}_jtensor.ncu-rep > $OUT/${TAG}_ncu_jtensor.txt 2>&1; head -30 $OUT/${TAG}_ncu_jtensor.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/${TAG}_nvidia_smi.txt 2>&1

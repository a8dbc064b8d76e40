#!/usr/bin/env bash
# Round-2 GPU call Q: one driver (the compiled one behind the Python launcher): whole GPU suite, smoke, driver timings.
set -u
TAG=${1:-r02q}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/${TAG}_pytest_gpu.log | head; grep -E "^E  " $OUT/${TAG}_pytest_gpu.log | head -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/time_driver.py --native 2>&1 | tail -12
timeout 200 python tools/legacy_latency.py 2>&1 | tail -1

#!/usr/bin/env bash
# Round-2 GPU call U: the whole parity file with poisoned workspaces (reads of unwritten memory show as NaN), the spherical test repeated,
# the scan probe (where does integrate_batch spend its time?).
set -u
OUT=gpurun_out; mkdir -p $OUT
( GIMIC_B200_POISON=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > $OUT/r02u_pytest_poison.log 2>&1; grep -E "passed|failed" $OUT/r02u_pytest_poison.log | tail -1; grep -E "^FAILED" $OUT/r02u_pytest_poison.log | head -20
GIMIC_B200_POISON=1 timeout 200 python tools/repro_small_sets.py 20 2>&1 | tail -3
for i in 1 2 3 4 5 6; do ( timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "spherical or random_molecules or more_active" -p no:randomly ) 2>&1 | tail -1; done
timeout 200 python tools/scan_probe.py 2>&1 | tail -8

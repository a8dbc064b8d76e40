#!/usr/bin/env bash
# Round-2 GPU call T: hunt the flaky small-set failure (run-to-run reproducibility with and without column slices, racecheck).
set -u
timeout 200 python tools/repro_small_sets.py 300 2>&1 | tail -7
GIMIC_B200_SLICES=0 timeout 200 python tools/repro_small_sets.py 300 2>&1 | tail -4
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/repro_small_sets.py 2 2>&1 | grep -v "^=========     at\|^=========     by\|^=========         in\|Saved host\|Host Frame" | tail -25

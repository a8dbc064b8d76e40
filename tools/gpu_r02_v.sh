#!/usr/bin/env bash
# Round-2 GPU call V: the spherical parity test's call sequence with poisoned workspaces, and under compute-sanitizer initcheck / memcheck
set -u
GIMIC_B200_POISON=1 timeout 120 python tools/repro_spherical_test.py 2>&1 | tail -5
timeout 120 python tools/repro_spherical_test.py 2>&1 | tail -5
GIMIC_B200_POISON=1 GIMIC_B200_SLICES=0 timeout 120 python tools/repro_spherical_test.py 2>&1 | tail -5
timeout 500 compute-sanitizer --tool initcheck --print-limit 8 python tools/repro_spherical_test.py 2>&1 | grep -v "Host Frame\|Saved host\|^=========         in " | tail -40
timeout 400 compute-sanitizer --tool memcheck --print-limit 8 python tools/repro_spherical_test.py 2>&1 | grep -v "Host Frame\|Saved host\|^=========         in " | tail -25

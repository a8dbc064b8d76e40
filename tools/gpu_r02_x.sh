#!/usr/bin/env bash
# Round-2 GPU call X: creation-time uploads stream-ordered (the flaky beta operand): the whole parity file six times, then the whole GPU suite.
set -u
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2 3 4 5 6; do ( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider ) > $OUT/r02x_parity_$i.log 2>&1; echo "parity run $i: $(tail -1 $OUT/r02x_parity_$i.log)"; grep -E "^E  .*rows off|^FAILED" $OUT/r02x_parity_$i.log | head -3; done
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/r02x_pytest_gpu.log 2>&1; grep -E "passed|failed" $OUT/r02x_pytest_gpu.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/r02x_pytest_gpu.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/scan_bench.py 2>&1 | tail -16 > $OUT/r02x_scan_bench.json; python -c "
import json; d=json.load(open('$OUT/r02x_scan_bench.json')); print({k:(round(v['loop_of_integrate_s']*1e3,2), round(v['integrate_batch_s']*1e3,2)) for k,v in d.items()})"

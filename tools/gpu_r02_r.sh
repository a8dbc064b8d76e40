#!/usr/bin/env bash
# Round-2 GPU call R (2 GPUs): the driver's rank mode over NCCL (torchrun -m gimic_b200 semantics): tools/dist_driver_check.py (cdens open-shell
# 3d files and the c4h4 integral against the reference's goldens on rank 0), the 2-GPU NCCL integral test, the driver GPU tests.
set -u
TAG=${1:-r02r}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 tools/dist_driver_check.py 2>&1 | grep -v "^NCCL\|Warning\|warn" | tail -4 | tee $OUT/${TAG}_dist_driver_check_n2.txt
( timeout 600 python -m pytest tests/test_native_driver_gpu.py tests/test_gpu_driver.py -q -m gpu -k "magnetizability or scan or read_grid or nccl" ) 2>&1 | tail -2
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nccl" ) 2>&1 | tail -1
# torchrun -m gimic_b200 on a reference input, then the same input on one GPU: same report and files
D=$(mktemp -d); python - "$D" <<'PY'
import os, shutil, sys
sys.path.insert(0, "tests"); import fixtures
c = fixtures.materialize(sys.argv[1])
for k in ("a", "b"):
    d = os.path.join(sys.argv[1], k); os.makedirs(d)
    shutil.copy(c["open_shell"]["mol"], d + "/MOL"); shutil.copy(c["open_shell"]["xdens"], d + "/XDENS")
    shutil.copy(os.path.join(fixtures.GOLD, "inputs", "open-shell_3d.inp"), d + "/gimic.inp")
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 -m gimic_b200 $D/a/gimic.inp > $D/a.out 2> $D/a.err; echo "torchrun -m gimic_b200 rc=$?"
timeout 300 python -m gimic_b200 $D/b/gimic.inp > $D/b.out 2> $D/b.err; echo "single rc=$?"
for f in jvec.vti jvecalpha.vti jmodspindens.vti; do cmp $D/a/$f $D/b/$f && echo "   $f identical"; done
grep -c "" $D/a.out $D/b.out; tail -3 $D/a.err

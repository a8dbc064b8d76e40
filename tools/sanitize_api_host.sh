#!/usr/bin/env bash
# The REAL host code of libgimic_b200.so (api.cu: contexts, staging, tile / batch / pool bookkeeping, quadrature and property drivers;
# host_basis.cpp; the nvcc launch stubs) under AddressSanitizer + UBSan + LeakSanitizer in a container without a GPU: linked against the
# FAKE CUDA runtime of tests/fake_cudart/ ("device" memory = zeroed host memory, kernel launches = no-ops reporting success).  Nothing is
# computed -- results are zeros; what runs is every entry point's host-side orchestration (tests/fake_cudart/api_harness.cpp) and every run
# mode of the native driver on top of it.  Test tooling only.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/asan}
FAKE=$OUT/fakecuda
mkdir -p "$OUT" "$FAKE"
cd "$ROOT/gimic_b200/csrc"
make -s all
g++ -O1 -g -std=c++17 -shared -fPIC -I/usr/local/cuda/include -o "$FAKE/libcudart.so.12" "$ROOT/tests/fake_cudart/fake_cudart.cpp"
ln -sf libcudart.so.12 "$FAKE/libcudart.so"
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g -O1"
g++ $SAN -std=c++17 -fPIC -c host_basis.cpp -o "$OUT/host_basis.o"
for f in inp grid writers driver; do g++ $SAN -std=c++17 -fPIC -ffp-contract=off -c driver/$f.cpp -o "$OUT/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 \
    -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer,-g --expt-relaxed-constexpr --extended-lambda \
    -ccbin /usr/bin/g++ -c api.cu -o "$OUT/api.o" 2>/dev/null
KOBJ="k_prepare.o k_jtensor.o k_fields.o"
g++ $SAN -std=c++17 -o "$OUT/api-harness-asan" "$ROOT/tests/fake_cudart/api_harness.cpp" "$OUT/api.o" "$OUT/host_basis.o" $KOBJ -I../../include -L"$FAKE" -lcudart -lpthread
g++ $SAN -std=c++17 -o "$OUT/gimic-b200-fakecuda-asan" driver/main.cpp "$OUT"/{inp,grid,writers,driver,api,host_basis}.o $KOBJ -I../../include -L"$FAKE" -lcudart -lpthread
export LD_LIBRARY_PATH="$FAKE" ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1
PY=$(command -v python3)
TMP=$(mktemp -d)
GIMIC_ROOT="$ROOT" LD_LIBRARY_PATH= "$PY" -c "
import sys, os
sys.path.insert(0, os.path.join('$ROOT', 'tests')); sys.path.insert(0, '$ROOT')
import fixtures
fixtures.materialize('$TMP')
"
"$OUT/api-harness-asan" "$TMP/c4h4/MOL" "$TMP/c4h4/XDENS" "$TMP/open_shell/MOL" "$TMP/open_shell/XDENS"
# once more with a 1 MB panel pool: hundreds of batches per call (the emulated tile counts make the batch / offset bookkeeping non-trivial)
GIMIC_B200_POOL_MB=1 "$OUT/api-harness-asan" "$TMP/c4h4/MOL" "$TMP/c4h4/XDENS" "$TMP/open_shell/MOL" "$TMP/open_shell/XDENS"
# the driver's run modes on the real C ABI host code (the sanitizer-built program is passed for both slots of the runner)
LD_LIBRARY_PATH= "$PY" -c "
import os, subprocess, sys
env = dict(os.environ, LD_LIBRARY_PATH='$FAKE', ASAN_OPTIONS='detect_leaks=1', UBSAN_OPTIONS='print_stacktrace=1')
sys.exit(subprocess.call([sys.executable, '$ROOT/tools/sanitize_host_driver.py', '$OUT/gimic-b200-fakecuda-asan', '$OUT/gimic-b200-fakecuda-asan'], env=env))
"

#!/usr/bin/env bash
# The native driver's complete run modes under AddressSanitizer + UBSan + LeakSanitizer and (for the multi-device thread partition)
# ThreadSanitizer, on CPU: the driver sources are linked with the oracle-backed TEST DOUBLE of the C ABI (tests/mock_backend/), so every
# orchestration path executes (cdens closed/open shell, ACID, property, integrals, scan, appended VTK, --devices).  Test tooling only.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/asan}
mkdir -p "$OUT"
make -s -C "$ROOT/oracle"
cd "$ROOT/gimic_b200/csrc"
SRC="driver/main.cpp driver/inp.cpp driver/grid.cpp driver/writers.cpp driver/driver.cpp host_basis.cpp ../../tests/mock_backend/mock_api.cpp"
LINK="-I../../include -L../../oracle -l:libgimic_oracle.so -Wl,-rpath,$ROOT/oracle -pthread"
g++ -fsanitize=address,undefined -fno-omit-frame-pointer -g -O1 -std=c++17 -o "$OUT/gimic-b200-mock-asan" $SRC $LINK
g++ -fsanitize=thread -fno-omit-frame-pointer -g -O1 -std=c++17 -o "$OUT/gimic-b200-mock-tsan" $SRC $LINK
export ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1 OMP_NUM_THREADS=2
python3 "$ROOT/tools/sanitize_host_driver.py" "$OUT/gimic-b200-mock-asan" "$OUT/gimic-b200-mock-tsan"

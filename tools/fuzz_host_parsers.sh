#!/usr/bin/env bash
# Builds the host side (MOL / XDENS / gimic.inp readers, grids, writers, driver, C ABI) with AddressSanitizer + UBSan into /tmp/asan
# and runs the gimic-b200 program over the reference inputs and over randomly damaged MOL / gimic.inp / grid files.  CPU only:
# everything that would need the device stops at "no CUDA device".  Findings of the round-1 run: a modulo by gauss_order = 0
# (SIGFPE) and a 16 GB allocation for an absurd primitive count -- both now errors (tests/test_native_driver_cpu.py).
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/asan}
mkdir -p "$OUT"
cd "$ROOT/gimic_b200/csrc"
make -s all
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g -O1"
g++ $SAN -std=c++17 -fPIC -c host_basis.cpp -o "$OUT/host_basis.o"
for f in inp grid writers driver; do g++ $SAN -std=c++17 -fPIC -c driver/$f.cpp -o "$OUT/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 \
    -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer,-g --expt-relaxed-constexpr --extended-lambda \
    -ccbin /usr/bin/g++ -c api.cu -o "$OUT/api.o" 2>/dev/null
g++ $SAN -std=c++17 -o "$OUT/gimic-b200-asan" driver/main.cpp "$OUT"/{inp,grid,writers,driver,api,host_basis}.o k_prepare.o k_jtensor.o k_fields.o \
    -I../../include -L/usr/local/cuda/lib64 -lcudart -lpthread
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1
python3 "$ROOT/tools/fuzz_host_parsers.py" "$OUT/gimic-b200-asan" "${2:-600}"

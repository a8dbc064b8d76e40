#!/usr/bin/env bash
# ncu full capture (with source) of k_tile_split on the octant step
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile_split -s 2 -c 1 -o $OUT/${TAG}_tile_split \
    python bench.py --mode octant --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_log.txt 2>&1
tail -2 $OUT/${TAG}_log.txt; ls -la $OUT/${TAG}_tile_split.ncu-rep

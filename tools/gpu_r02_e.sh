#!/usr/bin/env bash
# ncu full capture (with source) of the epilogue-warpgroup kernel on the prof_step chunk
set -u
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
GIMIC_B200_EPI=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jtensor -s 1 -c 1 -o $OUT/${TAG}_jtensor_e python tools/prof_step.py --reps 2 > $OUT/${TAG}_prof_step.log 2>&1
tail -2 $OUT/${TAG}_prof_step.log; ls -la $OUT/${TAG}_jtensor_e.ncu-rep

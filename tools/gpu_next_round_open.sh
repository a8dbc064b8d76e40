#!/usr/bin/env bash
# What round 2 left unmeasured (its GPU budget was spent): run first next round.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_next_round_open.sh'
# 1. N = 8 whole-grid line with drain groups (a rank's range is one panel batch: host outputs are copied out group by group):
#    compare e2e with profiles/r02_bench_n8_grid_h6.json (168.5 M points/s before the groups) and value with 173.1 M.
# 2. the same with GIMIC_B200_DRAIN_BATCHES=0 (no groups) for the A/B.
set -u
bash tools/gpu_r02_h.sh next_n8_groups 8
GIMIC_B200_DRAIN_BATCHES=0 bash tools/gpu_r02_h.sh next_n8_nogroups 8

#!/usr/bin/env bash
# Round-2 GPU call W: bisect the state-dependent failure of test_spherical_basis_vs_oracle[False] (fails inside the whole file, passes alone)
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { local name=$1; shift; ( "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider "${KARGS[@]}" ) > $OUT/r02w_$name.log 2>&1; echo "$name: $(tail -1 $OUT/r02w_$name.log)"; grep -E "^E  .*rows off" $OUT/r02w_$name.log | head -3; }
KARGS=(); run whole_poison env GIMIC_B200_POISON=1
KARGS=(); run whole_plain env
KARGS=(); run whole_poison_noslices env GIMIC_B200_POISON=1 GIMIC_B200_SLICES=0
KARGS=(-k "high_angular or spherical_basis"); run highl_poison env GIMIC_B200_POISON=1
KARGS=(-k "c4h4 or open_shell or spherical_basis"); run first_third_poison env GIMIC_B200_POISON=1
KARGS=(-k "synthetic or linearity or uhf_total or divj or spherical_basis"); run second_third_poison env GIMIC_B200_POISON=1
KARGS=(-k "edge or point_order or legacy or basis_vectors or closed_shell or spherical_basis"); run misc_poison env GIMIC_B200_POISON=1

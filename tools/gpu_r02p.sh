#!/bin/bash
set -u
mkdir -p gpurun_out
CIAOSR_HEAD_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp_pair' -c 1 \
    --profile-from-start off -f -o gpurun_out/r02p_prof_pair python tools/ncu_target.py > gpurun_out/r02p_ncu.log 2>&1
tail -2 gpurun_out/r02p_ncu.log

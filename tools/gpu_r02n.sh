#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/build_timing.sh 2>&1 | tail -2
for pair in 0 1; do
  echo "== CIAOSR_HEAD_PAIR=$pair"
  CIAOSR_HEAD_PAIR=$pair CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_timing.so timeout 300 python tools/wait_breakdown.py 2>&1 | grep -v "rdn\|conv" | head -14
done > gpurun_out/r02n_wait_pair.txt 2>&1
cat gpurun_out/r02n_wait_pair.txt

#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s -k "edsr or cross_scale or csattn or config3" 2>&1 | grep -v "^$" | tail -14 > gpurun_out/r02f_pytest.log
cat gpurun_out/r02f_pytest.log
timeout 300 python tools/time_csattn.py > gpurun_out/r02f_time_csattn.jsonl 2>&1; head -3 gpurun_out/r02f_time_csattn.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --profile-from-start off -f \
   -o gpurun_out/r02f_prof_linear python tools/ncu_linear.py > gpurun_out/r02f_ncu_linear.log 2>&1
tail -2 gpurun_out/r02f_ncu_linear.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attention -c 1 -f \
   -o gpurun_out/r02f_prof_winattn python tools/time_trunk_ops.py > gpurun_out/r02f_ncu_winattn.log 2>&1
tail -2 gpurun_out/r02f_ncu_winattn.log

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02i_pytest.log
cat gpurun_out/r02i_pytest.log
for fused in 1 0; do
  CIAOSR_HEAD_FUSED=$fused timeout 600 python bench.py --steps 10 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02i_bench_fused$fused.json 2> gpurun_out/r02i_bench_fused$fused.err
  tail -3 gpurun_out/r02i_bench_fused$fused.err | cut -c1-200
  python - <<PY
import json
d=json.load(open('gpurun_out/r02i_bench_fused$fused.json'))
print('fused=$fused', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'], d['clocks']['sm_mhz'])
PY
done
bash tools/build_timing.sh 2>&1 | tail -2
CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_timing.so timeout 300 python tools/wait_linear.py > gpurun_out/r02i_wait_linear.txt 2>&1; cat gpurun_out/r02i_wait_linear.txt

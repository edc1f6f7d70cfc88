#!/bin/bash
# ReLU folded into the operand split + tile-boundary software pipeline of the pair kernel: full GPU tests, bench, trace
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | grep -i "max-abs\|passed\|failed\|error\|golden" | tail -40 > gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
for ns in 0 0; do
  timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print(round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done
CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_trace.so timeout 300 python tools/trace_pair.py > gpurun_out/${TAG}_trace.txt 2> gpurun_out/${TAG}_trace.err
python tools/trace_stats.py gpurun_out/${TAG}_trace.txt

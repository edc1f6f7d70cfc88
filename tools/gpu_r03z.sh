#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "cta_pair" 2>&1 | grep -v "^$" | tail -5
for q in 1 0 1 0; do
  CIAOSR_QUERY_PAIR=$q timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r03z_bench.json 2> gpurun_out/r03z_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03z_bench.json'))
    st=d['roofline']['stage_ms_per_step']
    print('query_pair=$q: query %.3f pair %.3f step %.2f parity %.2e' % (st['query_mlp'], st['pair_mlp'], d['ms_per_step'], d['parity']['max_abs_vs_reference_golden']))
except Exception as e: print('ERR', e)
PY
done

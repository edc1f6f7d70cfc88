#!/bin/bash
set -u
mkdir -p gpurun_out
for f in 0 1 0 1; do
  CIAOSR_HEAD_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r04a_bench.json 2> gpurun_out/r04a_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r04a_bench.json'))
    st=d['roofline']['stage_ms_per_step']
    print('fused=$f: pair %.3f query %.3f head-sum %.3f step %.2f parity %.2e' % (st['pair_mlp'], st['query_mlp'], st['pair_mlp']+st['query_mlp'], d['ms_per_step'], d['parity']['max_abs_vs_reference_golden']))
except Exception as e: print('ERR', e)
PY
done

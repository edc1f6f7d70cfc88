#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head or cta_pair" 2>&1 | grep -v "^$" | tail -3
CIAOSR_HEAD_PAIR=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head" 2>&1 | grep -v "^$" | tail -2
for f in 1 0 1; do
  CIAOSR_HEAD_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r04b_bench.json 2> gpurun_out/r04b_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r04b_bench.json'))
    st=d['roofline']['stage_ms_per_step']
    print('fused=$f: pair %.3f query %.3f head-sum %.3f step %.2f parity %.2e' % (st['pair_mlp'], st['query_mlp'], st['pair_mlp']+st['query_mlp'], d['ms_per_step'], d['parity']['max_abs_vs_reference_golden']))
except Exception as e: print('ERR', e)
PY
done
timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head" 2>&1 | grep "ERROR SUMMARY\|passed\|failed" | head -3

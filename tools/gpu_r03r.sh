#!/bin/bash
# opt-in schedules again after the by-value fix
set -u
mkdir -p gpurun_out
for cfg in "2 0" "4 0" "2 1" "4 1" "2 0"; do
  set -- $cfg
  CIAOSR_HEAD_ROWPARTS=$1 CIAOSR_HEAD_NSPLIT=$2 timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r03r_bench.json 2> gpurun_out/r03r_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03r_bench.json'))
    st=d['roofline']['stage_ms_per_step']
    print('parts=$1 nsplit=$2: pair %.3f query %.3f step %.2f clocks %s parity %.2e' % (st['pair_mlp'], st['query_mlp'], d['ms_per_step'], d['clocks']['sm_mhz'], d['parity']['max_abs_vs_reference_golden']))
except Exception as e: print('ERR', e)
PY
done

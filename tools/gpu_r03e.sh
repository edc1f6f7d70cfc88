#!/bin/bash
# is the pair kernel limited by a chip-wide resource (power)?  same work on 148 / 74 / 36 persistent CTAs
set -u
mkdir -p gpurun_out
for sms in 148 74 36 148; do
  CIAOSR_DBG_MAXSMS=$sms timeout 300 python bench.py --steps 10 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r03e_bench_$sms.json 2> gpurun_out/r03e_bench_$sms.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03e_bench_$sms.json'))
    st=d['roofline']['stage_ms_per_step']
    print('CTAs $sms: pair %.3f ms (x%d CTAs = %.1f CTA-ms), query %.3f ms, clocks %s' % (st['pair_mlp'], $sms, st['pair_mlp']*$sms, st['query_mlp'], d['clocks']))
except Exception as e: print('ERR', e)
PY
done

#!/bin/bash
# what bounds convw_tc_kernel?  RDN stage time with parts of the kernel switched off (results are garbage in those runs)
set -u
mkdir -p gpurun_out
for dbg in 0 1 2 4 3 7 0; do
  CIAOSR_DBG_CONV=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r03h_bench_$dbg.json 2> gpurun_out/r03h_bench_$dbg.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03h_bench_$dbg.json'))
    st=d['roofline']['stage_ms_per_step']
    print('dbg $dbg: rdn %.3f ms  (step %.2f, clocks %s)' % (st['rdn_encoder'], d['ms_per_step'], d['clocks']['sm_mhz']))
except Exception as e: print('ERR', e)
PY
done

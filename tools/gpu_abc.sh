#!/bin/bash
# same-box A/B/C of library builds on the bench workload: $1 = tag; libs prev / alt / cur
set -u
TAG=$1
mkdir -p gpurun_out
for rep in 1 2; do for lib in prev alt cur; do
  if [ $lib = cur ]; then unset CIAOSR_LIB; else export CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_$lib.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/${TAG}_bench_$lib.json 2> gpurun_out/${TAG}_bench_$lib.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_$lib.json'))
    print('$lib', round(d['ms_per_step'],2), round(d['value'],2), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done; done

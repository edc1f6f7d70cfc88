#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r02t_pytest.log
cat gpurun_out/r02t_pytest.log
for pair in 1 0; do
  CIAOSR_HEAD_PAIR=$pair timeout 300 python bench.py --steps 20 --warmup 3 --other-configs '' --no-cpu-baseline > gpurun_out/r02t_bench_pair$pair.json 2> gpurun_out/r02t_bench_pair$pair.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02t_bench_pair$pair.json'))
    print('pair=$pair', round(d['ms_per_step'],2), d['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'].get('max_abs_vs_reference_golden'), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
except Exception as e: print('ERR', e)
PY
done

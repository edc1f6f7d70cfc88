#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r02s_pytest.log
cat gpurun_out/r02s_pytest.log
timeout 300 python tools/time_trunk_ops.py > gpurun_out/r02s_trunk_ops.json 2>&1; cat gpurun_out/r02s_trunk_ops.json | tr -d '\n ' ; echo
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
tail -7 gpurun_out/r02s_bench.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
print(round(d['ms_per_step'],2), d['value'], d['e2e']['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'], d['clocks'], round(d['roofline']['frac'],3))
PY

#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02g_pytest.log
cat gpurun_out/r02g_pytest.log
timeout 300 python tools/time_trunk_ops.py > gpurun_out/r02g_trunk_ops.json 2>&1; cat gpurun_out/r02g_trunk_ops.json | tr -d '\n ' ; echo
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -9 gpurun_out/r02g_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'], d['parity'])
PY

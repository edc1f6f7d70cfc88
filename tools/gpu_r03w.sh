#!/bin/bash
# evidence after the by-value fix: tests, bench (both arms), launch list, ncu --set full of the head kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -3 > gpurun_out/r03w_pytest.log
cat gpurun_out/r03w_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r03w_bench.json 2> gpurun_out/r03w_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03w_bench_ref.json 2> gpurun_out/r03w_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    --profile-from-start off python tools/ncu_target.py > gpurun_out/r03w_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp|query_mlp' -c 2 \
    --profile-from-start off -f -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/r03w_ncu_full.log 2>&1
tail -1 gpurun_out/r03w_ncu_full.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03w_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],d['clocks'])
print(d['roofline']['stage_ms_per_step'])
print('parity',d['parity']); print('cpu',d.get('cpu_baseline'))
for o in d.get('other_configs',[]): print({k:o.get(k) for k in ('config','case','ms','mpix_s','error')})
PY

#!/bin/bash
# round-end evidence: GPU tests, smoke(), bench (both arms), launch list, ncu --set full of the head kernels
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -3 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    --profile-from-start off python tools/ncu_target.py > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp|query_mlp' -c 2 \
    --profile-from-start off -f -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],d['clocks'],'launches',d.get('gpu_launches'))
print(d['roofline']['stage_ms_per_step'])
print('parity',d['parity'])
for o in d.get('other_configs',[]): print({k:o.get(k) for k in ('config','case','ms','mpix_s','error')})
PY

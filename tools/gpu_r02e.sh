#!/bin/bash
# Round 2, GPU session E: full GPU suite after the ring / window-attention / EDSR changes, trunk op timings,
# cs-attn timing, configs 3-5.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02e_pytest.log
cat gpurun_out/r02e_pytest.log
timeout 300 python tools/time_trunk_ops.py > gpurun_out/r02e_trunk_ops.json 2>&1; cat gpurun_out/r02e_trunk_ops.json
timeout 300 python tools/time_csattn.py > gpurun_out/r02e_time_csattn.jsonl 2>&1; cat gpurun_out/r02e_time_csattn.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_csattn_launches.csv \
   --profile-from-start off python tools/ncu_csattn.py 64 192 > gpurun_out/r02e_ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/r02e_csattn_launches.csv') if not l.startswith('==')]
for r in csv.DictReader(rows):
    print(r['Kernel Name'][:70], r['Metric Value'], r['Metric Unit'])
PY
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -9 gpurun_out/r02e_bench.err | cut -c1-200; cut -c1-1200 gpurun_out/r02e_bench.json

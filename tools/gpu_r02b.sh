#!/bin/bash
# Round 2, GPU session B: shifted-value cross-scale attention + training path.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s -k "cross_scale or csattn or head_golden or training or train_step or config3 or untiled or random_shapes or edge" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_pytest.log
timeout 300 python tools/time_csattn.py > gpurun_out/r02b_time_csattn.jsonl 2> gpurun_out/r02b_time_csattn.err; cat gpurun_out/r02b_time_csattn.jsonl; tail -3 gpurun_out/r02b_time_csattn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_csattn_launches.csv \
   --profile-from-start off python tools/ncu_csattn.py 64 192 > gpurun_out/r02b_ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/r02b_csattn_launches.csv') if not l.startswith('==')]
for r in csv.DictReader(rows):
    print(r['Kernel Name'][:70], r['Metric Value'], r['Metric Unit'])
PY

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | grep -v "^$" | tail -14 > gpurun_out/r02l_pytest.log
cat gpurun_out/r02l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
tail -8 gpurun_out/r02l_bench.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/r02l_bench.json'))
print(round(d['ms_per_step'],2), d['value'], d['e2e']['value'], {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['parity'], d['clocks'], d['roofline']['frac'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; cut -c1-300 gpurun_out/r02l_bench_ref.json
# ncu: launch list of one step + full capture of the head kernels and the cross-scale attention GEMMs
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    --profile-from-start off python tools/ncu_target.py > gpurun_out/r02l_ncu_list.log 2>&1
tail -2 gpurun_out/r02l_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp|query_mlp' -c 2 \
    --profile-from-start off -f -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/r02l_ncu_full.log 2>&1
tail -2 gpurun_out/r02l_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|softmax_rows|csa_' \
   --profile-from-start off -f -o gpurun_out/r02l_prof_csattn python tools/ncu_csattn.py 64 192 > gpurun_out/r02l_ncu_cs.log 2>&1
tail -2 gpurun_out/r02l_ncu_cs.log

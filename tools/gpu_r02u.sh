#!/bin/bash
# final profiles of the round: launch list, ncu full of the head kernels (pair mode), 2-GPU bench
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    --profile-from-start off python tools/ncu_target.py > gpurun_out/r02u_ncu_list.log 2>&1
tail -1 gpurun_out/r02u_ncu_list.log
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp|query_mlp' -c 2 \
    --profile-from-start off -f -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/r02u_ncu_full.log 2>&1
tail -1 gpurun_out/r02u_ncu_full.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02u_bench_n2.json 2> gpurun_out/r02u_bench_n2.err
tail -8 gpurun_out/r02u_bench_n2.err | cut -c1-200
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02u_bench_n2.json'))
    print('value',d['value'],'strong',d.get('strong'))
    for o in d.get('other_configs',[]): print({k:o.get(k) for k in ('config','case','ms','mpix_s','bit_equal','error')})
    print('parity',d.get('parity'))
except Exception as e: print('ERR',e)
PY

#!/bin/bash
# bench under torchrun on N GPUs of one box: weak scaling line + `strong` + sharded other_configs (bit-equality checked inside)
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r03p_bench_n$N.json 2> gpurun_out/r03p_bench_n$N.err
tail -3 gpurun_out/r03p_bench_n$N.err | cut -c1-200
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r03p_bench_n$N.json'))
    print('N=$N value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
    print('strong',d.get('strong'))
    for o in d.get('other_configs',[]): print({k:o.get(k) for k in ('config','case','ms','mpix_s','bit_equal','sharding','error')})
except Exception as e: print('ERR',e)
PY

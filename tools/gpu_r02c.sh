#!/bin/bash
# Round 2, GPU session C (2 GPUs): bench.py under torchrun (weak + strong legs + other configs), SwinIR trunk profile.
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
tail -15 gpurun_out/r02c_bench_n2.err | cut -c1-300; cut -c1-200 gpurun_out/r02c_bench_n2.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02c_bench_n2.json'))
    print('value',d['value'],'strong',d.get('strong'))
    for o in d.get('other_configs',[]): print(o)
    print('parity',d.get('parity'))
except Exception as e: print('ERR',e)
PY
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/profile_swinir.py 192 2 > gpurun_out/r02c_swinir_profile.txt 2>&1
head -50 gpurun_out/r02c_swinir_profile.txt | cut -c1-230

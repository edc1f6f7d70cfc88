#!/bin/bash
# Round 2, GPU session D: native SwinIR trunk pieces (window attention, LayerNorm, NHWC conv3x3, A-resident Linear).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s -k "native or swinir or full_size or linear" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02d_pytest.log
cat gpurun_out/r02d_pytest.log
timeout 300 python tools/profile_swinir.py 192 2 > gpurun_out/r02d_swinir_profile.txt 2>&1
head -36 gpurun_out/r02d_swinir_profile.txt | cut -c1-70,150-230
timeout 600 python tools/run_configs.py --configs 4,5 --check 1 --stages 1 --out gpurun_out/r02d_configs.jsonl 2>&1 | tail -4 | cut -c1-1200

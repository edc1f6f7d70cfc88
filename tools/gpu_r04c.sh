#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_head or cta_pair or config2_end_to_end" 2>&1 | grep -v "^$" | tail -2
bash tools/gpu_ab.sh $1

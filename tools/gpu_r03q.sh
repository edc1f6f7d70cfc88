#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
bash tools/gpu_ab.sh $1

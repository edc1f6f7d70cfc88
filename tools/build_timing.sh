#!/bin/bash
# diagnostic build of the library with mbarrier wait-time counters (see tools/wait_breakdown.py)
cd "$(dirname "$0")/../ciaosr_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -DCIAOSR_TC_TIMING -o libciaosr_b200_timing.so *.cu

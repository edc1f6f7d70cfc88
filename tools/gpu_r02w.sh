#!/bin/bash
# timeline of CTA 0 of the CTA-pair kernel (trace-only build: no wait counters), 2 and 4 row threads per row
set -u
mkdir -p gpurun_out
for parts in 2 4; do
  CIAOSR_HEAD_ROWPARTS=$parts CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_trace.so timeout 300 python tools/trace_pair.py > gpurun_out/r02w_trace_parts$parts.txt 2> gpurun_out/r02w_trace_parts$parts.err
  tail -3 gpurun_out/r02w_trace_parts$parts.err; wc -l gpurun_out/r02w_trace_parts$parts.txt
done

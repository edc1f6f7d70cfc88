#!/bin/bash
# what slows the UMMAs inside the kernel?  trace build with experiment switches (1 = rows skip arithmetic/stores, 2 = no weight loads)
set -u
mkdir -p gpurun_out
for ns in 0 1; do for flags in 0 1 2 3; do
  CIAOSR_HEAD_NSPLIT=$ns CIAOSR_DBG_FLAGS=$flags CIAOSR_HEAD_ROWPARTS=2 CIAOSR_LIB=ciaosr_b200/csrc/libciaosr_b200_trace.so timeout 300 python tools/trace_pair.py > gpurun_out/r02z_trace_ns${ns}_f$flags.txt 2> gpurun_out/r02z_trace_ns${ns}_f$flags.err
  python tools/trace_stats.py gpurun_out/r02z_trace_ns${ns}_f$flags.txt
done; done

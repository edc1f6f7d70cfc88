#!/bin/bash
# sanitizer sweep over the other native kernels (conv / RDN, cross-scale attention GEMMs, Linear, window attention, LayerNorm)
set -u
mkdir -p gpurun_out
for tool in synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -q \
     -k "native_rdn_encoder or cross_scale_attention_golden or native_linear_matches or native_window_attention or native_layernorm or native_split_activation or native_edsr" > gpurun_out/r03y_$tool.log 2>&1
  echo "== $tool"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r03y_$tool.log | head -4
  grep -A8 "error detected\|Invalid\|Barrier error" gpurun_out/r03y_$tool.log | grep "^=========     at\|Device Frame" | sort | uniq -c | sort -rn | head -6 | cut -c1-220
done

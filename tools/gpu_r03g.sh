#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/precision_frontier.py r03g 2>&1 | tail -5
cp profiles/r03g_precision_frontier.md gpurun_out/

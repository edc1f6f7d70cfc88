#!/bin/bash
# One GPU-box session: parity tests, bench (ours + reference arm), ncu launch list, ncu full capture of the
# top kernels.  Usage: tools/gpu_round.sh [noprof]
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -8 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${1:-}" != "noprof" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      --profile-from-start off python tools/ncu_target.py > gpurun_out/ncu_list.log 2>&1
  tail -3 gpurun_out/ncu_list.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_mlp|query_mlp' -c 2 \
      --profile-from-start off -f -o gpurun_out/prof_head python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json
fi

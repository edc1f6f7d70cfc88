#!/bin/bash
# Round 2, GPU session A: full GPU test-suite on the new BASELINE-size goldens, bench (both arms), sigma table,
# ncu of the cross-scale attention GEMMs on a 192x192 tile.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --durations=8 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02a_pytest.log
tail -40 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -12 gpurun_out/r02a_bench.err; cut -c1-3000 gpurun_out/r02a_bench.json
for t in 16 32; do
  CIAOSR_CPU_THREADS=$t timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref_$t.json 2> gpurun_out/r02a_bench_ref_$t.err
  cut -c1-400 gpurun_out/r02a_bench_ref_$t.json
done
timeout 600 python tools/sigma_table.py > gpurun_out/r02a_sigma.log 2>&1; cat gpurun_out/sigma_table.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm|softmax_rows|csa_' \
   --profile-from-start off -f -o gpurun_out/r02a_prof_csattn python tools/ncu_csattn.py 64 192 > gpurun_out/r02a_ncu_cs.log 2>&1
tail -3 gpurun_out/r02a_ncu_cs.log

#!/usr/bin/env bash
# ncu evidence for one bench step (B=32, 4 new tokens): launch list + full captures of the GEMM in both regimes.
set -u
mkdir -p gpurun_out
CMD="python bench.py --profile-mode --new-tokens 4"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 120 -c 4 -o gpurun_out/prof_gemm_prefill $CMD > gpurun_out/ncu_gemm_prefill.log 2>&1; echo "ncu gemm prefill exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 330 -c 4 -o gpurun_out/prof_gemm_decode $CMD > gpurun_out/ncu_gemm_decode.log 2>&1; echo "ncu gemm decode exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flash_fwd|attn_decode" -s 30 -c 3 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_gpu.log; ls -la gpurun_out

#!/usr/bin/env bash
# GPU visit: all GPU tests (no -x) + ncu launch list of one bench step with the native SAM encoder.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/t_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
CMD="python bench.py --profile-mode --new-tokens 4"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_v2.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/t_gpu.log

#!/usr/bin/env bash
# One GPU visit: tests, smoke, full bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/t_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
tail -25 gpurun_out/t_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 5000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err

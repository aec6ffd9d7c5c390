#!/usr/bin/env bash
# One GPU visit: tests, smoke, short bench, full bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --batch-per-gpu 4 --new-tokens 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench_small exit $?" | tee -a gpurun_out/summary.txt
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err; tail -c 4000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err

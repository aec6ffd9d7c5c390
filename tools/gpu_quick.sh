#!/usr/bin/env bash
# kernel-level GPU tests (no torch-library-heavy model tests) + bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_kernels.log
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
tail -c 2600 gpurun_out/bench_full.json | head -c 1700; tail -3 gpurun_out/bench_full.err

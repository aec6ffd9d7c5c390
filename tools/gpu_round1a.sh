#!/usr/bin/env bash
# first GPU contact: primitive-kernel parity + GEMM micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -x -k "single_tile" -p no:cacheprovider > gpurun_out/t_single_tile.log 2>&1
echo "single_tile exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider --timeout 120 > gpurun_out/t_kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.jsonl 2> gpurun_out/bench_gemm.err
echo "bench_gemm exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/t_single_tile.log; tail -40 gpurun_out/t_kernels.log; cat gpurun_out/bench_gemm.jsonl; cat gpurun_out/summary.txt

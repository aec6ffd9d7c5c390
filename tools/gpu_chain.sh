#!/usr/bin/env bash
# weight-streaming GEMM parity + decode-chain sweep (interleaved prefetch depths)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/t_gemm.log
timeout -s KILL 300 python tools/bench_decode_chain.py 2>&1 | tee gpurun_out/decode_chain.jsonl

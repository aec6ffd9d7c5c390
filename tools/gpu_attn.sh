#!/usr/bin/env bash
# attention-kernel parity + micro-benchmark on the GPU (bounded: a deadlocked kernel is killed with its process)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -q -k "attention" -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/t_attn.log
timeout -s KILL 300 python tools/bench_attn.py 2>&1 | tee gpurun_out/bench_attn.jsonl

#!/usr/bin/env bash
# GEMM parity on the GPU (bounded) + GEMM micro-benchmarks
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" | tee -a gpurun_out/summary.txt
tail -6 gpurun_out/t_gemm.log
timeout -s KILL 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_gemm.jsonl'):
    if not l.startswith('{'): print(l.strip()); continue
    r=json.loads(l); print(r['shape'], r['M'],r['N'],r['K'], 'ours', r['ms'], r['tflops'], 'cublas', r['cublas_ms'], r['cublas_tflops'], r.get('weight_GBps',''), r.get('cublas_weight_GBps',''))
PY

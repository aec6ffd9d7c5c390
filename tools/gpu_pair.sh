#!/usr/bin/env bash
# CTA-pair GEMM: parity (bounded: a deadlocked kernel dies with its process), then micro-benchmark pair vs single
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "cta_pair" -p no:cacheprovider > gpurun_out/t_pair.log 2>&1; rc=$?; echo "pair tests exit $rc" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/t_pair.log
if [ $rc -ne 0 ]; then exit 0; fi
for pair in 1 0; do
  ULLAVA_GEMM_PAIR=$pair SHAPES=b32 NO_CUBLAS=1 timeout -s KILL 200 python tools/bench_gemm.py 2>&1 | P=$pair python -c "
import sys, json, os
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('pair', os.environ['P'], r['shape'], r['ms'], r['tflops'])
"
done | tee gpurun_out/pair.txt

#!/usr/bin/env bash
# rasterisation-group sweep of the large-M GEMM on the bench step's shapes (two passes per setting)
set -u
mkdir -p gpurun_out
for rep in 1 2; do for g in 8 16 32 64; do
  ULLAVA_GROUP_M=$g SHAPES=b32 NO_CUBLAS=1 timeout -s KILL 200 python tools/bench_gemm.py 2>&1 | G=$g python -c "
import sys, json, os
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('g', os.environ['G'], r['shape'], r['ms'], r['tflops'])
"
done; done | tee gpurun_out/groupm.txt

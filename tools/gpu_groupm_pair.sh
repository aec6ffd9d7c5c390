for g in 0 16 64; do
  ULLAVA_GROUP_M=$g SHAPES=b32 NO_CUBLAS=1 timeout -s KILL 100 python tools/bench_gemm.py 2>&1 | G=$g python -c "
import sys, json, os
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l)
        if r['shape'].startswith('llama') or r['shape'].startswith('sam_qkv'): print('g', os.environ['G'], r['shape'], r['ms'], r['tflops'])
"
done

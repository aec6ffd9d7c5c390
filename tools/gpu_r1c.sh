#!/usr/bin/env bash
# GPU visit: weight-streaming GEMM parity (incl. the next-weight L2 prefetch hint), decode-chain sweep over the
# prefetch depth, then the full bench.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "gemm exit $?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/t_gemm.log
timeout -s KILL 240 python tools/bench_decode_chain.py 2>&1 | tee gpurun_out/decode_chain.jsonl
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print({k: r[k] for k in ('value', 'ms_per_step', 'stages_ms', 'clocks')}); print(r['e2e']); print(r['roofline_decode'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/bench_full.err

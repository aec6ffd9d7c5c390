#!/usr/bin/env bash
# decode kernels + model-level parity, then the full bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -x -k "gemm or decode or rope or generate or evaluate or greedy or core" -p no:cacheprovider > gpurun_out/t_dec.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/t_dec.log
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print({k: r[k] for k in ('value', 'ms_per_step', 'stages_ms', 'clocks')}); print(r['e2e'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/bench_full.err

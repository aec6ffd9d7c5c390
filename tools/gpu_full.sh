#!/usr/bin/env bash
# One GPU visit: every GPU test, smoke(), full bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q --durations=6 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/t_gpu.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench_full exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print({k: r[k] for k in ('value', 'ms_per_step', 'stages_ms', 'clocks')}); print(r['e2e'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/bench_full.err

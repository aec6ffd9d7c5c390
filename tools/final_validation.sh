#!/usr/bin/env bash
# One GPU box: the GPU test suite, smoke(), and every bench line quoted in DESIGN.md / profiles/ (JSON lines -> gpurun_out/).
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>/dev/null | tail -1 > gpurun_out/f6_b32.json
python bench.py --config c2 2>/dev/null | tail -1 > gpurun_out/f6_c2.json
python bench.py --config c3 2>/dev/null | tail -1 > gpurun_out/f6_c3.json
python bench.py --config c4 2>/dev/null | tail -1 > gpurun_out/f6_c4.json
python bench.py --no-overlap --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/f6_nooverlap.json
python tools/bench_attn.py 2>/dev/null | grep tcgen05 > gpurun_out/f6_attn.jsonl
for f in gpurun_out/f6_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print(sys.argv[1], round(d["value"], 2), round(d["ms_per_step"], 1), round((d.get("e2e") or {}).get("value", 0), 2),
      round((d.get("roofline") or {}).get("frac", 0), 3), d.get("stages_ms"))
PY
done

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" -p no:cacheprovider > gpurun_out/t_pair.log 2>&1; rc=$?; echo "gemm tests exit $rc"
tail -4 gpurun_out/t_pair.log
if [ $rc -ne 0 ]; then exit 0; fi
python - <<'PY'
import json, os, sys, torch
sys.path.insert(0, "u-llava_b200"); sys.path.insert(0, "tools")
import native
from bench_gemm import timeit
ctx = native.Context.get(0)
dt = torch.bfloat16
shapes = [("vit_qkv", 18464, 3072, 1024, native.EPI_NONE, True, False), ("vit_o", 18464, 1024, 1024, native.EPI_NONE, True, True),
          ("vit_fc1", 18464, 4096, 1024, native.EPI_QUICK_GELU, True, False), ("vit_fc2", 18464, 1024, 4096, native.EPI_NONE, True, True),
          ("sam_qkv", 131072, 3840, 1280, native.EPI_NONE, True, False), ("sam_proj", 131072, 1280, 1280, native.EPI_NONE, True, True),
          ("sam_fc1", 131072, 5120, 1280, native.EPI_GELU, True, False), ("sam_fc2", 131072, 1280, 5120, native.EPI_NONE, True, True),
          ("llama_o", 19456, 4096, 4096, native.EPI_NONE, False, True)]
for name, M, N, K, epi, has_bias, has_res in shapes:
    a = torch.randn((M, K), device="cuda", dtype=dt)
    ws = [torch.randn((N, K), device="cuda", dtype=dt) * K ** -0.5 for _ in range(2)]
    b = torch.randn((N,), device="cuda", dtype=dt) if has_bias else None
    out = torch.randn((M, N), device="cuda", dtype=dt)
    i = [0]
    def run():
        i[0] ^= 1
        ctx.gemm(a, ws[i[0]], bias=b, epilogue=epi, residual=out if has_res else None, out=out)
    t = timeit(run)
    print(name, "bias" if has_bias else "", "res" if has_res else "", round(t, 4), "ms", round(2.0 * M * N * K / t / 1e9, 1), "TFLOP/s", flush=True)
PY

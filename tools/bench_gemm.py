"""Micro-benchmark of the tcgen05 GEMM on the hot path's shapes (CUDA events, L2 flushed between
iterations by cycling through operand copies larger than L2).  Prints one JSON line per shape with
achieved TFLOP/s (or GB/s for the weight-streaming decode shapes) and the cuBLAS number beside it
(cuBLAS is the measured roofline denominator, not part of the product)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "u-llava_b200"))
import native  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


def main():
    ctx = native.Context.get(0)
    dt = torch.bfloat16
    shapes = [
        # (name, M, N, K, epilogue)
        ("vit_qkv_b32", 18464, 3072, 1024, native.EPI_NONE),
        ("vit_fc1_b32", 18464, 4096, 1024, native.EPI_QUICK_GELU),
        ("vit_fc2_b32", 18464, 1024, 4096, native.EPI_NONE),
        ("llama_qkv_b16", 9728, 12288, 4096, native.EPI_NONE),
        ("llama_o_b16", 9728, 4096, 4096, native.EPI_NONE),
        ("llama_gateup_b16", 9728, 22016, 4096, native.EPI_SILU_MUL),
        ("llama_down_b16", 9728, 4096, 11008, native.EPI_NONE),
        ("llama_qkv_b32", 19456, 12288, 4096, native.EPI_NONE),
        ("square_8192", 8192, 8192, 8192, native.EPI_NONE),
        ("sam_qkv_b8", 32768, 3840, 1280, native.EPI_NONE),
        ("sam_proj_b8", 32768, 1280, 1280, native.EPI_NONE),
        ("sam_fc1_b8", 32768, 5120, 1280, native.EPI_GELU),
        ("sam_fc2_b8", 32768, 1280, 5120, native.EPI_NONE),
        ("decode_qkv_b8", 8, 12288, 4096, native.EPI_NONE),
        ("decode_gateup_b8", 8, 22016, 4096, native.EPI_SILU_MUL),
        ("decode_down_b8", 8, 4096, 11008, native.EPI_NONE),
        ("decode_qkv_b32", 32, 12288, 4096, native.EPI_NONE),
        ("decode_down_b32", 32, 4096, 11008, native.EPI_NONE),
    ]
    if os.environ.get("SHAPES") == "b32":   # the large-M GEMMs of the bench step (32 images per GPU)
        shapes = [("vit_qkv_b32", 18464, 3072, 1024, native.EPI_NONE), ("vit_o_b32", 18464, 1024, 1024, native.EPI_NONE),
                  ("vit_fc1_b32", 18464, 4096, 1024, native.EPI_QUICK_GELU), ("vit_fc2_b32", 18464, 1024, 4096, native.EPI_NONE),
                  ("llama_qkv_b32", 19456, 12288, 4096, native.EPI_NONE), ("llama_o_b32", 19456, 4096, 4096, native.EPI_NONE),
                  ("llama_gateup_b32", 19456, 22016, 4096, native.EPI_SILU_MUL),
                  ("llama_down_b32", 19456, 4096, 11008, native.EPI_NONE),
                  ("sam_qkv_b32", 131072, 3840, 1280, native.EPI_NONE), ("sam_proj_b32", 131072, 1280, 1280, native.EPI_NONE),
                  ("sam_fc1_b32", 131072, 5120, 1280, native.EPI_GELU), ("sam_fc2_b32", 131072, 1280, 5120, native.EPI_NONE)]
    if os.environ.get("SHAPES") == "large":
        shapes = [s for s in shapes if s[1] > 32]
    if os.environ.get("EPI_NONE"):   # same shapes without their activation: what the epilogue costs
        shapes = [(n, M, N, K, native.EPI_NONE) for n, M, N, K, e in shapes if e != native.EPI_SILU_MUL]
    for name, M, N, K, epi in shapes:
        ncopies = max(2, int(160e6 // (N * K * 2)) + 1) if M <= 32 else 2
        a = torch.randn((M, K), device="cuda", dtype=dt)
        ws = [torch.randn((N, K), device="cuda", dtype=dt) * K ** -0.5 for _ in range(ncopies)]
        out = torch.empty((M, N // 2 if epi == native.EPI_SILU_MUL else N), device="cuda", dtype=dt)
        i = [0]

        def ours():
            i[0] = (i[0] + 1) % ncopies
            ctx.gemm(a, ws[i[0]], epilogue=epi, out=out)

        def cublas():
            i[0] = (i[0] + 1) % ncopies
            torch.matmul(a, ws[i[0]].t())

        t = timeit(ours)
        tc = timeit(cublas) if not os.environ.get("NO_CUBLAS") else float("nan")
        flops = 2.0 * M * N * K
        rec = {"shape": name, "M": M, "N": N, "K": K, "ms": round(t, 4), "tflops": round(flops / t / 1e9, 1),
               "cublas_ms": round(tc, 4), "cublas_tflops": round(flops / tc / 1e9, 1)}
        if M <= 32:
            rec["weight_GBps"] = round(N * K * 2 / t / 1e6, 1)
            rec["cublas_weight_GBps"] = round(N * K * 2 / tc / 1e6, 1)
        print(json.dumps(rec), flush=True)
        del a, ws, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

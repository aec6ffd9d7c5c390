"""Micro-benchmark of the single-query decode attention (LLaMA-7B geometry): B = 32 samples x 32 heads, hd 128,
context 640 of a 672-position KV cache, 32 distinct layer caches (10.7 GB, nothing L2 resident), captured in one CUDA
graph.  Prints us per layer and achieved KV GB/s against the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "u-llava_b200"))
import native  # noqa: E402


def main():
    ctx = native.Context.get(0)
    dt = torch.bfloat16
    B, H, D, L = int(os.environ.get("B", 32)), 32, 128, 32
    max_seq, ctx_len = 672, int(os.environ.get("CTX", 640))
    kc = torch.randn((L, B, H, max_seq, D), device="cuda", dtype=dt)
    vc = torch.randn((L, B, H, max_seq, D), device="cuda", dtype=dt)
    q = torch.randn((B, 3 * H * D), device="cuda", dtype=dt)
    out = torch.empty((B, H * D), device="cuda", dtype=dt)
    peak = 6554.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]

    def one_pass():
        for l in range(L):
            ctx.attention_decode(q[:, :H * D], kc[l], vc[l], ctx_len, out=out)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        one_pass()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            one_pass()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    kv_bytes = L * B * H * ctx_len * D * 2 * 2
    print(json.dumps({"kernel": "attn_decode", "B": B, "ctx": ctx_len, "us_per_layer": round(1e3 * ms / L, 1),
                      "kv_GBps": round(kv_bytes / ms / 1e6, 1), "frac_of_hbm_peak": round(kv_bytes / ms / 1e6 / peak, 3)}),
          flush=True)


if __name__ == "__main__":
    main()

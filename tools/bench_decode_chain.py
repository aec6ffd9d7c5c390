"""Micro-benchmark of the decode step's GEMM chain (LLaMA-7B geometry, B = 32): per layer
rmsnorm -> qkv -> o(+res) -> rmsnorm -> gate/up (SiLU*mul) -> down(+res), 32 layers with distinct weights
(12.9 GB streamed per pass, so nothing is L2 resident), captured in one CUDA graph and replayed.
Prints ms per pass and achieved weight GB/s, with programmatic dependent launch on and off."""
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
    B, H, F, L = int(os.environ.get("B", 32)), 4096, 11008, 32
    dev = "cuda"
    mk = lambda *s: (torch.randn(s, device=dev, dtype=torch.float32) * s[-1] ** -0.5).to(dt)
    layers = [dict(g1=torch.ones(H, device=dev, dtype=dt), wqkv=mk(3 * H, H), wo=mk(H, H),
                   g2=torch.ones(H, device=dev, dtype=dt), wgu=mk(2 * F, H), wd=mk(H, F)) for _ in range(L)]
    hid = torch.randn((B, H), device=dev, dtype=torch.float32).to(dt)
    xn = torch.empty_like(hid)
    qkv = torch.empty((B, 3 * H), device=dev, dtype=dt)
    act = torch.empty((B, F), device=dev, dtype=dt)
    wbytes = sum(sum(t.numel() * 2 for k, t in l.items() if k.startswith("w")) for l in layers)

    def one_pass():
        # next-weight hints exactly as csrc/models.cu sets them inside the decode step (no hint across attention,
        # which sits between qkv and o in the real step; here o follows qkv directly, so hint it as well)
        for i, l in enumerate(layers):
            ctx.rmsnorm(hid, l["g1"], 1e-6, out=xn)
            ctx.gemm_next_weight(l["wo"])
            ctx.gemm(xn, l["wqkv"], out=qkv)
            ctx.gemm_next_weight(l["wgu"])
            ctx.gemm(qkv[:, :H], l["wo"], residual=hid, out=hid)
            ctx.rmsnorm(hid, l["g2"], 1e-6, out=xn)
            ctx.gemm_next_weight(l["wd"])
            ctx.gemm(xn, l["wgu"], epilogue=native.EPI_SILU_MUL, out=act)
            ctx.gemm_next_weight(layers[(i + 1) % L]["wqkv"])
            ctx.gemm(act, l["wd"], residual=hid, out=hid)

    if os.environ.get("NO_GRAPH"):   # for ncu: plain launches, two passes
        ctx.set_pdl(os.environ.get("PDL", "1") == "1")
        one_pass()
        one_pass()
        torch.cuda.synchronize()
        return
    sweep = [(True, int(u)) for u in os.environ.get("PF_SWEEP", "0,12,0,12,24,0,24,48,0,48").split(",")] + [(False, 0)]
    for pdl, pf in sweep:
        ctx.set_pdl(pdl)
        ctx.set_weight_prefetch(pf)
        hid.normal_()
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
        print(json.dumps({"chain": "llama7b_decode_gemms", "B": B, "pdl": pdl, "prefetch_tiles_per_sm": pf, "ms_per_pass": round(ms, 3),
                          "us_per_layer": round(1e3 * ms / L, 1), "weight_GBps": round(wbytes / ms / 1e6, 1),
                          "frac_of_6554": round(wbytes / ms / 1e6 / 6554.2, 3)}), flush=True)
    ctx.set_pdl(True)
    ctx.set_weight_prefetch(12)


if __name__ == "__main__":
    main()

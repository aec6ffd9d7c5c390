"""Micro-benchmark of the attention kernels on the hot path's shapes (CUDA events, median of 10).
One JSON line per (shape, impl): impl 0 = tcgen05/TMEM kernel, 1 = warp-level mma.sync kernel."""
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
    only = sys.argv[1:] or None
    shapes = [
        # name, B, H, S, D, causal, relpos grid side
        ("clip_b32", 32, 16, 577, 64, False, 0),
        ("llama_prefill_b32", 32, 32, 608, 128, True, 0),
        ("sam_window_b32", 32 * 25, 16, 196, 80, False, 14),
        ("sam_global_b8", 8, 16, 4096, 80, False, 64),
        ("plain_4096_hd128_b4", 4, 16, 4096, 128, False, 0),
        ("plain_4096_hd64_b4", 4, 16, 4096, 64, False, 0),
    ]
    for name, B, H, S, D, causal, G in shapes:
        if only and name not in only:
            continue
        qkv = torch.randn((B, S, 3, H, D), device="cuda", dtype=torch.float32).to(dt)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        out = torch.empty((B, S, H, D), device="cuda", dtype=dt)
        flops = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        if G:
            both = (torch.randn((2 * (2 * G - 1), D), device="cuda") * 0.3).to(dt)
            rh, rw = both[:2 * G - 1], both[2 * G - 1:]
            fn = lambda: ctx.attention_relpos(q, k, v, rh, rw, G, out=out)
        else:
            fn = lambda: ctx.attention(q, k, v, causal=causal, out=out)
        if os.environ.get("TRACE"):
            # per-tile timeline of CTA (0, 0, 0): cycles relative to the first stamp
            ctx.set_attention_impl(2)
            fn()
            buf = torch.zeros((64, 48), dtype=torch.int64, device="cuda")
            ctx.lib.ullava_debug_fmha_trace(ctx.handle, buf.data_ptr())
            fn()
            torch.cuda.synchronize()
            ctx.lib.ullava_debug_fmha_trace(ctx.handle, None)
            t = buf.cpu().numpy()
            t0 = t[t > 0].min()
            rel = lambda v: [int(x - t0) if x > 0 else None for x in v]
            if name.startswith("sam_window"):
                print(name, "per tile: [loads issued, Q/K landed, QK issued, S landed, tables done, pass 1 done, P arrived, "
                      "P landed, PV issued, O landed, stored]")
                for tt in range(2):
                    print(tt, rel(t.reshape(-1)[tt * 16: tt * 16 + 11]))
                ctx.set_attention_impl(0)
                continue
            print(name, "per tile: [K issued, K landed, QK issued, V landed, P landed, PV issued]; then per softmax warp 0..7")
            for j in range(4, min(14, (S + 127) // 128)):
                print(j, rel(t[j][[0, 1, 2, 3, 4, 14]]))
                print("    S landed  ", rel(t[j][32:40]))
                print("    max done  ", rel(t[j][24:32]))
                print("    handed on ", rel(t[j][40:48]))
                print("    P arrived ", rel(t[j][16:24]))
            ctx.set_attention_impl(0)
            continue
        for impl in (2, 1):
            ctx.set_attention_impl(impl)
            ms = timeit(fn)
            print(json.dumps({"shape": name, "impl": "tcgen05" if impl == 2 else "mma.sync", "B": B, "H": H, "S": S,
                              "D": D, "causal": causal, "relpos": G, "ms": round(ms, 4),
                              "tflops": round(flops / ms * 1e-9, 1)}), flush=True)
        ctx.set_attention_impl(0)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Decode step of a 32-layer LLaMA-7B-width stack (B = 32, 640-position KV cache) through the CUDA graph: one kernel per
GEMM / norm vs the decode-layer chain kernel, sweep of the chain's L2 look-ahead (ULLAVA_PREFETCH_UNITS is read at context
creation, hence one process per value).   python tools/bench_chain.py           (GPU box)
NCU=1: eager launches only (for ncu -k regex:gemm_chain)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-llava_b200")]


def child():
    import torch
    import native
    from transformers import LlamaConfig, LlamaModel
    from models.engine import LlamaStack
    ctx = native.Context.get(0)
    dtype = torch.bfloat16
    layers, batch, P, new = int(os.environ.get("LAYERS", 32)), int(os.environ.get("B", 32)), 608, 40
    torch.manual_seed(0)
    lc = LlamaConfig(vocab_size=32011, hidden_size=4096, intermediate_size=11008, num_hidden_layers=layers,
                     num_attention_heads=32, num_key_value_heads=32, rms_norm_eps=1e-6)
    with torch.device("cuda"):
        mod = LlamaModel(lc).eval().to(dtype)
        head = torch.nn.Linear(4096, 32011, bias=False).to(dtype)
    stack = LlamaStack(mod, head)
    stack.ensure()
    ids = torch.randint(3, 32000, (batch, P), device="cuda")
    out = {"prefetch_units": os.environ.get("ULLAVA_PREFETCH_UNITS", "default"), "layers": layers, "batch": batch}
    for use_chain in (False, True):
        stack._session = None
        sess = stack.decode_session(ctx, batch, P + new + 2, stack.embed_w, stack.head_w, False)
        sess.use_chain = use_chain
        sess.begin(ids, None, 0)
        x = ctx.embed_gather(ids, stack.embed_w)
        final, _ = stack.run(ctx, x, sess.cache, batch, P)
        sess.first_token(final.view(batch, P, 4096)[:, -1].contiguous(), P)
        if os.environ.get("NCU"):
            sess.steps(2, use_graph=False)
            torch.cuda.synchronize()
            continue
        if os.environ.get("TRACE") and use_chain:
            # timeline of the chain kernel of the LAST layer-chain launched in one eager step (all chains write to the
            # same buffer, the last one wins): per step, medians over CTAs relative to the kernel's first stamp
            sess.steps(2, use_graph=False)
            G = ctx.sm_count()
            buf = torch.zeros((G, 4, 8), dtype=torch.int64, device="cuda")
            ctx.lib.ullava_debug_chain_trace(ctx.handle, buf.data_ptr())
            sess.steps(1, use_graph=False)
            torch.cuda.synchronize()
            ctx.lib.ullava_debug_chain_trace(ctx.handle, None)
            t = buf.cpu().numpy().astype("float64")
            t0 = t[t > 0].min()
            names = ["W issued", "X ready", "first MMA", "last MMA", "step done", "W first", "norm begin", "norm done"]
            for s_ in range(4):
                row = {}
                for k, nm in enumerate(names):
                    v = t[:, s_, k]
                    v = v[v > 0]
                    if len(v):
                        row[nm] = [round((v.min() - t0) / 1e3, 2), round((sorted(v)[len(v) // 2] - t0) / 1e3, 2),
                                   round((v.max() - t0) / 1e3, 2)]
                print("step", s_, json.dumps(row), flush=True)
            continue
        sess.steps(4, use_graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sess.steps(new - 4, use_graph=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (new - 4)
        out["chain_ms_per_step" if use_chain else "per_gemm_ms_per_step"] = round(ms, 4)
        out["chain_us_per_layer" if use_chain else "per_gemm_us_per_layer"] = round(1e3 * ms / layers, 2)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
    else:
        for pf in os.environ.get("PF_SWEEP", "12,0,4,8").split(","):
            subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, ULLAVA_PREFETCH_UNITS=pf), check=False)

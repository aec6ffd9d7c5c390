#!/usr/bin/env python
"""One launch (after a warm-up) of each large-M GEMM of the bench step, for DRAM-traffic capture under ncu:
  ULLAVA_GROUP_M=<g> ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      -k regex:gemm_tcgen05 --csv python tools/gemm_traffic.py
Prints the shapes in launch order (2 launches each: the second is the one to read)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "u-llava_b200"))
import native  # noqa: E402

SHAPES = [("llama_qkv_b32", 19456, 12288, 4096, native.EPI_NONE), ("llama_o_b32", 19456, 4096, 4096, native.EPI_NONE),
          ("llama_gateup_b32", 19456, 22016, 4096, native.EPI_SILU_MUL), ("llama_down_b32", 19456, 4096, 11008, native.EPI_NONE),
          ("sam_qkv_b32", 131072, 3840, 1280, native.EPI_NONE), ("sam_fc1_b32", 131072, 5120, 1280, native.EPI_GELU),
          ("sam_fc2_b32", 131072, 1280, 5120, native.EPI_NONE), ("vit_fc1_b32", 18464, 4096, 1024, native.EPI_QUICK_GELU)]


def main():
    ctx = native.Context.get(0)
    dt = torch.bfloat16
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name, M, N, K, epi in SHAPES:
        a = torch.randn((M, K), device="cuda", dtype=dt)
        w = torch.randn((N, K), device="cuda", dtype=dt) * K ** -0.5
        out = torch.empty((M, N // 2 if epi == native.EPI_SILU_MUL else N), device="cuda", dtype=dt)
        for _ in range(2):
            flush.zero_()
            ctx.gemm(a, w, epilogue=epi, out=out)
        torch.cuda.synchronize()
        print(json.dumps({"shape": name, "operand_MB": round((M * K + N * K) * 2 / 1e6, 1),
                          "output_MB": round(out.numel() * 2 / 1e6, 1)}), flush=True)
        del a, w, out


if __name__ == "__main__":
    main()

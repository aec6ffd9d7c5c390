"""ORACLE SUPPORT -- TEST INFRASTRUCTURE ONLY.

Deterministic synthetic weights and inputs, keyed by state_dict name, so that the real reference
(golden generation, build container only), the oracle and the B200 implementation can all be loaded
with identical tensors without shipping weight files: value(key, shape, seed) is a pure function.
All values are rounded to bf16-representable numbers, hence exactly loadable into bf16 models and
(up to fp16 subnormals) fp16 models.
"""
from __future__ import annotations

import zlib
from typing import Dict, Sequence, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def synth_tensor(key: str, shape: Sequence[int], seed: int = 0) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    g = _gen(key, seed)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = key.rsplit(".", 1)[-1]
    if "gaussian_matrix" in key:
        v = r
    elif len(shape) == 1:
        if leaf == "bias":
            v = 0.05 * r
        elif leaf == "weight":            # LayerNorm / RMSNorm / LayerNorm2d scales
            v = 1.0 + 0.1 * r
        else:                             # e.g. CLIP class_embedding
            v = 0.5 * r
    elif any(t in key for t in ("embed_tokens", "position_embedding", "pos_embed", "rel_pos", "iou_token",
                                "mask_tokens", "no_mask_embed", "point_embeddings", "not_a_point_embed")):
        v = 0.5 * r
    elif "output_upscaling" in key and len(shape) == 4:   # ConvTranspose2d [ci, co, kh, kw]
        v = r / (shape[0] ** 0.5)
    elif len(shape) == 4:                                 # Conv2d [co, ci, kh, kw]
        v = r / ((shape[1] * shape[2] * shape[3]) ** 0.5)
    elif len(shape) == 2:                                 # Linear [out, in]
        v = r / (shape[1] ** 0.5)
    else:
        v = 0.5 * r
    return v.to(torch.bfloat16).to(torch.float32)


def synth_state_dict(shapes: Dict[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_normal(name: str, shape: Sequence[int], seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """Seeded N(0, scale^2) input tensor (images etc.), bf16-representable."""
    r = torch.randn(tuple(shape), generator=_gen("input:" + name, seed), dtype=torch.float32) * scale
    return r.to(torch.bfloat16).to(torch.float32)


def synth_ids(name: str, shape: Sequence[int], low: int, high: int, seed: int = 0) -> torch.Tensor:
    return torch.randint(low, high, tuple(shape), generator=_gen("ids:" + name, seed), dtype=torch.int64)


def subsample(t: torch.Tensor, max_elems: int = 16384) -> Tuple[torch.Tensor, int]:
    """Strided subsample of a flattened tensor for compact golden fixtures: returns (values, stride)."""
    flat = t.reshape(-1)
    stride = max(1, (flat.numel() + max_elems - 1) // max_elems)
    return flat[::stride].clone(), stride

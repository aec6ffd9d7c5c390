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


# --------------------------------------------------------------------------------------------------------------------
# "Coherent mask" variant of the synthetic fixtures.
#
# With i.i.d. random weights and N(0,1) pixels the SAM mask logits are noise-like: every pixel is an independent draw
# around 0, a fixed share of them lies within rounding of the threshold, and a 16-bit implementation cannot reach
# north_star's IoU >= 0.999 against an fp32 reference on such a mask (nor could the reference's own 16-bit path).  A
# trained SAM produces COHERENT masks: large |logit| inside / outside an object and a thin boundary.  The overrides below
# give the synthetic model that property without any checkpoint, by removing every source of per-position variation
# that random weights turn into noise:
#   * pos_embed = 0 and rel_pos_h / rel_pos_w = 0 -- the reference's own initialisation (image_encoder.py:73-77,
#     216-219) -- and a piecewise-constant input image whose regions are unions of whole 14 x 14-token attention
#     windows: all tokens of a region then see the same patch AND the same window content, so the image embedding is
#     piecewise constant (a few classes: region, background, the padded windows at the right / bottom edge, one-token
#     borders from the 3 x 3 neck convolution);
#   * the random Fourier matrix of the dense positional encoding = 0 (prompt_encoder.py:195-201): the mask decoder's
#     image <-> token attention then depends on a token's content only;
#   * the four (dy, dx) sub-kernels of both ConvTranspose2d layers of output_upscaling are tied (mask_decoder.py:52-60):
#     an up-scaled pixel then depends on its token only, not on its position inside the token's 4 x 4 block (random
#     sub-kernels print a fixed sign checkerboard on every token).
# Everything else (attention, MLPs, norms, hyper-network, the LLM and the [SEG] prompt path) keeps its seeded random
# weights and runs the same kernels as always.
# --------------------------------------------------------------------------------------------------------------------
def coherent_overrides(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = dict(sd)
    for k, v in sd.items():
        if k.endswith("image_encoder.pos_embed") or "rel_pos" in k or "positional_encoding_gaussian_matrix" in k:
            out[k] = torch.zeros_like(v)
        elif "output_upscaling" in k and v.dim() == 4:          # ConvTranspose2d weight [ci, co, 2, 2]
            out[k] = v[:, :, :1, :1].expand_as(v).contiguous()
    return out


# Seeds of the coherent fixture (search: tests/golden/find_seed.py coherent): SAM image seed 96 and CLIP-image seed 2105
# are a pair for which all three [SEG] masks of the tiny_full prompts are two-signed (19 - 25 % foreground) with < 6e-4 of
# the pixels within 3 % of the largest |logit| -- i.e. only the interpolated boundary itself is near the threshold, no
# whole class of pixels sits there.
COHERENT_SEED = 96
COHERENT_CLIP_SEED = 2105


def coherent_clip_images(batch: int) -> torch.Tensor:
    """the CLIP-side images of the coherent fixture ([batch, 3, 28, 28] for the tiny vision tower)"""
    return synth_normal("images", (batch, 3, 28, 28), seed=COHERENT_CLIP_SEED)


def coherent_image(batch: int, size: int = 1024, seed: int = COHERENT_SEED, window_px: int = 224) -> torch.Tensor:
    """[batch, 3, size, size]: per sample a 2 x 2-window square of one constant colour on a background of another
    (bf16-representable values around +-1.5); position and colours are seeded per sample."""
    g = _gen("coherent:x", seed)
    out = []
    for _ in range(batch):
        fg = 1.5 * torch.randn(3, generator=g)
        bg = 1.5 * torch.randn(3, generator=g)
        im = bg[:, None, None].expand(3, size, size).clone()
        y0 = window_px * int(torch.randint(0, 2, (1,), generator=g))
        x0 = window_px * int(torch.randint(0, 2, (1,), generator=g))
        im[:, y0:y0 + 2 * window_px, x0:x0 + 2 * window_px] = fg[:, None, None]
        out.append(im)
    return torch.stack(out).to(torch.bfloat16).to(torch.float32)

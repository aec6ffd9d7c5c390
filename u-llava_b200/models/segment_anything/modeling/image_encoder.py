"""SAM ViT image encoder (reference: segment_anything/modeling/image_encoder.py:16-426).

Scope note: the encoder is ON the u-LLaVA path but is not one of
the kernels the north star names; in this round it runs as batched PyTorch-on-GPU library ops (cuBLAS
GEMMs + SDPA with the decomposed relative-position bias passed as an additive mask) and is the first
"next" row to be replaced by sm_100a kernels.  Same parameters / state_dict keys as the reference, so
SAM checkpoints load unchanged.  Differences from the reference implementation that do not change the
math: images are processed as one batch (the reference loops image by image with empty_cache() calls),
windows are attended through scaled_dot_product_attention instead of a materialised softmax."""
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn
import torch.nn.functional as F

from .common import LayerNorm2d, MLPBlock


class PatchEmbed(nn.Module):
    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)  # B C H W -> B H W C


def _rel_table(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """[q_size, k_size, hd] lookup of the relative-position embedding (linear resize when the stored table has a
    different length, as the reference does)."""
    max_rel = 2 * max(q_size, k_size) - 1
    if rel_pos.shape[0] != max_rel:
        rp = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1).float(), size=max_rel,
                           mode="linear").reshape(-1, max_rel).permute(1, 0).to(rel_pos.dtype)
    else:
        rp = rel_pos
    dev = rel_pos.device
    qc = torch.arange(q_size, device=dev)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size, device=dev)[None, :] * max(q_size / k_size, 1.0)
    idx = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rp[idx.long()]


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True, use_rel_pos=False, rel_pos_zero_init=True,
                 input_size: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, self.head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, self.head_dim))

    def forward(self, x):
        B, H, W, _ = x.shape
        nh, hd = self.num_heads, self.head_dim
        qkv = self.qkv(x).reshape(B, H * W, 3, nh, hd).permute(2, 0, 3, 1, 4)  # 3 B nh HW hd
        q, k, v = qkv[0], qkv[1], qkv[2]
        bias = None
        if self.use_rel_pos:
            rq = q.reshape(B, nh, H, W, hd)
            rel_h = torch.einsum("bnhwc,hkc->bnhwk", rq, _rel_table(H, H, self.rel_pos_h))
            rel_w = torch.einsum("bnhwc,wkc->bnhwk", rq, _rel_table(W, W, self.rel_pos_w))
            bias = (rel_h[..., :, None] + rel_w[..., None, :]).reshape(B, nh, H * W, H * W)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=self.scale)
        o = o.permute(0, 2, 1, 3).reshape(B, H, W, nh * hd)
        return self.proj(o)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, norm_layer: Type[nn.Module] = nn.LayerNorm,
                 act_layer: Type[nn.Module] = nn.GELU, use_rel_pos=False, rel_pos_zero_init=True, window_size=0,
                 input_size: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, qkv_bias, use_rel_pos, rel_pos_zero_init,
                              input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio), act_layer)
        self.window_size = window_size

    def forward(self, x):
        shortcut = x
        x = self.norm1(x)
        ws = self.window_size
        if ws > 0:
            B, H, W, C = x.shape
            ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
            if ph or pw:
                x = F.pad(x, (0, 0, 0, pw, 0, ph))
            Hp, Wp = H + ph, W + pw
            x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
        x = self.attn(x)
        if ws > 0:
            x = x.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
            if ph or pw:
                x = x[:, :H, :W].contiguous()
        x = shortcut + x
        return x + self.mlp(self.norm2(x))


class ImageEncoderViT(nn.Module):
    def __init__(self, img_size=1024, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 out_chans=256, qkv_bias=True, norm_layer: Optional[Type[nn.Module]] = None,
                 act_layer: Type[nn.Module] = nn.GELU, use_abs_pos=True, use_rel_pos=False, rel_pos_zero_init=True,
                 window_size=0, global_attn_indexes: Tuple[int, ...] = (), norm_eps: float = 1e-6):
        super().__init__()
        if norm_layer is None:
            def norm_layer(d):
                return nn.LayerNorm(d, eps=norm_eps)
        self.img_size = img_size
        self.patch_embed = PatchEmbed((patch_size, patch_size), (patch_size, patch_size), in_chans=in_chans,
                                      embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        g = img_size // patch_size
        self.blocks = nn.ModuleList(
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer, act_layer, use_rel_pos, rel_pos_zero_init,
                  window_size if i not in global_attn_indexes else 0, (g, g)) for i in range(depth))
        self.neck = nn.Sequential(nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False), LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
                                  LayerNorm2d(out_chans))
        self.max_images_per_pass = 8  # bounds the [B, heads, 4096, 4096] bias of the global blocks

    def _forward_chunk(self, x):
        x = self.patch_embed(x)
        if self.pos_embed is not None:
            x = x + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        x = x.permute(0, 3, 1, 2)
        if x.dtype == torch.float16:  # the reference runs the neck in fp32 for fp16 models (overflow guard)
            with torch.autocast(device_type="cuda", dtype=torch.float32):
                return self.neck(x).to(torch.float16)
        return self.neck(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = self.max_images_per_pass
        if x.shape[0] <= n:
            return self._forward_chunk(x)
        return torch.cat([self._forward_chunk(x[i:i + n]) for i in range(0, x.shape[0], n)], 0)

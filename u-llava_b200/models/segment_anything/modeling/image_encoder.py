"""SAM ViT image encoder (reference: segment_anything/modeling/image_encoder.py:16-426).

SURVEY.md section 8 row a12 / f1: the encoder is on the u-LLaVA path (models/ullava.py:139-150).
ImageEncoderViT.forward runs ullava_sam_encoder_forward (csrc/sam_encoder.cu): tcgen05 GEMMs for the patch
embedding, qkv / proj / MLP and the neck convolutions, flash attention with the decomposed relative-position
bias for the windowed and global blocks.  The nn.Modules below keep the reference's parameters and
state_dict keys, so SAM checkpoints load unchanged; they are parameter containers: their own forward() methods
raise (the blocks run fused inside ImageEncoderViT.forward), and there is no torch implementation of the encoder in
this package -- the tests cross-check the kernels against oracle/ullava_oracle.py:sam_image_encoder.
Images are processed as one batch (the reference loops image by image with empty_cache() calls)."""
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn
import torch.nn.functional as F

import native

from .common import LayerNorm2d, MLPBlock


class PatchEmbed(nn.Module):
    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        raise RuntimeError("PatchEmbed runs fused inside ImageEncoderViT.forward (ullava_sam_encoder_forward: im2col + "
                           "tcgen05 GEMM); this module only holds the parameters")


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
        raise RuntimeError("Attention runs fused inside ImageEncoderViT.forward (tcgen05 flash attention with the "
                           "decomposed relative-position bias); this module only holds the parameters")


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
        raise RuntimeError("Block runs fused inside ImageEncoderViT.forward (ullava_sam_encoder_forward); this module "
                           "only holds the parameters")


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
        self.patch_size, self.embed_dim, self.depth, self.num_heads = patch_size, embed_dim, depth, num_heads
        self.out_chans, self.window_size, self.norm_eps = out_chans, window_size, norm_eps
        self.global_attn_indexes = tuple(global_attn_indexes)
        self.use_rel_pos = use_rel_pos
        self.max_images_per_pass = 32  # bounds the activation scratch (about 0.2 GB per image for ViT-H)
        self._packed = None
        self._maps = {}
        self._scratch = None

    # ---- native path ---------------------------------------------------------------------------------
    @staticmethod
    def _rel_resized(rel_pos: torch.Tensor, size: int) -> torch.Tensor:
        """Relative-position table with 2*size-1 rows (linear resize like the reference's get_rel_pos :321-352)."""
        n = 2 * size - 1
        if rel_pos.shape[0] == n:
            return rel_pos
        return F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1).float(), size=n,
                             mode="linear").reshape(-1, n).permute(1, 0).to(rel_pos.dtype)

    def _pack(self):
        if not self.use_rel_pos:
            raise NotImplementedError("the native SAM encoder is built for use_rel_pos=True (all build_sam_* variants)")
        dev, dt = self.patch_embed.proj.weight.device, self.patch_embed.proj.weight.dtype
        g = self.img_size // self.patch_size
        D, hd = self.embed_dim, self.embed_dim // self.num_heads
        z = lambda n: torch.zeros(n, dtype=dt, device=dev)  # noqa: E731
        pe = self.patch_embed.proj
        t = [pe.weight.reshape(D, -1), pe.bias if pe.bias is not None else z(D),
             self.pos_embed.reshape(g * g, D) if self.pos_embed is not None else torch.zeros((g * g, D), dtype=dt, device=dev)]
        gmask = 0
        for i, blk in enumerate(self.blocks):
            is_global = blk.window_size == 0
            if is_global:
                gmask |= 1 << i
            S = g if is_global else blk.window_size
            a = blk.attn
            # [Rh; Rw] back to back in one buffer: the windowed-attention kernel loads both tables with one TMA box
            rel_hw = torch.cat([self._rel_resized(a.rel_pos_h, S), self._rel_resized(a.rel_pos_w, S)]).detach().to(dt).contiguous()
            n_rel = rel_hw.shape[0] // 2
            t += [blk.norm1.weight, blk.norm1.bias, a.qkv.weight, a.qkv.bias if a.qkv.bias is not None else z(3 * D),
                  rel_hw[:n_rel], rel_hw[n_rel:], a.proj.weight, a.proj.bias,
                  blk.norm2.weight, blk.norm2.bias, blk.mlp.lin1.weight, blk.mlp.lin1.bias, blk.mlp.lin2.weight,
                  blk.mlp.lin2.bias]
            if not isinstance(blk.mlp.act, nn.GELU):
                raise NotImplementedError("SAM encoder MLP activation must be GELU")
        c1, l1, c2, l2 = self.neck[0], self.neck[1], self.neck[2], self.neck[3]
        C = self.out_chans
        t += [c1.weight.reshape(C, D), l1.weight, l1.bias, c2.weight.permute(0, 2, 3, 1).reshape(C, 9 * C), l2.weight,
              l2.bias]
        t = [x.detach().to(dt).contiguous() for x in t]
        cfg = dict(img=self.img_size, patch=self.patch_size, embed_dim=D, depth=self.depth, heads=self.num_heads,
                   window=self.window_size, out_chans=C, global_mask=gmask, eps=float(self.norm_eps))
        return t, native.Context.pointer_table(t), cfg

    def _row_maps(self, B: int, device):
        """window_partition / window_unpartition (reference :263-318) as row maps over [B*g*g] tokens."""
        key = (B, str(device))
        if key not in self._maps:
            g, ws = self.img_size // self.patch_size, self.window_size
            if ws <= 0:
                self._maps[key] = (None, None)
            else:
                nw = (g + ws - 1) // ws
                b = torch.arange(B, device=device)[:, None, None]
                y = torch.arange(g, device=device)[None, :, None]
                x = torch.arange(g, device=device)[None, None, :]
                win = ((b * nw + y // ws) * nw + x // ws) * (ws * ws) + (y % ws) * ws + (x % ws)
                win = win.reshape(-1).to(torch.int32).contiguous()
                unwin = torch.full((B * nw * nw * ws * ws,), -1, dtype=torch.int32, device=device)
                unwin[win.long()] = torch.arange(B * g * g, dtype=torch.int32, device=device)
                self._maps[key] = (win, unwin)
        return self._maps[key]

    def _forward_native(self, x: torch.Tensor, blocks=None) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("ImageEncoderViT (B200 build) runs on a CUDA sm_100 device only; there is no CPU fallback")
        params = list(self.parameters())
        sig = tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in params)
        if self._packed is None or self._packed[0] != sig:
            self._packed = (sig,) + self._pack()
        _, tensors, table, cfg = self._packed
        # native_ctx: the context of the SM-partition lane this call is enqueued on (UllavaForCausalLM.evaluate)
        ctx = getattr(self, "native_ctx", None) or native.Context.get(x.device)
        win, unwin = self._row_maps(x.shape[0], x.device)
        out, self._scratch = ctx.sam_encoder_forward(table, len(tensors), x.to(tensors[0].dtype).contiguous(), cfg, win,
                                                     unwin, self._scratch, blocks=blocks)
        return out

    def forward_blocks(self, x: torch.Tensor, begin: int, end: int):
        """Blocks [begin, end) only (patch embedding with begin == 0, neck with end == depth): the encoder cut in two
        for UllavaForCausalLM.evaluate, which runs the first part on an SM-partition lane beside the decode steps.
        Returns the embeddings from the call that ends at depth, None before."""
        if x.shape[0] > self.max_images_per_pass:
            raise ValueError("forward_blocks works on one pass of at most max_images_per_pass images")
        with torch.no_grad():
            return self._forward_native(x, blocks=(begin, end))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = self.max_images_per_pass
        with torch.no_grad():
            if x.shape[0] <= n:
                return self._forward_native(x)
            return torch.cat([self._forward_native(x[i:i + n]) for i in range(0, x.shape[0], n)], 0)

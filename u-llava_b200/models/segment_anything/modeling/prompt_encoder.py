"""SAM prompt encoder, text-prompt path (reference: segment_anything/modeling/prompt_encoder.py).

u-LLaVA only ever calls forward(points=None, boxes=None, masks=None, text_embeds=...) and
get_dense_pe() (models/ullava.py:232-243,405-417).  Both are input-independent glue: the sparse
embedding IS the text embedding, the dense embedding is a broadcast view of no_mask_embed, and the
dense positional encoding is a constant of the module that is computed once (in fp32, rounded to the
buffer dtype) and cached."""
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .common import LayerNorm2d


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None):
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def _pe_encoding(self, coords: torch.Tensor) -> torch.Tensor:
        """sin/cos(2*pi*(2c-1)@G) evaluated in fp32 from the stored (possibly 16-bit) matrix and rounded once at
        the end.  The reference evaluates it in the buffer dtype (prompt_encoder.py:203-229), which under bf16
        carries ~0.1 absolute noise in the angle; the fp32 evaluation stays inside that noise band and matches
        the fp32 reference to the last bit of the 16-bit output."""
        g = self.positional_encoding_gaussian_matrix
        coords = 2 * coords.float() - 1
        coords = 2 * math.pi * (coords @ g.float())
        return torch.cat([coords.sin(), coords.cos()], dim=-1).to(g.dtype)

    def forward(self, size: Tuple[int, int]) -> torch.Tensor:
        h, w = size
        g = self.positional_encoding_gaussian_matrix
        ones = torch.ones((h, w), device=g.device, dtype=torch.float32)
        y = (ones.cumsum(dim=0) - 0.5) / h
        x = (ones.cumsum(dim=1) - 0.5) / w
        return self._pe_encoding(torch.stack([x, y], dim=-1)).permute(2, 0, 1)

    def forward_with_coords(self, coords_input, image_size):
        coords = coords_input.clone()
        coords[:, :, 0] = coords[:, :, 0] / image_size[1]
        coords[:, :, 1] = coords[:, :, 1] / image_size[0]
        return self._pe_encoding(coords.to(torch.float))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation=nn.GELU):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4
        self.point_embeddings = nn.ModuleList(nn.Embedding(1, embed_dim) for _ in range(4))
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans),
            activation(), nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)
        self._pe_cache = None

    def get_dense_pe(self) -> torch.Tensor:
        g = self.pe_layer.positional_encoding_gaussian_matrix
        key = (g.data_ptr(), g._version, g.dtype, str(g.device))
        if self._pe_cache is None or self._pe_cache[0] != key:
            self._pe_cache = (key, self.pe_layer(self.image_embedding_size).unsqueeze(0).contiguous())
        return self._pe_cache[1]

    def forward(self, points, boxes, masks, text_embeds):
        if points is not None or boxes is not None or masks is not None:
            raise NotImplementedError("point / box / mask prompts are outside the u-LLaVA path "
                                      "(the reference only passes text_embeds, models/ullava.py:232-238)")
        w = self.no_mask_embed.weight
        bs = text_embeds.shape[0] if text_embeds is not None else 1
        sparse = torch.empty((bs, 0, self.embed_dim), device=w.device)
        if text_embeds is not None:
            sparse = torch.cat([sparse, text_embeds], dim=1)
        dense = w.reshape(1, -1, 1, 1).expand(bs, -1, self.image_embedding_size[0], self.image_embedding_size[1])
        return sparse, dense

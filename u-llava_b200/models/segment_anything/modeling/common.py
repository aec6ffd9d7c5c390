"""Parameter containers shared by the SAM modules (reference: segment_anything/modeling/common.py)."""
import torch
import torch.nn as nn


class MLPBlock(nn.Module):
    """lin1 -> act -> lin2.  Holds parameters; the hot path reads them through the C ABI."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act=nn.GELU):
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()

    def forward(self, x):
        raise RuntimeError("MLPBlock runs as two tcgen05 GEMMs with fused epilogues inside the native SAM encoder / "
                           "mask decoder; this module only holds the parameters")


class LayerNorm2d(nn.Module):
    """Channel LayerNorm for NCHW tensors (biased variance, eps inside the sqrt)."""

    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):
        raise RuntimeError("LayerNorm2d runs inside the native SAM encoder neck / mask decoder (ullava_layernorm on "
                           "channels-last tokens); this module only holds the parameters")

"""Builders shared by the model-level GPU tests, __graft_entry__.smoke() and bench.py's parity check:
the B200 implementation loaded with the same synthetic state_dict (oracle/synth.py) that generated
tests/golden/*, plus the matching inputs."""
import json
import os

import numpy as np
import torch

from oracle.synth import synth_normal, synth_state_dict
from tests import configs as C

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        meta = json.load(f)
    return np.load(os.path.join(GOLD, name + ".npz")), meta


def core_cfg():
    cfg = dict(C.TINY_LLM)
    cfg["rope_theta"] = 10000.0
    return cfg


def build_tiny_core(dtype, device="cuda"):
    """UllavaCoreForCausalLM (B200 build) with the tiny_core golden weights.  Returns (model, sd_fp32, cfg)."""
    import models
    _, meta = load_golden("tiny_core")
    sd = synth_state_dict(meta["shapes"], meta["seed"])
    cfg = models.UllavaCoreConfig(**C.TINY_LLM)
    m = models.UllavaCoreForCausalLM(cfg).eval()
    m.load_state_dict(sd, strict=True)
    return m.to(device=device, dtype=dtype), sd, core_cfg()


def build_tiny_full(dtype, device="cuda"):
    """UllavaForCausalLM (B200 build) with the tiny_full golden weights (2-block SAM image encoder,
    full-geometry prompt encoder / mask decoder)."""
    import models
    from models.segment_anything.build_sam import _build_sam
    _, meta = load_golden("tiny_full")
    sd = synth_state_dict(meta["shapes"], meta["seed"])
    e = C.TINY_SAM_ENCODER
    cfg = models.UllavaConfig(llm_config=dict(C.TINY_LLM), seg_token_idx=C.SEG_ID, loc_token_idx=C.LOC_ID)

    class TinyUllava(models.UllavaForCausalLM):
        sam_builder = staticmethod(lambda checkpoint=None: _build_sam(e["embed_dim"], e["depth"], e["num_heads"],
                                                                     e["global_attn_indexes"]))

    m = TinyUllava(cfg).eval()
    m.load_state_dict(sd, strict=True)
    return m.to(device=device, dtype=dtype), sd, core_cfg()


def oracle_inputs_core(batch=2):
    return C.tiny_prompt(batch), synth_normal("images", (batch, 3, 28, 28))


def oracle_inputs_full():
    _, meta = load_golden("tiny_full")
    ids = C.tiny_prompt(2, seg_loc=True)
    images = synth_normal("images", (2, 3, 28, 28))
    images_sam = synth_normal("images_sam", (2, 3, 1024, 1024))
    sizes = [tuple(s) for s in meta["sizes"]]
    resizes = [tuple(s) for s in meta["resizes"]]
    return ids, images, images_sam, sizes, resizes


def iou(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a > 0, b > 0
    u = (a | b).sum().item()
    return 1.0 if u == 0 else (a & b).sum().item() / u

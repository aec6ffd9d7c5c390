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


def build_tiny_full(dtype, device="cuda", coherent=False):
    """UllavaForCausalLM (B200 build) with the tiny_full golden weights (2-block SAM image encoder,
    full-geometry prompt encoder / mask decoder).  coherent: with oracle.synth.coherent_overrides (masks with large
    margins, for the IoU >= 0.999 gate)."""
    import models
    from models.segment_anything.build_sam import _build_sam
    _, meta = load_golden("tiny_full")
    sd = synth_state_dict(meta["shapes"], meta["seed"])
    if coherent:
        from oracle.synth import coherent_overrides
        sd = coherent_overrides(sd)
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


class LoraLinearStub(torch.nn.Linear):
    """Attribute layout of peft 0.4's `lora.Linear` (the reference's pin; peft is not installed here): an nn.Linear
    whose `.weight` is the frozen base weight plus per-adapter lora_A / lora_B modules, `scaling`, `r`, `merged`."""

    def __init__(self, base: torch.nn.Linear, r: int, alpha: float, key: str, adapter="default"):
        super().__init__(base.in_features, base.out_features, bias=base.bias is not None)
        self.weight, self.bias = base.weight, base.bias
        self.lora_A = torch.nn.ModuleDict({adapter: torch.nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = torch.nn.ModuleDict({adapter: torch.nn.Linear(r, base.out_features, bias=False)})
        self.lora_embedding_A, self.lora_embedding_B = torch.nn.ParameterDict(), torch.nn.ParameterDict()
        self.r, self.scaling = {adapter: r}, {adapter: alpha / r}
        self.active_adapter, self.merged, self.disable_adapters, self.fan_in_fan_out = adapter, False, False, False
        with torch.no_grad():
            self.lora_A[adapter].weight.copy_(synth_normal(key + ".A", (r, base.in_features), scale=0.12))
            self.lora_B[adapter].weight.copy_(synth_normal(key + ".B", (base.out_features, r), scale=0.12))

    def delta(self) -> torch.Tensor:
        a = self.active_adapter
        return (self.lora_B[a].weight.float() @ self.lora_A[a].weight.float()) * self.scaling[a]


class PeftModelStub(torch.nn.Module):
    """What `PeftModel.from_pretrained(model.llm, ...)` leaves in `model.llm` (inference_ullava.py:43): a wrapper whose
    forward / generate / attribute lookups end up at the wrapped module (base_model.model)."""

    def __init__(self, model):
        super().__init__()
        self.base_model = torch.nn.Module()
        self.base_model.model = model

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)

    def forward(self, *a, **k):
        return self.base_model.model(*a, **k)

    def generate(self, **k):
        return self.base_model.model.generate(**k)


def inject_lora(llama_model, targets=("q_proj", "v_proj"), r=4, alpha=8.0):
    """Wraps the target projections of every decoder layer in LoraLinearStub (train_ullava.py:42 default targets);
    returns {state_dict key of the base weight: the stub} (stub.delta() = scaling * B @ A of its current weights)."""
    deltas = {}
    for i, lay in enumerate(llama_model.layers):
        for t in targets:
            parent = lay.self_attn if hasattr(lay.self_attn, t) else lay.mlp
            stub = LoraLinearStub(getattr(parent, t), r, alpha, f"lora.{i}.{t}")
            setattr(parent, t, stub)
            sub = "self_attn" if parent is lay.self_attn else "mlp"
            deltas[f"model.layers.{i}.{sub}.{t}.weight"] = stub
    return deltas


def greedy_walk(got: torch.Tensor, ref: torch.Tensor, margins: torch.Tensor, prompt_len: int, min_margin: float,
                got_prompt_len: int = None):
    """Per-token comparison of greedy ids with the reference's.  got [T_got], ref [T_ref] (one sample), margins [n] =
    the reference's top-2 logit margin of every generated token.  A token decided by a margin above `min_margin` (twice
    the logit tolerance) must be identical; at a narrower margin a flip is legitimate under a different reduction
    order, and everything after it is conditioned differently, so the walk stops there.
    Returns (n_checked_exact, n_compared_prefix): tokens asserted under a wide margin / length of the common prefix."""
    gp = prompt_len if got_prompt_len is None else got_prompt_len
    n = margins.shape[0]
    exact = 0
    for t in range(n):
        a, b = int(got[gp + t]), int(ref[prompt_len + t])
        if float(margins[t]) > min_margin:
            assert a == b, f"token {t}: got {a}, reference {b} at margin {float(margins[t]):.4f} > {min_margin}"
            exact += 1
        elif a != b:
            return exact, t
    return exact, n

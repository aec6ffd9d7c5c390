#!/usr/bin/env python
"""bench.py -- images/sec of the u-LLaVA inference hot path on B200 (contract: see DESIGN.md "Measurement").

One step = UllavaForCausalLM.evaluate() on one batch of synthetic inputs per GPU:
  336x336 image + 608-token prompt (576 visual + 32 text, one [SEG] in the prompt tail)
  -> CLIP ViT-L/14-336 -> projector -> LLaMA-7B prefill -> 64 greedy tokens (KV cached) -> SAM ViT-H image
  embedding -> [SEG] hidden state -> seg projector -> SAM mask decoder -> 336x336 mask logits,
followed (N > 1) by the single all-gather of token ids + bit-packed masks.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference ...                     the CPU arm: oracle port of the reference on host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "u-llava_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "images/sec (336px img + 32-tok prompt, 64-tok gen + mask)"
UNIT = "images/s"
IMG, SAM_IMG, N_PATCH, P_LEN = 336, 1024, 576, 608
MM_IDS = dict(IMG_PATCH=32001, VID_PATCH=32002, IMG_START=32003, IMG_END=32004, VID_START=32005, VID_END=32006)
SEG_ID, LOC_ID, VOCAB = 32007, 32008, 32011
VISION = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=IMG,
              patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
LLM = dict(vocab_size=VOCAB, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
           num_key_value_heads=32, rms_norm_eps=1e-6, max_position_embeddings=2048, vision_config=VISION,
           vision_hidden_layer=-2, projector_type="mlp", mm_token_ids=MM_IDS, bos_token_id=1, eos_token_id=None,
           pad_token_id=32000)

# algorithmic work per image (BASELINE.md section 4 / SURVEY.md section 8d)
FLOP_VIT, FLOP_PROJ, FLOP_PREFILL_LAST = 366.0e9, 4.83e9, 7.97e12
FLOP_SAM_ENC, FLOP_MASK_DEC = 5.67e12, 3.61e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=None, help="default: the config's (c2 32, c3 16, c4 8, c5 32)")
    ap.add_argument("--new-tokens", type=int, default=64)
    ap.add_argument("--config", default="c5", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs: c2 = ViT-L/14-336 only, B=32; c3 = LLaMA-7B prefill only, B=16, L=608; "
                         "c4 = full pipeline B=8; c5 (default) = full pipeline, 32 images per GPU (the metric's config)")
    ap.add_argument("--cpu-full-image", action="store_true",
                    help="--impl reference only: also run ONE whole image through the full-size oracle (no sampling, "
                         "about 27 GB of host memory and a few minutes) and report the extrapolation error")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    ap.add_argument("--no-overlap", action="store_true",
                    help="run the SAM image encoder after the decode steps instead of beside them (SM partition off)")
    ap.add_argument("--overlap-sms", type=int, default=0,
                    help="SMs of the decode lane of the SM partition (default: the model's, ULLAVA_OVERLAP_SMS)")
    ap.add_argument("--overlap-blocks", type=int, default=0,
                    help="SAM encoder blocks run on the second lane beside the decode steps (default: the model's)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="one resident step only, no e2e / stages / cpu baseline (for runs under ncu)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_prompt(global_index: int) -> torch.Tensor:
    """[BOS] + 13 text + <img_beg> + 576 x <image_patch> + </img_end> + 16 text (one of them [SEG]) = 608 ids."""
    g = torch.Generator().manual_seed(1234 + global_index)
    head = torch.randint(3, 32000, (13,), generator=g)
    tail = torch.randint(3, 32000, (16,), generator=g)
    tail[11] = SEG_ID
    ids = torch.cat([torch.tensor([1]), head, torch.tensor([MM_IDS["IMG_START"]]),
                     torch.full((N_PATCH,), MM_IDS["IMG_PATCH"]), torch.tensor([MM_IDS["IMG_END"]]), tail])
    assert ids.numel() == P_LEN
    return ids


def make_inputs(lo: int, hi: int, dtype):
    """Pinned host tensors for global images lo..hi-1 (seeded per global index: world-size invariant)."""
    n = hi - lo
    ids = torch.stack([make_prompt(i) for i in range(lo, hi)]).pin_memory()
    images = torch.empty((n, 3, IMG, IMG), dtype=dtype).pin_memory()
    images_sam = torch.empty((n, 3, SAM_IMG, SAM_IMG), dtype=dtype).pin_memory()
    for j, i in enumerate(range(lo, hi)):
        g = torch.Generator().manual_seed(1234 + i)
        images[j] = torch.randn((3, IMG, IMG), generator=g).to(dtype)
        images_sam[j] = torch.randn((3, SAM_IMG, SAM_IMG), generator=g).to(dtype)
    return ids, images, images_sam


def build_model(device, dtype):
    import models
    cfg = models.UllavaConfig(llm_config=dict(LLM), seg_token_idx=SEG_ID, loc_token_idx=LOC_ID)
    torch.manual_seed(0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = models.UllavaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(old)
    model = model.eval().to(device=device, dtype=dtype)
    # SAM's rel-pos tables are zero-initialised by the reference constructor; give them seeded values so the
    # bias path does real work (a trained checkpoint has them non-zero)
    g = torch.Generator(device=device).manual_seed(7)
    for n_, p in model.visual_model.image_encoder.named_parameters():
        if "rel_pos" in n_ or n_ == "pos_embed":
            p.data.copy_((0.02 * torch.randn(p.shape, generator=g, device=device)).to(dtype))
    return model


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi spends its first second initialising NVML (driver locks that can stall kernel launches): let
            # it deliver its first sample before the warm-up starts, so that only steady 200 ms polling overlaps the run
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on the host cores (bounded sample, extrapolated per image)
# ------------------------------------------------------------------------------------------------
class CpuSample:
    """Times full-width pieces of the reference algorithm (oracle/ullava_oracle.py, fp32, KV-cached greedy loop,
    which is favourable to the reference: its shipped checkpoints decode with use_cache=False) and scales them to
    one image of the bench workload:  2/23 CLIP layers, 1/32 LLaMA layers (prefill L=608 and 4 cached decode steps),
    lm_head (all 608 prefill positions as the reference computes them, and per decode step), 1 windowed + 1 global
    SAM ViT-H block of 32 (28 windowed + 4 global), patch embeds / neck / mask decoder / post-process in full."""

    description = ("oracle port, fp32, B=1: 2 of 23 CLIP-L/14-336 layers, 1 of 32 LLaMA-7B layers (prefill L=608 + 4 "
                   "KV-cached decode steps), lm_head V=32011, 1 windowed + 1 global of 32 SAM ViT-H blocks, full SAM "
                   "mask decoder + post-process; scaled by layer/token counts to one image with 64 generated tokens")

    def __init__(self):
        from oracle.synth import synth_normal, synth_tensor
        self.O = __import__("oracle.ullava_oracle", fromlist=["x"])
        torch.set_grad_enabled(False)
        st = lambda k, s: synth_tensor(k, s, 0)  # noqa: E731
        sd = {}
        # CLIP, 2 layers
        vp = "vision_encoder.vision_model."
        sd[vp + "embeddings.patch_embedding.weight"] = st("pe", (1024, 3, 14, 14))
        sd[vp + "embeddings.class_embedding"] = st("cls", (1024,))
        sd[vp + "embeddings.position_embedding.weight"] = st("pos", (577, 1024))
        for nm in ("pre_layrnorm",):
            sd[vp + nm + ".weight"] = st(nm + ".weight", (1024,)); sd[vp + nm + ".bias"] = st(nm + ".bias", (1024,))
        for l in range(2):
            q = vp + f"encoder.layers.{l}."
            for nm, (o, i) in dict(q_proj=(1024, 1024), k_proj=(1024, 1024), v_proj=(1024, 1024),
                                   out_proj=(1024, 1024)).items():
                sd[q + f"self_attn.{nm}.weight"] = st(q + nm + ".weight", (o, i))
                sd[q + f"self_attn.{nm}.bias"] = st(q + nm + ".bias", (o,))
            for nm, (o, i) in dict(fc1=(4096, 1024), fc2=(1024, 4096)).items():
                sd[q + f"mlp.{nm}.weight"] = st(q + nm + ".weight", (o, i)); sd[q + f"mlp.{nm}.bias"] = st(q + nm + ".bias", (o,))
            for nm in ("layer_norm1", "layer_norm2"):
                sd[q + nm + ".weight"] = st(q + nm + ".weight", (1024,)); sd[q + nm + ".bias"] = st(q + nm + ".bias", (1024,))
        sd["vision_projector.weight"] = st("proj.weight", (4096, 1024)); sd["vision_projector.bias"] = st("proj.bias", (4096,))
        # LLaMA, 1 layer + lm_head
        p = "model.layers.0."
        for nm, shp in {"self_attn.q_proj": (4096, 4096), "self_attn.k_proj": (4096, 4096), "self_attn.v_proj": (4096, 4096),
                        "self_attn.o_proj": (4096, 4096), "mlp.gate_proj": (11008, 4096), "mlp.up_proj": (11008, 4096),
                        "mlp.down_proj": (4096, 11008)}.items():
            sd[p + nm + ".weight"] = st(p + nm, shp)
        for nm in ("input_layernorm", "post_attention_layernorm"):
            sd[p + nm + ".weight"] = st(p + nm + ".weight", (4096,))
        sd["model.norm.weight"] = st("model.norm.weight", (4096,))
        sd["lm_head.weight"] = st("lm_head.weight", (VOCAB, 4096))
        # SAM ViT-H: 2 blocks (window, global) + patch embed + neck + prompt encoder / mask decoder
        from tests.util_models import load_golden
        _, meta = load_golden("sam_decoder")
        for k, shp in meta["shapes"].items():
            sd["visual_model." + k] = st(k, shp)
        e = "visual_model.image_encoder."
        sd[e + "patch_embed.proj.weight"] = st("sam.pe.w", (1280, 3, 16, 16)); sd[e + "patch_embed.proj.bias"] = st("sam.pe.bias", (1280,))
        sd[e + "pos_embed"] = st("sam.pos_embed", (1, 64, 64, 1280))
        for b, size in ((0, 14), (1, 64)):
            q = e + f"blocks.{b}."
            for nm, shp in {"norm1.weight": (1280,), "norm1.bias": (1280,), "norm2.weight": (1280,), "norm2.bias": (1280,),
                            "attn.qkv.weight": (3840, 1280), "attn.qkv.bias": (3840,), "attn.proj.weight": (1280, 1280),
                            "attn.proj.bias": (1280,), "attn.rel_pos_h": (2 * size - 1, 80), "attn.rel_pos_w": (2 * size - 1, 80),
                            "mlp.lin1.weight": (5120, 1280), "mlp.lin1.bias": (5120,), "mlp.lin2.weight": (1280, 5120),
                            "mlp.lin2.bias": (1280,)}.items():
                sd[q + nm] = st(q + nm, shp)
        sd[e + "neck.0.weight"] = st("neck0", (256, 1280, 1, 1)); sd[e + "neck.2.weight"] = st("neck2", (256, 256, 3, 3))
        for i in (1, 3):
            sd[e + f"neck.{i}.weight"] = st(f"neck{i}.weight", (256,)); sd[e + f"neck.{i}.bias"] = st(f"neck{i}.bias", (256,))
        self.sd = sd
        self.px = synth_normal("bench_px", (1, 3, IMG, IMG))
        self.px_sam = synth_normal("bench_px_sam", (1, 3, SAM_IMG, SAM_IMG))
        self.x = synth_normal("bench_x", (1, P_LEN, 4096))
        self.text = synth_normal("bench_text", (1, 256))

    @staticmethod
    def scale_factors(new_tokens: int) -> dict:
        """multipliers applied to the timed pieces (run_once): per image = sum(piece x factor)"""
        return {"clip_layer": 23, "prefill_layer": 32, "lm_head_prefill": 1, "decode_layer": 32 * (new_tokens - 1),
                "lm_head_step": new_tokens - 1, "sam_window_block": 28, "sam_global_block": 4, "sam_embed_neck": 1,
                "mask_decoder": 1, "postprocess": 1}

    @staticmethod
    def _t(fn):
        t0 = time.perf_counter()
        r = fn()
        return time.perf_counter() - t0, r

    def run_once(self, new_tokens: int) -> dict:
        O, sd = self.O, self.sd
        import torch.nn.functional as F
        vc = dict(VISION, num_hidden_layers=2)
        t_clip0, _ = self._t(lambda: O.clip_vit_hidden(sd, "vision_encoder.", self.px, vc, 0))
        t_clip2, h = self._t(lambda: O.clip_vit_hidden(sd, "vision_encoder.", self.px, vc, 2))
        t_proj, _ = self._t(lambda: F.linear(h[:, 1:], sd["vision_projector.weight"], sd["vision_projector.bias"]))
        lc = dict(hidden_size=4096, num_attention_heads=32, num_key_value_heads=32, num_hidden_layers=1)
        t_pre, (hl, past, _) = self._t(lambda: O.llama_layers(sd, self.x, lc))
        t_head_all, _ = self._t(lambda: F.linear(hl, sd["lm_head.weight"]))
        t_dec = 0.0
        for _ in range(4):
            dt, (h1, past, _) = self._t(lambda: O.llama_layers(sd, self.x[:, :1], lc, past=past))
            t_dec += dt / 4
        t_head1, _ = self._t(lambda: F.linear(h1, sd["lm_head.weight"]))
        ew = dict(embed_dim=1280, depth=1, num_heads=16, global_attn_indexes=[], window_size=14, patch_size=16)
        e0 = dict(ew, depth=0)
        t_s0, _ = self._t(lambda: O.sam_image_encoder(sd, "visual_model.", self.px_sam, e0))
        t_sw, emb = self._t(lambda: O.sam_image_encoder(sd, "visual_model.", self.px_sam, ew))
        sdg = dict(sd)
        for k in list(sd):
            if ".blocks.1." in k:
                sdg[k.replace(".blocks.1.", ".blocks.0.")] = sd[k]
        eg = dict(ew, global_attn_indexes=[0])
        t_sg, _ = self._t(lambda: O.sam_image_encoder(sdg, "visual_model.", self.px_sam, eg))
        t_md, (masks, _) = self._t(lambda: O.sam_mask_decoder(sd, "visual_model.", emb, self.text))
        t_post, _ = self._t(lambda: O.postprocess_masks(masks[:, 0:1], (SAM_IMG, SAM_IMG), (IMG, IMG)))
        per_image = (t_clip0 + (t_clip2 - t_clip0) / 2 * 23 + t_proj + 32 * t_pre + t_head_all +
                     (new_tokens - 1) * (32 * t_dec + t_head1) + t_s0 + 28 * max(t_sw - t_s0, 0) + 4 * max(t_sg - t_s0, 0) +
                     t_md + t_post)
        return dict(per_image_s=per_image, parts=dict(clip2=t_clip2, prefill_layer=t_pre, lm_head_prefill=t_head_all,
                                                      decode_layer=t_dec, lm_head_step=t_head1, sam_window_block=t_sw - t_s0,
                                                      sam_global_block=t_sg - t_s0, sam_embed_neck=t_s0, mask_decoder=t_md,
                                                      postprocess=t_post))


def config_dict(args, world_size):
    if args.config == "c2":
        return {"workload": f"c2: CLIP ViT-L/14-336 encoder only (23 layers, CLS dropped), batch {args.batch_per_gpu}",
                "global_batch": args.batch_per_gpu * world_size, "batch_per_gpu": args.batch_per_gpu, "image": IMG,
                "parallelism": f"dp{world_size}", "l2": "L2 flushed between timed iterations (256 MB write)"}
    if args.config == "c3":
        return {"workload": f"c3: LLaMA-7B prefill only, 576 visual + 32 text positions from resident embeddings, batch "
                            f"{args.batch_per_gpu}, KV cache written, last-position logits",
                "global_batch": args.batch_per_gpu * world_size, "batch_per_gpu": args.batch_per_gpu, "prompt_len": P_LEN,
                "parallelism": f"dp{world_size}", "l2": "inputs larger than L2 (13.5 GB of weights stream through every step)"}
    tag = "c4" if args.config == "c4" else "c5 shard"
    return {"workload": f"{tag}: full u-LLaVA-7B pipeline (CLIP ViT-L/14-336 + projector + LLaMA-7B prefill 608 + "
                        f"{args.new_tokens} greedy tokens + SAM ViT-H + mask decoder), {args.batch_per_gpu} images per GPU",
            "global_batch": args.batch_per_gpu * world_size, "batch_per_gpu": args.batch_per_gpu, "prompt_len": P_LEN,
            "new_tokens": args.new_tokens, "image": IMG, "sam_image": SAM_IMG, "parallelism": f"dp{world_size}",
            "l2": "inputs larger than L2 (13.5 GB of weights stream through every step)"}


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process that owns the host."""
    n = os.cpu_count() or 1
    os.environ.pop("OMP_NUM_THREADS", None)
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_full_image(new_tokens: int) -> dict:
    """ONE image through the whole full-size oracle (23 CLIP layers, 32 LLaMA-7B layers with a KV-cached greedy loop,
    32 SAM ViT-H blocks, mask decoder), nothing sampled or scaled: the check on CpuSample's extrapolation."""
    from oracle import ullava_oracle as O
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(0)
    rn = lambda *s, std=0.02: torch.randn(s, generator=g) * std  # noqa: E731
    sd = {}
    vp = "llm.vision_encoder.vision_model."
    sd[vp + "embeddings.patch_embedding.weight"] = rn(1024, 3, 14, 14)
    sd[vp + "embeddings.class_embedding"] = rn(1024)
    sd[vp + "embeddings.position_embedding.weight"] = rn(577, 1024)
    for nm in ("pre_layrnorm",):
        sd[vp + nm + ".weight"], sd[vp + nm + ".bias"] = torch.ones(1024), torch.zeros(1024)
    for l in range(23):
        q = vp + f"encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[q + f"self_attn.{nm}.weight"], sd[q + f"self_attn.{nm}.bias"] = rn(1024, 1024), torch.zeros(1024)
        sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = rn(4096, 1024), torch.zeros(4096)
        sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = rn(1024, 4096), torch.zeros(1024)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[q + nm + ".weight"], sd[q + nm + ".bias"] = torch.ones(1024), torch.zeros(1024)
    sd["llm.vision_projector.weight"], sd["llm.vision_projector.bias"] = rn(4096, 1024), torch.zeros(4096)
    sd["llm.model.embed_tokens.weight"] = rn(VOCAB, 4096)
    for l in range(32):
        p = f"llm.model.layers.{l}."
        for nm, shp in {"self_attn.q_proj": (4096, 4096), "self_attn.k_proj": (4096, 4096), "self_attn.v_proj": (4096, 4096),
                        "self_attn.o_proj": (4096, 4096), "mlp.gate_proj": (11008, 4096), "mlp.up_proj": (11008, 4096),
                        "mlp.down_proj": (4096, 11008)}.items():
            sd[p + nm + ".weight"] = rn(*shp)
        sd[p + "input_layernorm.weight"] = sd[p + "post_attention_layernorm.weight"] = torch.ones(4096)
    sd["llm.model.norm.weight"] = torch.ones(4096)
    sd["llm.lm_head.weight"] = rn(VOCAB, 4096)
    from tests.util_models import load_golden
    _, meta = load_golden("sam_decoder")
    for k, shp in meta["shapes"].items():
        sd["visual_model." + k] = rn(*shp) if len(shp) > 1 else torch.ones(shp)
    e = "visual_model.image_encoder."
    sd[e + "patch_embed.proj.weight"], sd[e + "patch_embed.proj.bias"] = rn(1280, 3, 16, 16), torch.zeros(1280)
    sd[e + "pos_embed"] = rn(1, 64, 64, 1280)
    for b in range(32):
        size = 64 if b in (7, 15, 23, 31) else 14
        q = e + f"blocks.{b}."
        for nm, shp in {"attn.qkv.weight": (3840, 1280), "attn.proj.weight": (1280, 1280), "mlp.lin1.weight": (5120, 1280),
                        "mlp.lin2.weight": (1280, 5120), "attn.rel_pos_h": (2 * size - 1, 80),
                        "attn.rel_pos_w": (2 * size - 1, 80)}.items():
            sd[q + nm] = rn(*shp)
        for nm, n in {"attn.qkv.bias": 3840, "attn.proj.bias": 1280, "mlp.lin1.bias": 5120, "mlp.lin2.bias": 1280,
                      "norm1.bias": 1280, "norm2.bias": 1280}.items():
            sd[q + nm] = torch.zeros(n)
        sd[q + "norm1.weight"] = sd[q + "norm2.weight"] = torch.ones(1280)
    sd[e + "neck.0.weight"], sd[e + "neck.2.weight"] = rn(256, 1280, 1, 1), rn(256, 256, 3, 3)
    for i in (1, 3):
        sd[e + f"neck.{i}.weight"], sd[e + f"neck.{i}.bias"] = torch.ones(256), torch.zeros(256)
    for nm in ("seg_projector",):
        sd[nm + ".0.weight"], sd[nm + ".0.bias"] = rn(4096, 4096), torch.zeros(4096)
        sd[nm + ".2.weight"], sd[nm + ".2.bias"] = rn(256, 4096), torch.zeros(256)
    cfg = dict(LLM)
    cfg["rope_theta"] = 10000.0
    ids = make_prompt(0)[None]
    px = torch.randn((1, 3, IMG, IMG), generator=g)
    px_sam = torch.randn((1, 3, SAM_IMG, SAM_IMG), generator=g)
    vith = dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=[7, 15, 23, 31], window_size=14, patch_size=16)
    t0 = time.perf_counter()
    seqs, hid, _ = O.greedy_generate(sd, cfg, ids, px, new_tokens, prefix="llm.")
    t1 = time.perf_counter()
    emb = O.sam_image_encoder(sd, "visual_model.", px_sam, vith)
    t2 = time.perf_counter()
    h = hid[:, -1]                                     # any one hidden state: the mask head's cost does not depend on it
    text = O.seg_project(sd, "seg_projector.", h)
    masks, _ = O.sam_mask_decoder(sd, "visual_model.", emb, text)
    O.postprocess_masks(masks[:, 0:1], (SAM_IMG, SAM_IMG), (IMG, IMG))
    t3 = time.perf_counter()
    return {"per_image_s": t3 - t0, "generate_s": t1 - t0, "sam_encoder_s": t2 - t1, "mask_head_s": t3 - t2}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = use_all_host_threads()
    cpu = CpuSample()
    for _ in range(min(args.warmup, 1)):
        cpu.run_once(args.new_tokens)
    rs = [cpu.run_once(args.new_tokens) for _ in range(args.steps)]
    per_image = statistics.median(r["per_image_s"] for r in rs)
    v = 1.0 / per_image
    B = args.batch_per_gpu
    # ms_per_step: what ONE step of the arm's config (B images) costs at this rate -- the sample itself is a fraction of
    # one image (CpuSample.description) scaled by layer / token counts, hence "extrapolated"
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * B / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, ws), "extrapolated": True,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": CpuSample.description, "host_cpus": os.cpu_count(), "parts_s": rs[-1]["parts"],
                             "scale_factors": CpuSample.scale_factors(args.new_tokens),
                             "sample_wall_s": round(sum(sum(r["parts"].values()) for r in rs) / max(len(rs), 1), 3)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.cpu_full_image:
        full = cpu_full_image(args.new_tokens)
        line["cpu_baseline"]["full_image"] = dict(full, images_per_s=1.0 / full["per_image_s"],
                                                  extrapolation_error=per_image / full["per_image_s"] - 1.0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import native
    import dist_eval

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16
    B = args.batch_per_gpu
    lo, hi = rank * B, (rank + 1) * B
    ctx = native.Context.get(local)
    model = build_model(dev, dtype)
    model.pack_mask_bits = True
    if args.no_overlap or args.profile_mode:
        model.overlap_sam = False     # (ncu cannot attach to kernels launched into green-context streams)
    if args.overlap_sms:
        model.overlap_sms_decode = args.overlap_sms
    if args.overlap_blocks:
        model.overlap_sam_blocks = args.overlap_blocks
    h_ids, h_img, h_sam = make_inputs(lo, hi, dtype)
    sizes = [(IMG, IMG)] * B
    resizes = [(SAM_IMG, SAM_IMG)] * B
    words = (IMG * IMG + 31) // 32
    T = P_LEN + args.new_tokens
    MAX_MASKS = 2
    graph_launches = [0]

    def core(ids, img, sam):
        out_ids, masks, boxes = model.evaluate(sam, img, ids, sizes, resizes, max_new_tokens=args.new_tokens,
                                               temperature=0)
        payload = dist_eval.pack_results(out_ids, model.last_mask_bits, MAX_MASKS, words)
        gathered = dist_eval.all_gather_results(payload, [B] * ws)
        return out_ids, masks, gathered

    d_ids, d_img, d_sam = h_ids.to(dev), h_img.to(dev), h_sam.to(dev)

    def step_resident():
        return core(d_ids, d_img, d_sam)

    h_out_ids = torch.empty((B, T), dtype=torch.int64).pin_memory()
    h_masks = torch.empty((B, IMG, IMG), dtype=torch.float32).pin_memory()
    h_gather = torch.empty((B * ws, 1 + 2 * T + MAX_MASKS * words), dtype=torch.int32).pin_memory()

    def step_e2e():
        ids = h_ids.to(dev, non_blocking=True)
        img = h_img.to(dev, non_blocking=True)
        sam = h_sam.to(dev, non_blocking=True)
        out_ids, masks, gathered = core(ids, img, sam)
        h_out_ids.copy_(out_ids, non_blocking=True)
        h_masks.copy_(torch.stack([m[0] for m in masks]), non_blocking=True)
        h_gather.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host owns the result before the next step starts
        return out_ids, masks, gathered

    def timed(fn, warmup, steps):
        # warm-up with the SAME liveness pattern as the timed loop (the previous step's outputs stay referenced while the
        # next step runs): otherwise the caching allocator meets that footprint for the first time at the second timed
        # step and grows the pool there (8-9 cudaMalloc calls, +60..300 ms on that one step)
        r = None
        for _ in range(warmup):
            r = fn()
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        n0 = native.Context.total_launches(local) + model.llm.graph_kernel_launches()
        st0 = torch.cuda.memory_stats(dev)
        e0.record()
        for i in range(steps):
            r = fn()
            marks[i].record()
        e1.record()
        torch.cuda.synchronize()
        st1 = torch.cuda.memory_stats(dev)
        prev = e0
        step_ms = []
        for m in marks:
            step_ms.append(round(prev.elapsed_time(m), 1))
            prev = m
        timed.last_info = {"step_ms": step_ms,
                           "cuda_mallocs": int(st1.get("num_device_alloc", 0) - st0.get("num_device_alloc", 0)),
                           "cuda_frees": int(st1.get("num_device_free", 0) - st0.get("num_device_free", 0)),
                           "alloc_retries": int(st1.get("num_alloc_retries", 0) - st0.get("num_alloc_retries", 0))}
        if ws > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), native.Context.total_launches(local) + model.llm.graph_kernel_launches() - n0, r

    if args.config in ("c2", "c3"):
        run_stage_config(args, model, ctx, timed, (h_ids, h_img, d_ids, d_img), rank, ws, local, dev)
        if ws > 1:
            dist.destroy_process_group()
        return

    if args.profile_mode:
        step_resident()
        torch.cuda.synchronize()
        print(json.dumps({"profile_mode": True, "launches": native.Context.total_launches(local) + model.llm.graph_kernel_launches()}))
        return

    sampler = ClockSampler(physical_gpu_index(local))
    if rank == 0 and not os.environ.get("ULLAVA_BENCH_NO_CLOCKS"):
        sampler.start()
    ms_res, launches, res = timed(step_resident, args.warmup, args.steps)
    info_res = timed.last_info
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, res2 = timed(step_e2e, 2, args.steps)
    info_e2e = timed.last_info

    out_ids, masks, gathered = res
    assert out_ids.shape == (B, T), out_ids.shape
    n_masks = sum(int(m.shape[0]) for m in masks)
    assert n_masks >= B, "every image must produce at least one mask"
    assert gathered.shape[0] == B * ws

    # ---- stage timeline + per-kernel-class CUDA-event profile of one more step (not part of the timed runs) ----
    stages, prof = None, None
    if not args.no_stages:
        model.timeline = []
        step_resident()
        torch.cuda.synchronize()
        tl = model.timeline
        model.timeline = None
        stages = {tl[i][0]: tl[i - 1][1].elapsed_time(tl[i][1]) for i in range(1, len(tl))}
        # per-kernel-class events: one context, one stream, eager launches (no graph, no SM partition)
        overlap = model.overlap_sam
        model.llm.use_cuda_graph, model.overlap_sam = False, False
        ctx.profile_begin()
        step_resident()
        torch.cuda.synchronize()
        prof = ctx.profile_end()
        model.llm.use_cuda_graph, model.overlap_sam = True, overlap

    if ws > 1:
        dist.barrier()
    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_src = load_peaks()

    images = B * ws * args.steps
    value = images / (ms_res / 1e3)
    e2e_v = images / (ms_e2e / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (seeded random weights, N(0,1) images, random prompt ids)",
            "config": config_dict(args, ws), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h_ids.numel() * 8 + h_img.numel() * 2 + h_sam.numel() * 2),
                    "d2h_bytes_per_step": int(h_out_ids.numel() * 8 + h_masks.numel() * 4 + h_gather.numel() * 4)},
            "masks_per_step": n_masks, "timed_region": info_res, "timed_region_e2e": info_e2e}
    parts = list(native.Partition._by_key.values())
    if model.overlap_sam and parts:
        line["overlap"] = {"sam_encoder_beside_decode": True, "sms_decode_lane": parts[0].sms[0],
                           "sms_sam_lane": parts[0].sms[1], "sam_blocks_on_lane": model.overlap_sam_blocks,
                           "how": "CUDA green contexts (ullava_partition)"}
    else:
        line["overlap"] = {"sam_encoder_beside_decode": False}
    if stages:
        line["stages_ms"] = {k: round(v, 3) for k, v in stages.items()}
    if prof:
        gt, gs = prof["gemm_tensor"], prof["gemm_stream"]
        line["kernel_classes"] = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                      "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1),
                                      "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)}
                                  for k, v in prof.items() if v["launches"]}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        a_t = gt["flops"] / max(gt["ms"], 1e-9) / 1e9
        line["roofline"] = {"kernel": "gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel (large-M GEMMs: ViT, prefill, SAM encoder + decoder)",
                            "bound": "tensor", "achieved": a_t, "peak": peaks["bf16_tflops_sustained"],
                            "unit": "TFLOP/s", "frac": a_t / peaks["bf16_tflops_sustained"],
                            "traffic": (traffic or {}).get("gemm_tensor"), "launches": gt["launches"],
                            "avg_launch_ms": gt["ms"] / max(gt["launches"], 1), "peak_source": peaks_src + ", sustained"}
        a_s = gs["bytes"] / max(gs["ms"], 1e-9) / 1e6
        line["roofline_decode"] = {"kernel": "gemm_stream_kernel (stream-K weight streaming, decode M <= 32)",
                                   "bound": "hbm", "achieved": a_s, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": a_s / peaks["hbm_gbs"], "traffic": (traffic or {}).get("gemm_stream"),
                                   "launches": gs["launches"], "avg_launch_ms": gs["ms"] / max(gs["launches"], 1),
                                   "peak_source": peaks_src}
        if stages and "decode" in stages:
            # the decode stage as a whole against the HBM roofline: per generated token every LLaMA weight (fused
            # qkv, o, gate/up, down, lm_head) and the K/V rows 0..pos of every layer are read once
            hs, ff, nl = LLM["hidden_size"], LLM["intermediate_size"], LLM["num_hidden_layers"]
            w_bytes = 2.0 * (nl * (4 * hs * hs + 3 * hs * ff) + VOCAB * hs)
            steps = max(args.new_tokens - 1, 0)          # the first token comes from the prefill
            kv_bytes = sum(B * nl * 2.0 * (P_LEN + t + 1) * hs * 2.0 for t in range(steps))
            dec_s = stages["decode"] / 1e3
            gbs = (steps * w_bytes + kv_bytes) / dec_s / 1e9 if dec_s > 0 else 0.0
            line["decode_stage"] = {"ms": round(stages["decode"], 3), "tokens": args.new_tokens, "gbs": round(gbs, 1),
                                    "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 3),
                                    "note": "in the CUDA graph (PDL overlap on): algorithmic weight + KV bytes of the "
                                            "whole stage / its duration; includes lm_head, sampling and the first-token step"}
        if stages and "prefill" in stages and "vit_projector_splice" in stages:
            fl = B * (FLOP_VIT + FLOP_PROJ + FLOP_PREFILL_LAST)
            sec = (stages["prefill"] + stages["vit_projector_splice"]) / 1e3
            line["vit_prefill"] = {"tflops": fl / sec / 1e12, "frac_of_burst_peak": fl / sec / 1e12 / peaks["bf16_tflops"],
                                   "frac_of_sustained_peak": fl / sec / 1e12 / peaks["bf16_tflops_sustained"],
                                   "batch": B, "ms": sec * 1e3}
    if not args.no_cpu_baseline and ws == 1:
        threads = use_all_host_threads()
        cpu = CpuSample()
        r = cpu.run_once(args.new_tokens)
        line["cpu_baseline"] = {"value": 1.0 / r["per_image_s"], "unit": UNIT, "cores": threads, "extrapolated": True,
                                "kind": "port", "sample": CpuSample.description, "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def load_peaks():
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        with open(pk_path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def run_stage_config(args, model, ctx, timed, tensors, rank, ws, local, dev):
    """BASELINE.json configs[1] / configs[2]: one stage of the path on its own, against the tensor roofline.
    c2: ViT-L/14-336 (UllavaCoreForCausalLM.encode_image) on 32 images; c3: LLaMA-7B prefill of 16 x 608 positions
    (forward(inputs_embeds=..., use_cache=True), last-position logits)."""
    import native
    h_ids, h_img, d_ids, d_img = tensors
    B = args.batch_per_gpu
    llm = model.llm
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if args.config == "c2":
        flops = B * FLOP_VIT
        h_out = torch.empty((B, N_PATCH, 1024), dtype=torch.bfloat16).pin_memory()

        def resident():
            flush.zero_()                      # ViT weights (0.6 GB) + activations: flush L2 between iterations
            return llm.encode_image(d_img)

        def e2e():
            out = llm.encode_image(h_img.to(dev, non_blocking=True))
            h_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        h2d, d2h = h_img.numel() * 2, h_out.numel() * 2
    else:
        flops = B * FLOP_PREFILL_LAST
        _, emb = llm.embed_images_videos(d_ids, d_img, None)        # ViT + projector + splice: outside the timed region
        emb = emb.contiguous()
        h_emb = emb.cpu().pin_memory()
        h_out = torch.empty((B, 1, VOCAB), dtype=torch.float32).pin_memory()

        def resident():
            return llm(inputs_embeds=emb, use_cache=True, logits_to_keep=1, logits_fp32=True, return_dict=True).logits

        def e2e():
            out = llm(inputs_embeds=h_emb.to(dev, non_blocking=True), use_cache=True, logits_to_keep=1, logits_fp32=True,
                      return_dict=True).logits
            h_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        h2d, d2h = h_emb.numel() * 2, h_out.numel() * 4
    sampler = ClockSampler(physical_gpu_index(local))
    if rank == 0 and not os.environ.get("ULLAVA_BENCH_NO_CLOCKS"):
        sampler.start()
    ms_res, launches, _ = timed(resident, args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(e2e, 2, args.steps)
    ctx.profile_begin()
    resident()
    torch.cuda.synchronize()
    prof = ctx.profile_end()
    if rank != 0:
        return
    peaks, peaks_src = load_peaks()
    images = B * ws * args.steps
    sec = ms_res / 1e3 / args.steps
    gt = prof["gemm_tensor"]
    a_t = gt["flops"] / max(gt["ms"], 1e-9) / 1e9
    line = {"metric": METRIC, "value": images / (ms_res / 1e3), "unit": UNIT, "n_gpus": ws, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (seeded random weights, N(0,1) images, random prompt ids)",
            "config": config_dict(args, ws), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": images / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "stage": {"tflops": flops / sec / 1e12, "frac_of_burst_peak": flops / sec / 1e12 / peaks["bf16_tflops"],
                      "frac_of_sustained_peak": flops / sec / 1e12 / peaks["bf16_tflops_sustained"],
                      "algorithmic_flop_per_step": flops},
            "roofline": {"kernel": "gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel", "bound": "tensor", "achieved": a_t,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": a_t / peaks["bf16_tflops_sustained"], "traffic": None, "launches": gt["launches"],
                         "avg_launch_ms": gt["ms"] / max(gt["launches"], 1), "peak_source": peaks_src + ", sustained"},
            "kernel_classes": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                   "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1),
                                   "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)}
                               for k, v in prof.items() if v["launches"]}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.batch_per_gpu is None:
        args.batch_per_gpu = {"c2": 32, "c3": 16, "c4": 8, "c5": 32}[args.config]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

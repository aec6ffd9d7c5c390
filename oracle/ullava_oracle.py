"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement (plain torch functional ops, no nn.Module, no transformers import) of the
u-LLaVA inference forward: CLIP ViT -> projector -> embedding splice -> LLaMA decoder (KV cached)
-> lm_head -> greedy loop -> [SEG] hidden state -> seg projector -> SAM prompt encoder / two-way
mask decoder -> mask post-processing.

It operates on a *state_dict with the reference's key names* (the same keys the reference's
UllavaForCausalLM.state_dict() has) plus a plain config dict, so the same tensors can be loaded
into (a) the real reference modules, (b) this oracle, (c) the B200 implementation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module -- as the checker or the timed CPU baseline, never as a fallback.

Pinning: the reference has no tests or golden vectors, and its arithmetic
for CLIP/LLaMA lives in the un-vendored dependency `transformers` (pinned 4.29.1 by the reference,
5.5.0 installed here).  tests/golden/make_golden.py therefore runs the REAL reference modules
(/root/reference + transformers 5.5.0 eager, fp32 CPU, seeded synthetic weights) in the build
container and commits their outputs; tests/test_oracle_golden.py checks this restatement against
them.  Parity is pinned to "reference code + transformers 5.5.0 eager", not to transformers 4.29.1.

Each function cites the reference (or `hf:` = transformers 5.5.0) lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# CLIP vision tower
# --------------------------------------------------------------------------------------------
def _act(x: torch.Tensor, name: str) -> torch.Tensor:
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)  # hf:activations.py QuickGELUActivation
    if name == "gelu":
        return F.gelu(x)
    if name == "relu":
        return F.relu(x)
    raise NotImplementedError(name)


def clip_vit_hidden(sd: SD, prefix: str, pixel_values: torch.Tensor, vcfg: dict, hidden_layer: int) -> torch.Tensor:
    """CLIPVisionModel(..., output_hidden_states=True).hidden_states[hidden_layer]
    (hf:models/clip/modeling_clip.py:138-218 embeddings, :282-385 encoder layer, :647-690 transformer).
    Layers that cannot influence the selected hidden state are not evaluated."""
    p = prefix + "vision_model."
    H, heads = vcfg["hidden_size"], vcfg["num_attention_heads"]
    hd = H // heads
    eps = vcfg.get("layer_norm_eps", 1e-5)
    patch = vcfg["patch_size"]
    n_layers = vcfg["num_hidden_layers"]
    idx = hidden_layer if hidden_layer >= 0 else n_layers + 1 + hidden_layer
    assert 0 <= idx <= n_layers
    B = pixel_values.shape[0]
    x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embedding.weight"], stride=patch)  # no bias (:148-154)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (H,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], eps)
    for l in range(idx):
        q = p + f"encoder.layers.{l}."
        r = x
        h = F.layer_norm(x, (H,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], eps)
        qq = F.linear(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"])
        kk = F.linear(h, sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"])
        vv = F.linear(h, sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"])
        S = h.shape[1]
        qq, kk, vv = (t.view(B, S, heads, hd).transpose(1, 2) for t in (qq, kk, vv))
        # eager_attention_forward :261-279 (softmax in fp32, cast back to the query dtype)
        w = torch.softmax((qq @ kk.transpose(-1, -2) * hd ** -0.5).float(), dim=-1).to(qq.dtype)
        a = (w @ vv).transpose(1, 2).reshape(B, S, H)
        x = r + F.linear(a, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        r = x
        h = F.layer_norm(x, (H,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], eps)
        h = _act(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"]), vcfg.get("hidden_act", "quick_gelu"))
        x = r + F.linear(h, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    return x


def encode_image(sd: SD, images: torch.Tensor, cfg: dict, prefix: str = "") -> torch.Tensor:
    """UllavaCoreForCausalLM.encode_image (models/ullava_core.py:146-158): drop CLS."""
    h = clip_vit_hidden(sd, prefix + "vision_encoder.", images, cfg["vision_config"], cfg["vision_hidden_layer"])
    return h[:, 1:]


def project(sd: SD, feats: torch.Tensor, cfg: dict, prefix: str = "") -> torch.Tensor:
    """build_vision_projector (models/ullava_core.py:117-129)."""
    p = prefix + "vision_projector."
    if cfg.get("projector_type", "mlp") == "mlp":
        return F.linear(feats, sd[p + "weight"], sd[p + "bias"])
    if cfg["projector_type"] == "mlp2x":
        h = F.gelu(F.linear(feats, sd[p + "0.weight"], sd[p + "0.bias"]))
        return F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"])
    raise NotImplementedError


def encode_video(sd: SD, videos: torch.Tensor, cfg: dict, prefix: str = "") -> torch.Tensor:
    """UllavaCoreForCausalLM.encode_video (models/ullava_core.py:160-180): [bs, C, T, H, W] -> per-frame patch
    features, temporal means (over patches) followed by spatial means (over frames): [bs, T + N, D]."""
    bs, c, t, h, w = videos.shape
    frames = videos.permute(0, 2, 1, 3, 4).reshape(bs * t, c, h, w)
    f = encode_image(sd, frames, cfg, prefix)
    f = f.view(bs, t, f.shape[1], f.shape[2])
    return torch.cat([f.mean(dim=2), f.mean(dim=1)], dim=1)


def embed_images(sd: SD, input_ids: torch.Tensor, images: Optional[torch.Tensor], cfg: dict,
                 prefix: str = "", videos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """embed_images_videos (models/ullava_core.py:182-277): token embeddings with the rows after <img_beg> /
    <vid_beg> overwritten by projected image / video features."""
    emb = sd[prefix + "model.embed_tokens.weight"][input_ids]
    ids = cfg["mm_token_ids"]
    feats = encode_image(sd, images, cfg, prefix) if images is not None else None
    vfeats = encode_video(sd, videos, cfg, prefix) if videos is not None else None
    out = []
    img_i = vid_i = 0
    for b in range(input_ids.shape[0]):
        cur = emb[b]
        n_s = int((input_ids[b] == ids["IMG_START"]).sum())
        n_e = int((input_ids[b] == ids["IMG_END"]).sum())
        n_vs = int((input_ids[b] == ids["VID_START"]).sum())
        n_ve = int((input_ids[b] == ids["VID_END"]).sum())
        assert n_s == n_e and n_vs == n_ve, "Number of image start and end tokens should be the same"
        if n_s == 0 and n_vs > 0:   # video branch (:248-269)
            pos = int(torch.where(input_ids[b] == ids["VID_START"])[0][0])
            f = project(sd, vfeats[vid_i], cfg, prefix)
            out.append(torch.cat([cur[:pos + 1], f, cur[pos + f.shape[0] + 1:]], dim=0))
            vid_i += 1
            continue
        if n_s == 0:
            out.append(cur)  # text only: dummy projector term is exactly zero (:213-220)
            continue
        pos = int(torch.where(input_ids[b] == ids["IMG_START"])[0][0])
        f = project(sd, feats[img_i], cfg, prefix)
        n = f.shape[0]
        out.append(torch.cat([cur[:pos + 1], f, cur[pos + n + 1:]], dim=0))  # (:234-245, both branches equal in value)
        img_i += 1
    return torch.stack(out, 0)


# --------------------------------------------------------------------------------------------
# LLaMA decoder
# --------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm (hf:models/llama/modeling_llama.py:52-69): statistics in fp32, one cast back to the input dtype
    before the weight multiply (no-ops for the fp32 oracle; they matter when the restatement is evaluated in 16 bits
    as the yardstick of what the reference's own bf16 / fp16 path does)."""
    xf = x.float()
    v = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(v + eps)).to(x.dtype)


def rope_tables(positions: torch.Tensor, hd: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding default rope (hf:...modeling_llama.py:74-135): returns cos, sin [S, hd/2]."""
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    fr = positions.float()[:, None] * inv.to(positions.device)[None, :]
    return fr.cos(), fr.sin()


def _rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_pos_emb with rotate_half (hf:...modeling_llama.py:137-168); x: [B, heads, S, hd]."""
    c = torch.cat([cos, cos], -1)[None, None]
    s = torch.cat([sin, sin], -1)[None, None]
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return x * c + rot * s


def llama_layers(sd: SD, x: torch.Tensor, cfg: dict, prefix: str = "", past: Optional[List] = None,
                 collect_hidden: bool = False, n_layers: Optional[int] = None):
    """LlamaModel.forward on inputs_embeds x [B,S,H] at positions len(past)..; causal eager attention
    (hf:models/llama/modeling_llama.py:199-291 attention, :171-184 MLP, :292-333 layer, :355-424 model).
    Returns (last_hidden_state after final norm, new past, hidden_states tuple or None)."""
    H, heads = cfg["hidden_size"], cfg["num_attention_heads"]
    assert cfg.get("num_key_value_heads", heads) == heads, "MHA only (LLaMA-7B / Vicuna-7B)"
    hd = H // heads
    eps = cfg.get("rms_norm_eps", 1e-6)
    L = cfg["num_hidden_layers"] if n_layers is None else n_layers
    B, S, _ = x.shape
    pos0 = 0 if past is None else past[0][0].shape[2]
    dev = x.device  # the oracle also runs in fp32 on the GPU for the full-size checks (tests/test_fullsize_gpu.py)
    cos, sin = rope_tables(torch.arange(pos0, pos0 + S, device=dev), hd, cfg.get("rope_theta", 10000.0))
    cos, sin = cos.to(x.dtype), sin.to(x.dtype)   # LlamaRotaryEmbedding returns cos / sin in the activation dtype
    qi = torch.arange(S, device=dev)[:, None] + pos0
    kj = torch.arange(pos0 + S, device=dev)[None, :]
    mask = torch.zeros(S, pos0 + S, device=dev).masked_fill(kj > qi, float("-inf"))
    new_past = []
    hiddens = [x] if collect_hidden else None
    for l in range(L):
        p = prefix + f"model.layers.{l}."
        r = x
        h = rms_norm(x, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, heads, hd).transpose(1, 2)
        q, k = _rope(q, cos, sin), _rope(k, cos, sin)
        if past is not None:
            k = torch.cat([past[l][0], k], dim=2)
            v = torch.cat([past[l][1], v], dim=2)
        new_past.append((k, v))
        # eager_attention_forward (:199-222): softmax in fp32, cast back to the query dtype
        w = torch.softmax((q @ k.transpose(-1, -2) * hd ** -0.5 + mask.to(q.dtype)).float(), dim=-1).to(q.dtype)
        a = (w @ v).transpose(1, 2).reshape(B, S, H)
        x = r + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        r = x
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps)
        h = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = r + F.linear(h, sd[p + "mlp.down_proj.weight"])
        if collect_hidden and l < L - 1:
            hiddens.append(x)
    x = rms_norm(x, sd[prefix + "model.norm.weight"], eps)
    if collect_hidden:
        hiddens.append(x)  # HF: the last entry of hidden_states is post-final-norm
    return x, new_past, (tuple(hiddens) if collect_hidden else None)


def core_forward(sd: SD, cfg: dict, input_ids: torch.Tensor, images: Optional[torch.Tensor] = None,
                 past=None, prefix: str = "", collect_hidden: bool = False, videos: Optional[torch.Tensor] = None):
    """UllavaCoreForCausalLM.forward (models/ullava_core.py:279-355) without the loss.
    Returns dict(logits [B,S,V], last_hidden [B,S,H], past, hidden_states)."""
    if input_ids.shape[1] == 1 and past is not None:
        x = sd[prefix + "model.embed_tokens.weight"][input_ids]  # decode step: vision tower skipped (:188-189)
    else:
        x = embed_images(sd, input_ids, images, cfg, prefix, videos)
    h, new_past, hs = llama_layers(sd, x, cfg, prefix, past, collect_hidden)
    logits = F.linear(h, sd[prefix + "lm_head.weight"])
    return {"logits": logits, "last_hidden": h, "past": new_past, "hidden_states": hs}


def greedy_generate(sd: SD, cfg: dict, input_ids: torch.Tensor, images: torch.Tensor, max_new_tokens: int,
                    prefix: str = "", eos_token_id: Optional[int] = None):
    """Greedy decode equivalent to self.llm.generate(do_sample=False, output_hidden_states=True,
    return_dict_in_generate=True) as used by UllavaForCausalLM.evaluate (models/ullava.py:349-365), KV cached.
    Returns (sequences [B, P+T], hidden [B, P+T-1, H]): the post-final-norm hidden state of every
    processed position -- what outputs.hidden_states[-1][-1] holds when the checkpoint's use_cache=False
    makes the last step re-forward the whole sequence."""
    out = core_forward(sd, cfg, input_ids, images, None, prefix)
    seqs = input_ids
    hid = [out["last_hidden"]]
    past = out["past"]
    nxt = out["logits"][:, -1].argmax(-1)
    margins = [_top2_margin(out["logits"][:, -1])]
    for t in range(max_new_tokens):
        seqs = torch.cat([seqs, nxt[:, None]], dim=1)
        if t == max_new_tokens - 1:
            break
        if eos_token_id is not None and bool((nxt == eos_token_id).all()):
            break
        out = core_forward(sd, cfg, nxt[:, None], None, past, prefix)
        past = out["past"]
        hid.append(out["last_hidden"])
        nxt = out["logits"][:, -1].argmax(-1)
        margins.append(_top2_margin(out["logits"][:, -1]))
    return seqs, torch.cat(hid, dim=1), torch.stack(margins, 1)


def _top2_margin(logits: torch.Tensor) -> torch.Tensor:
    t = logits.topk(2, dim=-1).values
    return t[:, 0] - t[:, 1]


# --------------------------------------------------------------------------------------------
# SAM prompt encoder / mask decoder / post-processing
# --------------------------------------------------------------------------------------------
def sam_dense_pe(sd: SD, prefix: str, size: int = 64) -> torch.Tensor:
    """PromptEncoder.get_dense_pe -> PositionEmbeddingRandom.forward
    (segment_anything/modeling/prompt_encoder.py:67-76,203-229): [1, 256, size, size]."""
    g = sd[prefix + "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones((size, size), dtype=g.dtype, device=g.device)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    c = torch.stack([x, y], -1)
    c = (2 * c - 1) @ g
    c = 2 * math.pi * c
    return torch.cat([c.sin(), c.cos()], -1).permute(2, 0, 1)[None]


def _sam_attn(sd: SD, p: str, q, k, v, heads: int = 8):
    """Attention (segment_anything/modeling/transformer.py:185-242)."""
    q = F.linear(q, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(k, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(v, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    b, n, c = q.shape

    def sep(t):
        return t.reshape(t.shape[0], t.shape[1], heads, c // heads).transpose(1, 2)

    q, k, v = sep(q), sep(k), sep(v)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(c // heads), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ln(sd: SD, p: str, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def sam_mask_decoder(sd: SD, prefix: str, image_embedding: torch.Tensor, text_embeds: torch.Tensor):
    """prompt_encoder(text_embeds) + MaskDecoder.predict_masks for ONE image and n prompts
    (prompt_encoder.py:164-186, mask_decoder.py:116-164, transformer.py:62-182).
    image_embedding [1,256,64,64]; text_embeds [n,256].  Returns (masks [n,4,256,256], iou [n,4])."""
    md = prefix + "mask_decoder."
    n = text_embeds.shape[0]
    sparse = text_embeds[:, None, :]
    dense = sd[prefix + "prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1)
    pe = sam_dense_pe(sd, prefix)
    tokens = torch.cat([sd[md + "iou_token.weight"], sd[md + "mask_tokens.weight"]], 0)[None].expand(n, -1, -1)
    tokens = torch.cat([tokens, sparse], dim=1)
    src = image_embedding.expand(n, -1, -1, -1) + dense
    b, c, h, w = src.shape
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pe.expand(n, -1, -1, -1).flatten(2).permute(0, 2, 1)
    queries, point = tokens, tokens
    for i in range(2):
        t = md + f"transformer.layers.{i}."
        if i == 0:
            queries = _sam_attn(sd, t + "self_attn.", queries, queries, queries)
        else:
            q = queries + point
            queries = queries + _sam_attn(sd, t + "self_attn.", q, q, queries)
        queries = _ln(sd, t + "norm1.", queries)
        q, k = queries + point, keys + key_pe
        queries = _ln(sd, t + "norm2.", queries + _sam_attn(sd, t + "cross_attn_token_to_image.", q, k, keys))
        m = F.linear(F.relu(F.linear(queries, sd[t + "mlp.lin1.weight"], sd[t + "mlp.lin1.bias"])),
                     sd[t + "mlp.lin2.weight"], sd[t + "mlp.lin2.bias"])
        queries = _ln(sd, t + "norm3.", queries + m)
        q, k = queries + point, keys + key_pe
        keys = _ln(sd, t + "norm4.", keys + _sam_attn(sd, t + "cross_attn_image_to_token.", k, q, queries))
    q, k = queries + point, keys + key_pe
    queries = queries + _sam_attn(sd, md + "transformer.final_attn_token_to_image.", q, k, keys)
    queries = _ln(sd, md + "transformer.norm_final_attn.", queries)
    iou_tok, mask_tok = queries[:, 0], queries[:, 1:5]
    src = keys.transpose(1, 2).reshape(b, c, h, w)
    u = F.conv_transpose2d(src, sd[md + "output_upscaling.0.weight"], sd[md + "output_upscaling.0.bias"], stride=2)
    mu = u.mean(1, keepdim=True)  # LayerNorm2d (common.py:31-43)
    var = (u - mu).pow(2).mean(1, keepdim=True)
    u = (u - mu) / torch.sqrt(var + 1e-6)
    u = sd[md + "output_upscaling.1.weight"][:, None, None] * u + sd[md + "output_upscaling.1.bias"][:, None, None]
    u = F.gelu(u)
    u = F.gelu(F.conv_transpose2d(u, sd[md + "output_upscaling.3.weight"], sd[md + "output_upscaling.3.bias"], stride=2))
    hyper = []
    for i in range(4):
        x = mask_tok[:, i]
        hp = md + f"output_hypernetworks_mlps.{i}.layers."
        x = F.relu(F.linear(x, sd[hp + "0.weight"], sd[hp + "0.bias"]))
        x = F.relu(F.linear(x, sd[hp + "1.weight"], sd[hp + "1.bias"]))
        hyper.append(F.linear(x, sd[hp + "2.weight"], sd[hp + "2.bias"]))
    hyper = torch.stack(hyper, 1)
    bb, cc, hh, ww = u.shape
    masks = (hyper @ u.view(bb, cc, hh * ww)).view(bb, 4, hh, ww)
    x = iou_tok
    ip = md + "iou_prediction_head.layers."
    x = F.relu(F.linear(x, sd[ip + "0.weight"], sd[ip + "0.bias"]))
    x = F.relu(F.linear(x, sd[ip + "1.weight"], sd[ip + "1.bias"]))
    iou = F.linear(x, sd[ip + "2.weight"], sd[ip + "2.bias"])
    return masks, iou


def postprocess_masks(masks: torch.Tensor, input_size: Sequence[int], original_size: Sequence[int],
                      img_size: int = 1024) -> torch.Tensor:
    """Sam.postprocess_masks (segment_anything/modeling/sam.py:137-172)."""
    m = F.interpolate(masks.float(), (img_size, img_size), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)


def seg_project(sd: SD, prefix: str, h: torch.Tensor) -> torch.Tensor:
    """seg_projector / det_projector: Linear-ReLU-Linear (models/ullava.py:83-132)."""
    x = F.relu(F.linear(h, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    return F.linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


def det_decode(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """det_decoder (models/ullava.py:96-102)."""
    x = F.relu(F.linear(x, sd["det_decoder.0.weight"], sd["det_decoder.0.bias"]))
    x = F.relu(F.linear(x, sd["det_decoder.2.weight"], sd["det_decoder.2.bias"]))
    return F.linear(x, sd["det_decoder.4.weight"], sd["det_decoder.4.bias"])


def masks_from_hidden(sd: SD, cfg: dict, token_ids: torch.Tensor, hidden: torch.Tensor,
                      image_embeddings: torch.Tensor, raw_size_list, resize_list):
    """Shared tail of UllavaForCausalLM.forward(inference=True) and .evaluate
    (models/ullava.py:168-256 and :364-432): token_ids [B,T]; hidden [B,T-1 or T,H] post-final-norm.
    The hidden state at position j is used when token j+1 is [SEG]/[LOC]."""
    B, T = token_ids.shape
    seg_mask = token_ids[:, 1:] == cfg["seg_token_idx"]
    loc_mask = token_ids[:, 1:] == cfg["loc_token_idx"]
    hid = hidden[:, : T - 1]
    pred_masks, pred_boxes, low_res = [], [], []
    for i in range(B):
        e = seg_project(sd, "seg_projector.", hid[i][seg_mask[i]])
        if e.shape[0] > 0:
            m, _ = sam_mask_decoder(sd, "visual_model.", image_embeddings[i:i + 1], e)
            m = m[:, 0:1]
        else:
            m = torch.zeros((0, 1, 256, 256), device=hidden.device)
        low_res.append(m)
        pm = postprocess_masks(m, resize_list[i], raw_size_list[i]) if e.shape[0] > 0 else \
            torch.zeros((0, 1) + tuple(raw_size_list[i]), device=hidden.device)
        pred_masks.append(pm[:, 0])
        le = seg_project(sd, "det_projector.", hid[i][loc_mask[i]])
        pred_boxes.append(det_decode(sd, le))
    return pred_masks, pred_boxes, low_res


# --------------------------------------------------------------------------------------------
# SAM ViT image encoder (on the path; segment_anything/modeling/image_encoder.py:110-426)
# --------------------------------------------------------------------------------------------
def _get_rel_pos(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """image_encoder.py:321-352 (no interpolation needed when the table already has 2*max-1 rows)."""
    max_rel = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel:
        rp = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel, mode="linear")
        rp = rp.reshape(-1, max_rel).permute(1, 0)
    else:
        rp = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rp[rel.long()]


def sam_image_encoder(sd: SD, prefix: str, x: torch.Tensor, ecfg: dict) -> torch.Tensor:
    """ImageEncoderViT.forward: [B,3,S,S] -> [B,256,S/16,S/16]."""
    p = prefix + "image_encoder."
    D, heads, depth = ecfg["embed_dim"], ecfg["num_heads"], ecfg["depth"]
    win, glob = ecfg["window_size"], set(ecfg["global_attn_indexes"])
    hd = D // heads
    x = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=ecfg["patch_size"])
    x = x.permute(0, 2, 3, 1) + sd[p + "pos_embed"]
    for i in range(depth):
        b = p + f"blocks.{i}."
        sc = x
        h = F.layer_norm(x, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)
        B, Hh, Ww, _ = h.shape
        w_ = 0 if i in glob else win
        if w_ > 0:  # window_partition :263-286
            ph, pw = (w_ - Hh % w_) % w_, (w_ - Ww % w_) % w_
            h = F.pad(h, (0, 0, 0, pw, 0, ph))
            Hp, Wp = Hh + ph, Ww + pw
            h = h.view(B, Hp // w_, w_, Wp // w_, w_, D).permute(0, 1, 3, 2, 4, 5).reshape(-1, w_, w_, D)
        Bw, hh, ww, _ = h.shape
        qkv = F.linear(h, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"])
        qkv = qkv.reshape(Bw, hh * ww, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.reshape(3, Bw * heads, hh * ww, hd).unbind(0)
        attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
        Rh = _get_rel_pos(hh, hh, sd[b + "attn.rel_pos_h"])  # add_decomposed_rel_pos :355-392
        Rw = _get_rel_pos(ww, ww, sd[b + "attn.rel_pos_w"])
        rq = q.reshape(Bw * heads, hh, ww, hd)
        rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
        rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
        attn = (attn.view(-1, hh, ww, hh, ww) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(
            -1, hh * ww, hh * ww)
        attn = attn.softmax(dim=-1)
        o = (attn @ v).view(Bw, heads, hh, ww, hd).permute(0, 2, 3, 1, 4).reshape(Bw, hh, ww, D)
        o = F.linear(o, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        if w_ > 0:  # window_unpartition :289-318
            o = o.view(B, Hp // w_, Wp // w_, w_, w_, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, D)
            o = o[:, :Hh, :Ww]
        x = sc + o
        h = F.layer_norm(x, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[b + "mlp.lin1.weight"], sd[b + "mlp.lin1.bias"])),
                     sd[b + "mlp.lin2.weight"], sd[b + "mlp.lin2.bias"])
        x = x + h
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + "neck.0.weight"])

    def ln2d(t, w, bb):
        mu = t.mean(1, keepdim=True)
        var = (t - mu).pow(2).mean(1, keepdim=True)
        return w[:, None, None] * ((t - mu) / torch.sqrt(var + 1e-6)) + bb[:, None, None]

    x = ln2d(x, sd[p + "neck.1.weight"], sd[p + "neck.1.bias"])
    x = F.conv2d(x, sd[p + "neck.2.weight"], padding=1)
    return ln2d(x, sd[p + "neck.3.weight"], sd[p + "neck.3.bias"])

"""Generates the golden fixtures in this directory by running the REAL reference
(/root/reference, transformers 5.5.0 eager attention, fp32, CPU) on seeded synthetic weights/inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Shims needed to run the unmodified reference here (stated with every result):
  * torch.Tensor.cuda -> identity (models/ullava.py hard-codes .cuda());
  * manual greedy loop over forward() with past_key_values (reference generate() does not run under
    transformers >= 5);
  * stub modules for peft / omegaconf / pycocotools imports that the reference's package imports pull in.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)          # reference `models`, `utils`
sys.path.insert(1, ROOT)         # oracle/, tests/
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.empty_cache = lambda: None

from oracle.synth import subsample, synth_normal, synth_state_dict  # noqa: E402
from tests import configs as C  # noqa: E402

import models as ref_models  # noqa: E402  (the reference package)
from models.segment_anything.build_sam import _build_sam as ref_build_sam  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)

# Weight seeds of the two tiny end-to-end fixtures.  They were picked (tests/golden/find_seed.py, a search over the
# oracle) so that EVERY greedy decision the tests compare has a top-2 margin above 0.16 = twice the bf16 logit
# tolerance: the exact-id / hidden-state / mask assertions of the generate() and evaluate() tests then always run
# (a near-tie could legitimately flip under a different reduction order and used to gate them off).
SEED_TINY_CORE = 14400   # min margin 0.163 over 2 x 8 golden tokens + the right-padded prompt variant (6 tokens)
SEED_TINY_FULL = 332     # min margin 0.252 over 2 x 8 tokens of the [SEG]/[LOC] prompt


def load_synth(module, seed=0):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = synth_state_dict(shapes, seed)
    missing = module.load_state_dict(sd, strict=True)
    return shapes, sd


def save(name, arrays, meta):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


def tiny_core():
    cfg = ref_models.UllavaCoreConfig(**C.TINY_LLM)
    cfg._attn_implementation = "eager"
    cfg.vision_config._attn_implementation = "eager"
    m = ref_models.UllavaCoreForCausalLM(cfg).eval().float()
    shapes, _ = load_synth(m, SEED_TINY_CORE)
    B = 2
    ids = C.tiny_prompt(B)
    images = synth_normal("images", (B, 3, 28, 28))
    out = m(input_ids=ids, images=images, output_hidden_states=True, use_cache=True, return_dict=True)
    feats = m.encode_image(images)
    # manual greedy loop (KV cached) -- 8 tokens
    seqs = ids
    past = out.past_key_values
    nxt = out.logits[:, -1].argmax(-1)
    hid = [out.hidden_states[-1]]
    margins = []
    for t in range(8):
        top2 = out.logits[:, -1].topk(2).values
        margins.append((top2[:, 0] - top2[:, 1]).numpy())
        seqs = torch.cat([seqs, nxt[:, None]], 1)
        if t == 7:
            break
        out = m(input_ids=nxt[:, None], past_key_values=past, use_cache=True, output_hidden_states=True,
                return_dict=True)
        past = out.past_key_values
        hid.append(out.hidden_states[-1])
        nxt = out.logits[:, -1].argmax(-1)
    first = m(input_ids=ids, images=images, output_hidden_states=True, return_dict=True)
    save("tiny_core", dict(input_ids=ids.numpy(), image_features=feats.numpy(), logits=first.logits.numpy(),
                           last_hidden=first.hidden_states[-1].numpy(), hidden1=first.hidden_states[1].numpy(),
                           greedy=seqs.numpy(), greedy_hidden=torch.cat(hid, 1).numpy(),
                           margins=np.stack(margins, 1)),
         dict(shapes={k: list(v) for k, v in shapes.items()}, seed=SEED_TINY_CORE, config="TINY_LLM",
              min_margin=float(np.stack(margins, 1).min()),
              note="reference UllavaCoreForCausalLM, transformers 5.5.0 eager, fp32 CPU"))


def tiny_full():
    llm = dict(C.TINY_LLM)
    cfg = ref_models.UllavaConfig(llm_config=llm, seg_token_idx=C.SEG_ID, loc_token_idx=C.LOC_ID)
    cfg.llm_config._attn_implementation = "eager"
    cfg.llm_config.vision_config._attn_implementation = "eager"
    e = C.TINY_SAM_ENCODER
    import models.ullava as ref_ullava
    ref_ullava.build_sam_vit_h = lambda checkpoint=None: ref_build_sam(e["embed_dim"], e["depth"], e["num_heads"],
                                                                        e["global_attn_indexes"])
    m = ref_models.UllavaForCausalLM(cfg).eval().float()
    shapes, _ = load_synth(m, SEED_TINY_FULL)
    B = 2
    ids = C.tiny_prompt(B, seg_loc=True)
    images = synth_normal("images", (B, 3, 28, 28))
    images_sam = synth_normal("images_sam", (B, 3, 1024, 1024))
    sizes = [(40, 56), (33, 47)]
    resizes = [(731, 1024), (719, 1024)]
    emb = m.get_visual_embs(images_sam)
    out = m(images_sam=images_sam, images=images, input_ids=ids, labels=ids.clone(),
            attention_mask=torch.ones_like(ids).bool(), mask_list=[None] * B, size_list=sizes, resize_list=resizes,
            bbox_list=[None] * B, inference=True)
    arrays = dict(input_ids=ids.numpy(), logits=out["logits"].numpy(), sam_embeddings_sub=subsample(emb, 32768)[0].numpy())
    for i in range(B):
        arrays[f"pred_mask_{i}"] = out["pred_masks"][i].numpy()
        arrays[f"pred_box_{i}"] = out["pred_boxes"][i].numpy()
    save("tiny_full", arrays,
         dict(shapes={k: list(v) for k, v in shapes.items()}, seed=SEED_TINY_FULL, sizes=sizes, resizes=resizes,
              note="reference UllavaForCausalLM.forward(inference=True); SAM image encoder reduced to 2 blocks, "
                   "prompt encoder / mask decoder at build_sam geometry; .cuda() shim"))


def tiny_full_coherent():
    """The tiny_full model with the coherent-mask overrides (oracle/synth.py) on a piecewise-constant SAM image: the REAL
    reference's forward(inference=True) masks at 336 x 336 / 300 x 420, for the IoU >= 0.999 gate."""
    from oracle.synth import coherent_clip_images, coherent_image, coherent_overrides
    llm = dict(C.TINY_LLM)
    cfg = ref_models.UllavaConfig(llm_config=llm, seg_token_idx=C.SEG_ID, loc_token_idx=C.LOC_ID)
    cfg.llm_config._attn_implementation = "eager"
    cfg.llm_config.vision_config._attn_implementation = "eager"
    e = C.TINY_SAM_ENCODER
    import models.ullava as ref_ullava
    ref_ullava.build_sam_vit_h = lambda checkpoint=None: ref_build_sam(e["embed_dim"], e["depth"], e["num_heads"],
                                                                        e["global_attn_indexes"])
    m = ref_models.UllavaForCausalLM(cfg).eval().float()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(coherent_overrides(synth_state_dict(shapes, SEED_TINY_FULL)), strict=True)
    B = 2
    ids = C.tiny_prompt(B, seg_loc=True)
    images = coherent_clip_images(B)
    images_sam = coherent_image(B)
    sizes = [(336, 336), (300, 420)]
    resizes = [(1024, 1024), (731, 1024)]
    out = m(images_sam=images_sam, images=images, input_ids=ids, labels=ids.clone(),
            attention_mask=torch.ones_like(ids).bool(), mask_list=[None] * B, size_list=sizes, resize_list=resizes,
            bbox_list=[None] * B, inference=True)
    arrays, stats = {}, []
    for i in range(B):
        pm = out["pred_masks"][i]
        arrays[f"mask_bits_{i}"] = np.packbits((pm > 0).numpy().reshape(pm.shape[0], -1), axis=1)
        arrays[f"mask_sub_{i}"] = pm.reshape(pm.shape[0], -1)[:, ::37].numpy().astype(np.float32)
        stats.append([[float((x > 0).float().mean()), float(x.abs().max())] for x in pm])
    save("tiny_full_coherent", arrays,
         dict(seed=SEED_TINY_FULL, sizes=sizes, resizes=resizes, sub_stride=37, positive_share_and_max=stats,
              note="reference UllavaForCausalLM.forward(inference=True) with oracle.synth.coherent_overrides / "
                   "coherent_image: thresholded masks (packbits) + every 37th logit"))


def sam_decoder_full():
    sam = ref_build_sam(64, 1, 2, [0]).float()
    shapes, sd = load_synth(sam, seed=1)
    n = 3
    emb = synth_normal("sam_emb", (1, 256, 64, 64), seed=1)
    text = synth_normal("sam_text", (n, 1, 256), seed=1)
    sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
    pe = sam.prompt_encoder.get_dense_pe()
    masks, iou = sam.mask_decoder.predict_masks(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                                dense_prompt_embeddings=dense)
    post = sam.postprocess_masks(masks[:, 0:1], input_size=(768, 1024), original_size=(120, 160))
    m_sub, stride = subsample(masks, 65536)
    save("sam_decoder", dict(masks_sub=m_sub.numpy(), iou=iou.numpy(), dense_pe_sub=subsample(pe, 8192)[0].numpy(),
                             post=post.numpy(), masks_mean=masks.mean((2, 3)).numpy(),
                             masks_absmean=masks.abs().mean((2, 3)).numpy()),
         dict(shapes={k: list(v) for k, v in shapes.items() if not k.startswith("image_encoder")}, seed=1,
              masks_stride=stride, note="reference Sam prompt_encoder(text_embeds) + MaskDecoder.predict_masks + "
                                        "postprocess_masks, fp32 CPU"))


def clip_layer_full():
    """One full-width CLIP ViT-L/14-336 encoder layer (hidden 1024, 16 heads, 577 tokens)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    vc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=1, num_attention_heads=16,
                          image_size=336, patch_size=14, hidden_act="quick_gelu")
    vc._attn_implementation = "eager"
    m = CLIPVisionModel(vc).eval().float()
    shapes, _ = load_synth(m, seed=2)
    px = synth_normal("clip_px", (1, 3, 336, 336), seed=2)
    out = m(px, output_hidden_states=True)
    h1 = out.hidden_states[1]
    save("clip_layer_full", dict(hidden1_sub=subsample(h1, 32768)[0].numpy(), hidden0_sub=subsample(out.hidden_states[0], 32768)[0].numpy()),
         dict(shapes={k: list(v) for k, v in shapes.items()}, seed=2, vision_config=vc.to_dict()["hidden_size"],
              note="HF CLIPVisionModel 1 layer at ViT-L/14-336 width, eager, fp32 CPU"))


def llama_layer_full():
    """One full-width LLaMA-7B decoder layer at L=608 (hidden 4096, 32 heads, ffn 11008)."""
    from transformers import LlamaConfig, LlamaModel
    lc = LlamaConfig(vocab_size=64, hidden_size=4096, intermediate_size=11008, num_hidden_layers=1,
                     num_attention_heads=32, num_key_value_heads=32, rms_norm_eps=1e-6)
    lc._attn_implementation = "eager"
    m = LlamaModel(lc).eval().float()
    shapes, _ = load_synth(m, seed=3)
    x = synth_normal("llama_x", (1, 608, 4096), seed=3)
    out = m(inputs_embeds=x, output_hidden_states=True, use_cache=False)
    save("llama_layer_full", dict(last_sub=subsample(out.last_hidden_state, 32768)[0].numpy()),
         dict(shapes={k: list(v) for k, v in shapes.items()}, seed=3,
              note="HF LlamaModel 1 layer at LLaMA-7B width, L=608, eager, fp32 CPU (post final norm)"))


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny_core", "tiny_full", "sam_decoder_full", "clip_layer_full", "llama_layer_full"]
    for w in which:
        globals()[w]()

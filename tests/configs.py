"""Shared test configurations (tiny shapes for CI, full-width single-layer shapes for kernel-size parity)."""
MM_IDS = dict(IMG_PATCH=301, VID_PATCH=302, IMG_START=303, IMG_END=304, VID_START=305, VID_END=306)
SEG_ID, LOC_ID = 307, 308

TINY_VISION = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=2, image_size=28,
                   patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
TINY_LLM = dict(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                num_key_value_heads=2, rms_norm_eps=1e-6, max_position_embeddings=512,
                vision_config=TINY_VISION, vision_hidden_layer=-2, projector_type="mlp", mm_token_ids=MM_IDS,
                bos_token_id=1, eos_token_id=2, pad_token_id=0)
# SAM with a 2-block image encoder; prompt encoder / mask decoder are the real build_sam geometry
TINY_SAM_ENCODER = dict(embed_dim=64, depth=2, num_heads=2, global_attn_indexes=[1], window_size=14, patch_size=16)


def tiny_prompt(batch: int, n_patch: int = 4, seg_loc: bool = False):
    """input_ids [batch, L]: BOS + 3 text + IMG_START + n_patch*IMG_PATCH + IMG_END + tail."""
    import torch
    from oracle.synth import synth_ids
    head = synth_ids("head", (batch, 3), 3, 300)
    tail = synth_ids("tail", (batch, 7), 3, 300)
    if seg_loc:
        tail[:, 2] = SEG_ID
        tail[:, 4] = LOC_ID
        if batch > 1:
            tail[1, 5] = SEG_ID  # second sample: two masks
    bos = torch.full((batch, 1), 1, dtype=torch.int64)
    img = torch.tensor([MM_IDS["IMG_START"]] + [MM_IDS["IMG_PATCH"]] * n_patch + [MM_IDS["IMG_END"]],
                       dtype=torch.int64)[None].expand(batch, -1)
    return torch.cat([bos, head, img, tail], dim=1)

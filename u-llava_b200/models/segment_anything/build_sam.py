"""SAM factory functions with the hyper-parameters of the reference
(/root/reference/models/segment_anything/build_sam.py:15-108)."""
import torch

from .modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer

_PROMPT_DIM, _IMG, _PATCH = 256, 1024, 16


def _build_sam(encoder_embed_dim, encoder_depth, encoder_num_heads, encoder_global_attn_indexes, checkpoint=None):
    grid = _IMG // _PATCH
    sam = Sam(
        image_encoder=ImageEncoderViT(depth=encoder_depth, embed_dim=encoder_embed_dim, img_size=_IMG, mlp_ratio=4,
                                      num_heads=encoder_num_heads, patch_size=_PATCH, qkv_bias=True, use_rel_pos=True,
                                      global_attn_indexes=encoder_global_attn_indexes, window_size=14,
                                      out_chans=_PROMPT_DIM, norm_eps=1e-6),
        prompt_encoder=PromptEncoder(embed_dim=_PROMPT_DIM, image_embedding_size=(grid, grid),
                                     input_image_size=(_IMG, _IMG), mask_in_chans=16),
        mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                 transformer=TwoWayTransformer(depth=2, embedding_dim=_PROMPT_DIM, mlp_dim=2048,
                                                               num_heads=8),
                                 transformer_dim=_PROMPT_DIM, iou_head_depth=3, iou_head_hidden_dim=256),
        pixel_mean=[123.675, 116.28, 103.53], pixel_std=[58.395, 57.12, 57.375])
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            state_dict = torch.load(f)
        sam.load_state_dict(state_dict, strict=False)
    return sam


def build_sam_vit_h(checkpoint=None):
    return _build_sam(1280, 32, 16, [7, 15, 23, 31], checkpoint)


def build_sam_vit_l(checkpoint=None):
    return _build_sam(1024, 24, 16, [5, 11, 17, 23], checkpoint)


def build_sam_vit_b(checkpoint=None):
    return _build_sam(768, 12, 12, [2, 5, 8, 11], checkpoint)


build_sam = build_sam_vit_h
sam_model_registry = {"default": build_sam_vit_h, "vit_h": build_sam_vit_h, "vit_l": build_sam_vit_l,
                      "vit_b": build_sam_vit_b}

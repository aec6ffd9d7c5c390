"""B200-native counterpart of the reference's vendored models/segment_anything package
(only what the u-LLaVA path touches: build_sam_*, Sam, ImageEncoderViT, PromptEncoder, MaskDecoder,
TwoWayTransformer).  SamPredictor / automatic mask generation / ONNX export are out of scope."""
from .build_sam import build_sam, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l, sam_model_registry

__all__ = ["build_sam", "build_sam_vit_h", "build_sam_vit_l", "build_sam_vit_b", "sam_model_registry"]

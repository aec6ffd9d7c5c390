"""Sam container (reference: segment_anything/modeling/sam.py).  postprocess_masks runs the fused
double-bilinear kernel (ullava_sam_postprocess) instead of materialising the 1024x1024 intermediate."""
from typing import Tuple

import torch
import torch.nn as nn

import native


class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder, prompt_encoder, mask_decoder, pixel_mean=(123.675, 116.28, 103.53),
                 pixel_std=(58.395, 57.12, 57.375)):
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self):
        return self.pixel_mean.device

    def postprocess_masks(self, masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...],
                          pack_bits: bool = False):
        """[n,C,h,w] low-res logits -> [n,C,H,W] fp32 logits at the original image size."""
        n, c, h, w = masks.shape
        if masks.device.type != "cuda":
            raise RuntimeError("Sam.postprocess_masks (B200 build) runs on a CUDA sm_100 device only")
        if h != w:
            raise NotImplementedError("square low-res masks only")
        if masks.dtype == torch.float32:
            masks = masks.to(torch.bfloat16)  # the decoder emits 16-bit logits; fp32 input is never on the path
        ctx = native.Context.get(masks.device)
        if c == 1 and masks.stride(3) == 1 and masks.stride(2) == w:
            m, mstride = masks, masks.stride(0)  # e.g. the [:, 0:1] slice of the decoder's [n,4,h,w] output
        else:
            m, mstride = masks.contiguous(), h * w
        oh, ow = int(original_size[0]), int(original_size[1])
        out, bits = ctx.sam_postprocess(m, mstride, n * c, h, self.image_encoder.img_size,
                                        (int(input_size[0]), int(input_size[1])), (oh, ow), pack_bits=pack_bits)
        out = out.view(n, c, oh, ow)
        return (out, bits) if pack_bits else out

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        x = (x - self.pixel_mean) / self.pixel_std
        h, w = x.shape[-2:]
        s = self.image_encoder.img_size
        return nn.functional.pad(x, (0, s - w, 0, s - h))

    def forward(self, batched_input, multimask_output):
        raise NotImplementedError("Sam.forward (point/box prompting) is outside the u-LLaVA path")

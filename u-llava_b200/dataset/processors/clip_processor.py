"""CLIPProcessor of /root/reference/dataset/processors/clip_processor.py:23-101 on the device.

__call__(item) takes the decoded RGB image (numpy uint8 HWC, a PIL image, or a uint8 HWC CUDA tensor) and returns the
CLIP pixel_values [3, size, size] in `dtype` on the GPU: optional white square padding ('pad' aspect ratio), Pillow
BICUBIC resize of the shortest edge, center crop, rescale and normalise -- the arithmetic of
CLIPImageProcessor.preprocess (transformers 4.29.1) reproduced bit for bit by csrc/preprocess.cu.
Only the host->device copy of the uint8 image and the padding fill are done with torch (memory plumbing)."""
import numpy as np
import torch

import native

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _register(cls):
    """registry.register_processor('clip_image') (reference clip_processor.py:23) when the reference's utils package
    is importable, so that tasks/base_task.py:49 builds THIS processor from the eval / train configs."""
    try:
        from utils.registry import registry
        return registry.register_processor('clip_image')(cls)
    except Exception:
        return cls


@_register
class CLIPProcessor:
    def __init__(self, checkpoint_path=None, aspect_ratio=None, size: int = 336, image_mean=OPENAI_CLIP_MEAN,
                 image_std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255, dtype=torch.float32, device=None):
        """checkpoint_path: directory with a preprocessor_config.json (size / crop_size / mean / std are read from it
        when present, like CLIPImageProcessor.from_pretrained); aspect_ratio: 'pad', 'keep' or None.
        dtype: fp32 by default, like the reference (whose callers then do `.cuda().to(model dtype)`,
        inference_ullava.py:78); pass the model dtype to get that single rounding done by the kernel instead."""
        self.aspect_ratio = aspect_ratio
        self.size, self.mean, self.std, self.rescale = int(size), tuple(image_mean), tuple(image_std), rescale_factor
        if checkpoint_path is not None:
            import json
            import os
            cfg_path = os.path.join(checkpoint_path, "preprocessor_config.json")
            if os.path.exists(cfg_path):
                with open(cfg_path) as f:
                    cfg = json.load(f)
                sz = cfg.get("size", self.size)
                self.size = int(sz["shortest_edge"] if isinstance(sz, dict) else sz)
                self.mean = tuple(cfg.get("image_mean", self.mean))
                self.std = tuple(cfg.get("image_std", self.std))
                self.rescale = cfg.get("rescale_factor", self.rescale)
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    @staticmethod
    def resize_shape(h: int, w: int, size: int):
        """get_resize_output_image_size(shortest_edge, default_to_square=False) of transformers."""
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = size, int(size * long / short)
        return (new_long, new_short) if w <= h else (new_short, new_long)

    def _to_device(self, item) -> torch.Tensor:
        if isinstance(item, torch.Tensor):
            t = item
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(item)))
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "expected an RGB uint8 HWC image"
        return t.to(self.device, non_blocking=True).contiguous()

    def pad(self, img: torch.Tensor, background_color=(255, 255, 255)) -> torch.Tensor:
        """pad_pil / pad_cv2 (clip_processor.py:36-79): centre the image on a square background."""
        h, w = img.shape[:2]
        if h == w:
            return img
        size = max(h, w)
        out = torch.empty((size, size, 3), dtype=torch.uint8, device=img.device)
        out[:] = torch.tensor(background_color, dtype=torch.uint8, device=img.device)
        if w > h:
            y0 = (w - h) // 2
            out[y0:y0 + h, :w] = img
        else:
            x0 = (h - w) // 2
            out[:h, x0:x0 + w] = img
        return out

    def __call__(self, item) -> torch.Tensor:
        ctx = native.Context.get(self.device)
        img = self._to_device(item)
        if self.aspect_ratio == 'pad':
            img = self.pad(img)
        h, w = int(img.shape[0]), int(img.shape[1])
        oh, ow = self.resize_shape(h, w, self.size)
        if min(oh, ow) < self.size:
            raise ValueError(f"image {h}x{w} resizes to {oh}x{ow}: smaller than the {self.size} crop")
        resized = ctx.resize_u8(img, oh, ow, bicubic=True)
        top, left = (oh - self.size) // 2, (ow - self.size) // 2
        return ctx.clip_preprocess(resized, top, left, self.size, self.mean, self.std, self.dtype, self.rescale)

    @classmethod
    def from_config(cls, cfg=None):
        cfg = cfg or {}
        return cls(checkpoint_path=cfg.get("path", None), aspect_ratio=cfg.get("aspect_ratio", None))

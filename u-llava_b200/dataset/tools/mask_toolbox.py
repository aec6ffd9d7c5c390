"""SegToolBox and DetToolBox of /root/reference/dataset/tools/mask_toolbox.py:8-28,31-90 (+ ResizeLongestSide.apply_image,
models/segment_anything/utils/transforms.py:29-37,61-70) on the device.

apply_image: Pillow BILINEAR resize of the longest side to 1024 (uint8 HWC in, uint8 HWC out, on the GPU);
preprocess : (x - mean) / std in fp32 and zero padding to [3, 1024, 1024], in `dtype`.
Both are bit-exact with the reference's host code (csrc/preprocess.cu)."""
import numpy as np
import torch

import native


class SegToolBox:
    def __init__(self, dtype=torch.float32, device=None):
        """dtype: fp32 by default, like the reference's preprocess (callers cast with `.to(model dtype)`,
        inference_ullava.py:85); pass the model dtype to let the kernel do that single rounding."""
        self.sam_mean = (123.675, 116.28, 103.53)
        self.sam_std = (58.395, 57.12, 57.375)
        self.sam_size = 1024
        self.dtype = dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int):
        scale = long_side_length * 1.0 / max(oldh, oldw)
        newh, neww = oldh * scale, oldw * scale
        return int(newh + 0.5), int(neww + 0.5)

    def _to_device(self, image) -> torch.Tensor:
        t = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(image)))
        assert t.dtype == torch.uint8 and t.dim() == 3, "expected a uint8 image"
        return t.to(self.device, non_blocking=True)

    def apply_image(self, image) -> torch.Tensor:
        """uint8 HWC (numpy / PIL / tensor) -> uint8 HWC CUDA tensor with the longest side at 1024."""
        img = self._to_device(image).contiguous()
        oh, ow = self.get_preprocess_shape(int(img.shape[0]), int(img.shape[1]), self.sam_size)
        return native.Context.get(self.device).resize_u8(img, oh, ow, bicubic=False)

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        """Normalise and pad.  Accepts the reference's CHW uint8 tensor (torch.from_numpy(img).permute(2, 0, 1)) or
        the HWC tensor apply_image returns."""
        x = self._to_device(x)
        if x.shape[0] == 3 and x.shape[-1] != 3:
            x = x.permute(1, 2, 0)
        x = x.contiguous()
        return native.Context.get(self.device).sam_preprocess(x, self.sam_size, self.sam_mean, self.sam_std, self.dtype)

    def __call__(self, image):
        """image -> (images_sam [3, 1024, 1024], resize (h, w)): the two values inference_ullava.py:82-85 needs."""
        r = self.apply_image(image)
        return self.preprocess(r), (int(r.shape[0]), int(r.shape[1]))


class DetToolBox:
    """Box bookkeeping of the reference (mask_toolbox.py:31-90): boxes are normalised on the image padded to a square
    (the CLIP 'pad' aspect ratio).  Host-side scalar arithmetic on four numbers; nothing here touches the device."""

    @staticmethod
    def get_pad_length(width, height):
        if width > height:
            return 0, (width - height) / 2.0
        return (height - width) / 2.0, 0

    @staticmethod
    def xywh2xyxy(xywh):
        x, y, w, h = xywh
        return [x, y, x + w, y + h]

    def pad_normalize_xyxy(self, xyxy, width, height):
        side = max(width, height)
        pad_x, pad_y = self.get_pad_length(width, height)
        x0, y0, x1, y1 = xyxy
        return [(x0 + pad_x) / side, (y0 + pad_y) / side, (x1 + pad_x) / side, (y1 + pad_y) / side]

    def denormalize_padded_xyxy(self, normalized_xyxy, width, height):
        side = max(width, height)
        pad_x, pad_y = self.get_pad_length(width, height)
        x0, y0, x1, y1 = normalized_xyxy
        return [x0 * side - pad_x, y0 * side - pad_y, x1 * side - pad_x, y1 * side - pad_y]

    @staticmethod
    def mask2bbox(binary_mask):
        """[x0, y0, x1, y1] (inclusive) of the non-zero pixels, what pycocotools' encode + toBbox give the reference."""
        m = np.asarray(binary_mask) != 0
        if not m.any():
            return [0.0, 0.0, -1.0, -1.0]
        ys, xs = np.nonzero(m)
        return [float(xs.min()), float(ys.min()), float(xs.max()), float(ys.max())]

"""ResizeLongestSide of the reference (models/segment_anything/utils/transforms.py:17-102) on the B200 path.

apply_image keeps the reference contract (uint8 HWC numpy array in, uint8 HWC numpy array out) but resamples on the
device with the Pillow-exact kernel (ullava_resize_u8, BILINEAR) -- the reference does
np.array(resize(to_pil_image(image), target_size)); callers that want the tensor to stay on the GPU use
dataset.tools.mask_toolbox.SegToolBox.apply_image.  The coordinate / box helpers are host arithmetic on a few numbers.
The *_torch image variant (F.interpolate, "may not exactly match apply_image" in the reference's own words) is not on
the u-LLaVA path and raises."""
from copy import deepcopy
from typing import Tuple

import numpy as np
import torch

import native


class ResizeLongestSide:
    def __init__(self, target_length: int) -> None:
        self.target_length = target_length

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        """uint8 HxWxC -> uint8 with the longest side at target_length (Pillow BILINEAR, bit-exact)."""
        oh, ow = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        dev = torch.device("cuda", torch.cuda.current_device())
        src = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
        return native.Context.get(dev).resize_u8(src, oh, ow, bicubic=False).cpu().numpy()

    def apply_coords(self, coords: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(original_size[0], original_size[1], self.target_length)
        coords = deepcopy(coords).astype(float)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes(self, boxes: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        return self.apply_coords(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)

    def apply_image_torch(self, image: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("apply_image_torch (F.interpolate resize) is not on the u-LLaVA path; apply_image / "
                                  "SegToolBox.apply_image give the resize the model was trained with")

    def apply_coords_torch(self, coords: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(original_size[0], original_size[1], self.target_length)
        coords = deepcopy(coords).to(torch.float)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes_torch(self, boxes: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        return self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        scale = long_side_length * 1.0 / max(oldh, oldw)
        newh, neww = oldh * scale, oldw * scale
        return int(newh + 0.5), int(neww + 0.5)

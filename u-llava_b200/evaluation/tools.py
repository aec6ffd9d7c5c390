"""Meters of u-LLaVA evaluation on the device.

Same names as /root/reference/evaluation/tools.py (bbox_iou, intersectionAndUnionGPU, Summary, to_cuda, dict_to_cuda,
AverageMeter) so that `from evaluation.tools import ...` in evaluation/eval_ullava.py resolves here, plus `SegMeter`:
the three AverageMeters of validate() (evaluation/eval_ullava.py:41-102) kept in device memory.  The reference
copies two histograms to the host per sentence; here one launch counts every mask of an image, one more folds a whole
batch of images into the running fp64 meters, and the only device->host copy is result().

Everything goes through the C ABI (native.py); there is no torch fallback.
"""
from enum import Enum
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

import native


def _labels(t: torch.Tensor) -> torch.Tensor:
    """int32 / uint8 label tensors are consumed as they are; anything else (the reference casts the 16-bit gt masks
    with .int(), evaluation/eval_ullava.py:62) is converted once."""
    t = t.contiguous()
    return t if t.dtype in (torch.int32, torch.uint8) else t.to(torch.int32)


def bbox_iou(pred_boxes: torch.Tensor, target_boxes: torch.Tensor):
    """evaluation/tools.py:13-26: torchvision box_iou(pred * 1000, target * 1000) diagonal -> accuracy / miou / num."""
    ctx = native.Context.get(pred_boxes.device)
    ious = ctx.box_iou_diag(pred_boxes, target_boxes.to(pred_boxes.dtype))
    num = len(target_boxes)
    return {'accuracy': 1.0 * (ious > 0.5).sum().item() / num, 'miou': ious.mean().item(), "num": num}


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """evaluation/tools.py:29-41 for K = 2 (the only value the reference passes).  Unlike the reference, `output` is
    not modified in place.  Returns (area_intersection, area_union, area_target), int32 tensors of K entries."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    if K != 2:
        raise NotImplementedError("intersectionAndUnionGPU: the device kernel counts K = 2 classes")
    ctx = native.Context.get(output.device)
    pred = output.reshape(1, -1)
    pred = pred.contiguous() if pred.dtype in (torch.int32, torch.uint8) else pred.to(torch.int32)
    c = ctx.mask_iou_counts(pred, _labels(target.reshape(1, -1)), ignore_index)[0]
    return c[0:2], c[2:4], c[4:6]


class Summary(Enum):
    NONE = 0
    AVERAGE = 1
    SUM = 2
    COUNT = 3


def to_cuda(tensor, torch_type=torch.float32):
    return tensor.to(torch_type).cuda(non_blocking=True)


def dict_to_cuda(input_dict, torch_type=torch.float32):
    """evaluation/tools.py:55-67 (ids / labels / attention_mask keep their dtype, every other tensor is cast)."""
    for k, v in input_dict.items():
        if k in ['input_ids', 'labels', 'attention_mask']:
            input_dict[k] = v.cuda(non_blocking=True)
        elif isinstance(input_dict[k], torch.Tensor):
            input_dict[k] = v.cuda(non_blocking=True).to(torch_type)
        elif isinstance(input_dict[k], list) and len(input_dict[k]) > 0 and isinstance(input_dict[k][0], torch.Tensor):
            input_dict[k] = [ele.cuda(non_blocking=True).to(torch_type) for ele in v]
    return input_dict


class AverageMeter(object):
    """Host-side meter with the reference's interface (evaluation/tools.py:70-115); kept for scripts that use it
    directly.  validate() below uses SegMeter instead."""

    def __init__(self, name, fmt=":f", summary_type=Summary.AVERAGE):
        self.name, self.fmt, self.summary_type = name, fmt, summary_type
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def all_reduce(self):
        device = "cuda" if torch.cuda.is_available() else "cpu"
        if isinstance(self.sum, np.ndarray):
            total = torch.tensor(self.sum.tolist() + [self.count], dtype=torch.float32, device=device)
        else:
            total = torch.tensor([self.sum, self.count], dtype=torch.float32, device=device)
        dist.all_reduce(total, dist.ReduceOp.SUM, async_op=False)
        if total.shape[0] > 2:
            self.sum, self.count = total[:-1].cpu().numpy(), total[-1].cpu().item()
        else:
            self.sum, self.count = total.tolist()
        self.avg = self.sum / (self.count + 1e-5)

    def __str__(self):
        fmtstr = "{name} {val" + self.fmt + "} ({avg" + self.fmt + "})"
        return fmtstr.format(**self.__dict__)


class SegMeter:
    """intersection / union / gIoU / Prec@0.5 meters of validate() in device memory (fp64).

    update() takes what UllavaForCausalLM.forward(inference=True) returns for a BATCH of images:
    pred_masks[i] fp32 logits [n_i, H_i, W_i], gt_masks[i] [n_i, H_i, W_i] with 0 / 1 / 255, and optionally the
    normalised xyxy boxes.  No host synchronisation happens before result()."""

    def __init__(self, device=None):
        self.ctx = native.Context.get(device)
        self.device = self.ctx.device
        self.state = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.box = torch.zeros(3, dtype=torch.float64, device=self.device)

    def update(self, pred_masks: Sequence[torch.Tensor], gt_masks: Sequence[torch.Tensor],
               pred_boxes: Optional[Sequence[torch.Tensor]] = None, gt_boxes: Optional[Sequence[torch.Tensor]] = None):
        assert len(pred_masks) == len(gt_masks)
        sizes = [int(p.shape[0]) for p in pred_masks]
        offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32).to(self.device,
                                                                                               non_blocking=True)
        counts = torch.empty((max(sum(sizes), 1), 6), dtype=torch.int32, device=self.device)
        lo = 0
        for p, g, n in zip(pred_masks, gt_masks, sizes):
            if n:
                assert p.shape == g.shape, (p.shape, g.shape)
                p = p.contiguous() if p.dtype == torch.float32 else p.float().contiguous()
                self.ctx.mask_iou_counts(p, _labels(g), 255, out=counts[lo:lo + n])
            lo += n
        self.ctx.seg_meter_update(counts, offsets, self.state)
        if pred_boxes is not None and gt_boxes is not None:
            for p, g in zip(pred_boxes, gt_boxes):
                if p.numel():
                    self.ctx.box_iou_diag(p, g.to(p.dtype), meter=self.box)

    def all_reduce(self):
        """One NCCL all-reduce of the 11 accumulators (the reference all-reduces each AverageMeter separately)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            buf = torch.cat([self.state, self.box])
            dist.all_reduce(buf, dist.ReduceOp.SUM)
            self.state, self.box = buf[:8].clone(), buf[8:].clone()

    def result(self):
        s = torch.cat([self.state, self.box]).cpu().numpy()   # the only device -> host copy
        iou_class = s[0:2] / (s[2:4] + 1e-10)
        ciou = iou_class[1] * 100.0
        giou = (s[4:6] / max(s[7], 1.0))[1] * 100.0
        prec05 = 100.0 * s[8] / max(s[9], 1.0)
        return {"ciou": float(ciou), "giou": float(giou), "prec05": float(prec05), "images": int(s[6]),
                "masks": int(s[7]), "boxes": int(s[9])}

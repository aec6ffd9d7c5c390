"""CPU restatements (numpy / torch-CPU) of the reference code on EITHER SIDE of the hot path: the evaluation
metrics, the sampling step of generate and the image preprocessing.

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path (u-llava_b200/ calls the CUDA library and
raises when it is missing).

Pinning: tests/test_oracle_golden.py checks these functions against
  * the reference's own evaluation/tools.py + evaluation/eval_ullava.py:validate, dataset/tools/mask_toolbox.py and
    dataset/processors/clip_processor.py run unmodified in the build container (fixtures tests/golden/callers_*.npz,
    written by tests/golden/make_golden_callers.py);
  * the third-party libraries the reference calls, which ARE installed where the tests run: Pillow's
    Image.resize (BICUBIC / BILINEAR), torchvision.ops.box_iou, transformers' CLIPImageProcessor.
The sampling restatement has no reference known-answer (torch.multinomial's random stream is not reproducible):
it is pinned on the distribution (HF TemperatureLogitsWarper + TopPLogitsWarper output) only.
"""
import math

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------------------------
# evaluation/tools.py:29-41  intersectionAndUnionGPU
# ------------------------------------------------------------------------------------------------------------------
def intersection_and_union(output: np.ndarray, target: np.ndarray, K: int = 2, ignore_index: int = 255):
    """Integer restatement: histc(x, bins=K, min=0, max=K-1) counts the values 0..K-1 and drops the rest."""
    output = output.reshape(-1).astype(np.int64).copy()
    target = target.reshape(-1).astype(np.int64)
    output[target == ignore_index] = ignore_index
    inter = output[output == target]
    hist = lambda x: np.array([(x == k).sum() for k in range(K)], dtype=np.int64)
    area_i, area_o, area_t = hist(inter), hist(output), hist(target)
    return area_i, area_o + area_t - area_i, area_t


# ------------------------------------------------------------------------------------------------------------------
# evaluation/eval_ullava.py:66-102  the accumulation inside validate() and its final ratios.
# counts_per_image: list over images of int arrays [n_i, 6] = I0 I1 U0 U1 T0 T1.
# The reference's histc runs on CUDA int tensors and returns int counts; python float 0.0 + int array promotes the
# running sums to float64, which is what is restated here.
# ------------------------------------------------------------------------------------------------------------------
def validate_meters(counts_per_image, box_hits_per_image=None):
    inter_sum = np.zeros(2, np.float64)
    union_sum = np.zeros(2, np.float64)
    acc_sum = np.zeros(2, np.float64)
    images = masks = 0
    for counts in counts_per_image:
        n = len(counts)
        if n == 0:
            continue
        intersection, union, acc_iou = 0.0, 0.0, 0.0
        for c in counts:
            i_i = np.asarray(c[0:2], np.int64)
            u_i = np.asarray(c[2:4], np.int64)
            intersection = intersection + i_i
            union = union + u_i
            acc_iou = acc_iou + i_i / (u_i + 1e-5)
            acc_iou[u_i == 0] += 1.0
        acc_iou = acc_iou / n
        inter_sum += intersection
        union_sum += union
        acc_sum += acc_iou * n
        images += 1
        masks += n
    out = {"state": np.concatenate([inter_sum, union_sum, acc_sum, [images, masks]]).astype(np.float64)}
    iou_class = inter_sum / (union_sum + 1e-10)
    out["ciou"] = iou_class[1] * 100.0
    out["giou"] = (acc_sum / max(masks, 1))[1] * 100.0
    if box_hits_per_image is not None:
        hits = sum(int(h) for img in box_hits_per_image for h in img)
        nbox = sum(len(img) for img in box_hits_per_image)
        out["prec05"] = 100.0 * hits / max(nbox, 1)
    return out


# ------------------------------------------------------------------------------------------------------------------
# evaluation/tools.py:13-26  bbox_iou = diag(torchvision.ops.box_iou(pred * 1000, gt * 1000)); torchvision 0.26:
# areas and the quotient in fp32 (_upcast), `* 1000` and `rb - lt` in the tensors' own dtype.
# ------------------------------------------------------------------------------------------------------------------
def box_iou_diag(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    a, b = pred * 1000, gt * 1000
    up = lambda t: t if t.dtype in (torch.float32, torch.float64) else t.float()
    a32, b32 = up(a), up(b)
    area1 = (a32[:, 2] - a32[:, 0]) * (a32[:, 3] - a32[:, 1])
    area2 = (b32[:, 2] - b32[:, 0]) * (b32[:, 3] - b32[:, 1])
    lt = torch.max(a[:, :2], b[:, :2])
    rb = torch.min(a[:, 2:], b[:, 2:])
    wh = up(rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (area1 + area2 - inter)


# ------------------------------------------------------------------------------------------------------------------
# Sampling step (GenerationMixin.generate with do_sample=True as called by models/ullava.py:350-362):
# TemperatureLogitsWarper, TopKLogitsWarper (GenerationConfig default top_k = 50 in the pinned transformers 4.29.1; the
# reference never overrides it), TopPLogitsWarper (transformers generation/logits_process.py), then one categorical draw.
# ------------------------------------------------------------------------------------------------------------------
def filtered_distribution(logits: np.ndarray, temperature: float, top_p=None, top_k=None) -> np.ndarray:
    """probabilities [V] (fp64) of the warped distribution; dropped tokens get 0."""
    s = logits.astype(np.float64) / temperature
    if top_k is not None and 0 < top_k < s.shape[0]:
        kth = np.sort(s)[-top_k]                       # scores < k-th largest are removed (ties stay)
        s = np.where(s < kth, -np.inf, s)
    e = np.exp(s - s.max())
    p = e / e.sum()
    if top_p is not None and 0.0 <= top_p < 1.0:
        order = np.argsort(p, kind="stable")           # ascending, like torch.sort(descending=False)
        cum = np.cumsum(p[order])
        remove = cum <= (1.0 - top_p)
        remove[-1:] = False                            # min_tokens_to_keep = 1
        p = p.copy()
        p[order[remove]] = 0.0
        p = p / p.sum()
    return p


def sample_inverse_cdf(logits: np.ndarray, temperature: float, top_p, u: float, top_k=None) -> int:
    """smallest index i with cdf(i) > u over the filtered distribution taken in vocabulary order."""
    p = filtered_distribution(logits, temperature, top_p, top_k)
    cdf = np.cumsum(p)
    idx = int(np.searchsorted(cdf, u * cdf[-1], side="right"))
    kept = np.nonzero(p > 0)[0]
    return int(min(idx, kept[-1]))


# ------------------------------------------------------------------------------------------------------------------
# Pillow 8-bit resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
# ImagingResampleHorizontal_8bpc / Vertical_8bpc), the code under PIL.Image.resize that CLIPImageProcessor.resize
# (BICUBIC) and ResizeLongestSide.apply_image (BILINEAR, models/segment_anything/utils/transforms.py:29-37) call.
# ------------------------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int, bicubic: bool):
    filt, fsupport = (_bicubic, 2.0) if bicubic else (_bilinear, 1.0)
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int, bicubic: bool) -> np.ndarray:
    """resample along axis 0 of an int64 array [n, ...] holding uint8 values."""
    bounds, kk = pil_coeffs(img.shape[0], out_size, bicubic)
    out = np.empty((out_size,) + img.shape[1:], np.int64)
    for yy in range(out_size):
        lo, n = bounds[yy]
        acc = np.tensordot(kk[yy, :n], img[lo:lo + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def pil_resize(img: np.ndarray, out_h: int, out_w: int, bicubic: bool) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C]; horizontal pass first, each pass skipped when the size is kept."""
    x = img.astype(np.int64)
    if out_w != img.shape[1]:
        x = _resample_axis0(x.transpose(1, 0, 2), out_w, bicubic).transpose(1, 0, 2)
    if out_h != img.shape[0]:
        x = _resample_axis0(x, out_h, bicubic)
    return x.astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# dataset/processors/clip_processor.py:55-94 + CLIPImageProcessor.preprocess (transformers 4.29.1 arithmetic:
# resize shortest edge (BICUBIC) -> center crop -> image * (1/255) in float64 rounded to float32 -> (x - mean) / std
# in float32), channels first.
# ------------------------------------------------------------------------------------------------------------------
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def pad_square_white(img: np.ndarray) -> np.ndarray:
    h, w = img.shape[:2]
    if h == w:
        return img
    size = max(h, w)
    out = np.full((size, size, 3), 255, dtype=img.dtype)
    if w > h:
        y0 = (w - h) // 2
        out[y0:y0 + h, :w] = img
    else:
        x0 = (h - w) // 2
        out[:h, x0:x0 + w] = img
    return out


def clip_resize_shape(h: int, w: int, size: int):
    """transformers get_resize_output_image_size(default_to_square=False): short side -> size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def clip_preprocess(img: np.ndarray, size: int = 336, pad: bool = False) -> np.ndarray:
    if pad:
        img = pad_square_white(img)
    oh, ow = clip_resize_shape(img.shape[0], img.shape[1], size)
    r = pil_resize(img, oh, ow, bicubic=True)
    top, left = (oh - size) // 2, (ow - size) // 2
    r = r[top:top + size, left:left + size]
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(CLIP_MEAN, np.float32)) / np.asarray(CLIP_STD, np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


# ------------------------------------------------------------------------------------------------------------------
# dataset/tools/mask_toolbox.py:8-28 + ResizeLongestSide (models/segment_anything/utils/transforms.py:29-37,61-70)
# ------------------------------------------------------------------------------------------------------------------
SAM_MEAN = (123.675, 116.28, 103.53)
SAM_STD = (58.395, 57.12, 57.375)


def sam_resize_shape(h: int, w: int, long_side: int = 1024):
    scale = long_side * 1.0 / max(h, w)
    return int(h * scale + 0.5), int(w * scale + 0.5)


def sam_preprocess(img: np.ndarray, sam_size: int = 1024) -> np.ndarray:
    """uint8 HWC (already resized by sam_apply_image) -> float32 [3, sam_size, sam_size]."""
    x = torch.from_numpy(img).permute(2, 0, 1).contiguous()
    x = (x - torch.tensor(SAM_MEAN).view(-1, 1, 1)) / torch.tensor(SAM_STD).view(-1, 1, 1)
    h, w = x.shape[-2:]
    x = torch.nn.functional.pad(x, (0, sam_size - w, 0, sam_size - h))
    return x.numpy()


def sam_apply_image(img: np.ndarray, long_side: int = 1024) -> np.ndarray:
    oh, ow = sam_resize_shape(img.shape[0], img.shape[1], long_side)
    return pil_resize(img, oh, ow, bicubic=False)

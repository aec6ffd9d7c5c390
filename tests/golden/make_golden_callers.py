"""Golden fixtures for the callers on either side of the hot path, produced by the
REAL reference code in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden_callers.py

  callers_metrics.npz / .json
      evaluation/eval_ullava.py:validate run UNMODIFIED on a stub model + synthetic dataset (random mask logits,
      targets with 0 / 1 / 255, random boxes), plus evaluation/tools.py:intersectionAndUnionGPU per mask.
  callers_preprocess.npz / .json
      dataset/tools/mask_toolbox.py:SegToolBox.apply_image + preprocess and
      dataset/processors/clip_processor.py:CLIPProcessor.pad_cv2 on random uint8 images (sha256 of the full fp32
      output + a strided sample; the arithmetic is exact, so the oracle must hit the hash).

Shims (stated with every result): torch.Tensor.cuda -> identity; stub modules for peft / pycocotools / tasks /
train_ullava / utils.config_builder / omegaconf / decord / imageio (imports of the reference scripts that are not on this path);
DataLoader(num_workers=0); torch.histc on integer CPU tensors -> histc of the float copy cast back to the input
dtype (what the CUDA kernel the reference runs on returns; the CPU kernel does not implement integer inputs).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
os.chdir(REF)  # the reference scripts do sys.path.append(os.getcwd())

torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.empty_cache = lambda: None
_histc = torch.histc


def histc_like_cuda(t, bins=100, min=0, max=0):
    if t.is_floating_point():
        return _histc(t, bins=bins, min=min, max=max)
    return _histc(t.float(), bins=bins, min=min, max=max).to(t.dtype)


torch.histc = histc_like_cuda

for name in ("peft", "pycocotools", "pycocotools.mask", "tasks", "train_ullava", "utils.config_builder", "deepspeed", "omegaconf",
             "decord", "imageio", "imageio.v3"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            sys.modules[name] = m
sys.modules["peft"].PeftModel = object
if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
    sys.modules["omegaconf"].OmegaConf = object
sys.modules["decord"].VideoReader = object
sys.modules["decord"].bridge = types.SimpleNamespace(set_bridge=lambda *a, **k: None)
sys.modules["tasks"].setup_task = lambda *a, **k: None
sys.modules["utils.config_builder"].Config = object
sys.modules["train_ullava"].ModelArguments = object
sys.modules["train_ullava"].TrainingArguments = object
if hasattr(sys.modules["pycocotools"], "__path__") is False:
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]

import torch.utils.data as tud  # noqa: E402

_DL = tud.DataLoader
tud.DataLoader = lambda ds, **kw: _DL(ds, **{**kw, "num_workers": 0})

import evaluation.eval_ullava as ref_eval  # noqa: E402
from evaluation.tools import intersectionAndUnionGPU, bbox_iou  # noqa: E402


def save(name, arrays, meta):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


def metrics():
    g = torch.Generator().manual_seed(20240)
    items = []
    shapes = [(3, 24, 40), (1, 33, 17), (2, 16, 16), (4, 20, 30), (1, 8, 8)]
    for i, (n, h, w) in enumerate(shapes):
        logits = torch.randn((n, h, w), generator=g)
        gt = (torch.rand((n, h, w), generator=g) > 0.6).float()
        ignore = torch.rand((n, h, w), generator=g) > 0.9
        gt[ignore] = 255.0
        if i == 2:           # an empty target AND an empty prediction: union == 0 -> acc_iou += 1 branch
            logits[0] = -1.0
            gt[0] = 0.0
        box = torch.rand((n, 2), generator=g) * 0.5
        pred_boxes = torch.cat([box, box + 0.1 + torch.rand((n, 2), generator=g) * 0.4], 1)
        gt_boxes = (pred_boxes + torch.randn((n, 4), generator=g) * 0.06).clamp(0, 1)
        items.append({"pred_masks": logits, "gt_masks": gt, "pred_boxes": pred_boxes, "gt_boxes": gt_boxes})

    class Stub:
        def eval(self):
            return self

        def __call__(self, idx=None, inference=None, **kw):
            it = items[int(idx)]
            return {"pred_masks": [it["pred_masks"]], "gt_masks": [it["gt_masks"]],
                    "pred_boxes": [it["pred_boxes"]], "gt_boxes": [it["gt_boxes"]]}

    dataset = [{"idx": i} for i in range(len(items))]
    ciou, giou, prec05 = ref_eval.validate(Stub(), dataset, lambda batch: dict(batch[0]), torch.float32)

    arrays, counts = {}, []
    for i, it in enumerate(items):
        arrays[f"logits_{i}"] = it["pred_masks"].numpy()
        arrays[f"gt_{i}"] = it["gt_masks"].numpy().astype(np.uint8)
        arrays[f"pred_boxes_{i}"] = it["pred_boxes"].numpy()
        arrays[f"gt_boxes_{i}"] = it["gt_boxes"].numpy()
        out = (it["pred_masks"] > 0).int()
        tgt = it["gt_masks"].int()
        per = []
        for o, t in zip(out, tgt):
            a_i, a_u, a_t = intersectionAndUnionGPU(o.contiguous().clone(), t.contiguous(), 2, ignore_index=255)
            per.append(np.concatenate([a_i.numpy(), a_u.numpy(), a_t.numpy()]))
        counts.append(np.stack(per))
        arrays[f"counts_{i}"] = counts[-1]
        arrays[f"box_iou_{i}"] = np.array(
            [bbox_iou(p.unsqueeze(0), q.unsqueeze(0))["miou"] for p, q in zip(it["pred_boxes"], it["gt_boxes"])],
            dtype=np.float64)
    save("callers_metrics", arrays,
         {"n_images": len(items), "ciou": float(ciou), "giou": float(giou), "prec05": float(prec05),
          "source": "evaluation/eval_ullava.py:validate + evaluation/tools.py, unmodified, CPU, shims in the header"})


def preprocess():
    from dataset.tools.mask_toolbox import SegToolBox
    from dataset.processors.clip_processor import CLIPProcessor
    rng = np.random.default_rng(7)
    tool = SegToolBox()
    arrays, meta = {}, {"cases": []}
    for i, (h, w) in enumerate([(48, 64), (75, 50), (40, 40)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        resized = tool.apply_image(img)
        x = tool.preprocess(torch.from_numpy(resized).permute(2, 0, 1).contiguous()).numpy()
        padded = CLIPProcessor.pad_cv2(img)
        arrays[f"img_{i}"] = img
        arrays[f"sam_sample_{i}"] = x.reshape(-1)[::997].copy()
        arrays[f"pad_{i}"] = padded
        meta["cases"].append({"h": h, "w": w, "resized": list(resized.shape[:2]),
                              "resized_sha256": hashlib.sha256(np.ascontiguousarray(resized).tobytes()).hexdigest(),
                              "sam_sha256": hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(),
                              "sam_shape": list(x.shape)})
    meta["source"] = "dataset/tools/mask_toolbox.py SegToolBox + dataset/processors/clip_processor.py pad_cv2, unmodified"
    save("callers_preprocess", arrays, meta)


def tiny_core_video():
    """UllavaCoreForCausalLM.forward on a batch mixing a video row and an image row (reference
    models/ullava_core.py:160-180,248-269), tiny config, seeded synthetic weights: logits + encode_video output."""
    import models as ref_models
    from oracle.synth import synth_normal, synth_state_dict, synth_ids
    from tests import configs as C
    with open(os.path.join(HERE, "tiny_core.json")) as f:
        meta = json.load(f)
    cfg = ref_models.UllavaCoreConfig(**C.TINY_LLM)
    cfg._attn_implementation = "eager"
    cfg.vision_config._attn_implementation = "eager"
    m = ref_models.UllavaCoreForCausalLM(cfg).eval().float()
    m.load_state_dict(synth_state_dict(meta["shapes"], meta["seed"]), strict=True)
    T, N = 3, 4                      # frames, patches per 28 x 28 frame
    videos = synth_normal("videos", (1, 3, T, 28, 28))
    images = synth_normal("images", (1, 3, 28, 28))
    ids_img = C.tiny_prompt(1)
    L = ids_img.shape[1]
    vid = [1] + synth_ids("vhead", (1, 2), 3, 300)[0].tolist() + [C.MM_IDS["VID_START"]] + \
          [C.MM_IDS["VID_PATCH"]] * (T + N) + [C.MM_IDS["VID_END"]]
    vid = vid + synth_ids("vtail", (1, L - len(vid)), 3, 300)[0].tolist()
    ids = torch.cat([torch.tensor([vid], dtype=torch.int64), ids_img], 0)
    with torch.no_grad():
        out = m(input_ids=ids, images=images, videos=videos, return_dict=True)
        vf = m.encode_video(videos)
    save("tiny_core_video", {"ids": ids.numpy(), "logits": out.logits.numpy(), "video_features": vf.numpy()},
         {"frames": T, "patches": N, "seed": meta["seed"],
          "source": "models/ullava_core.py forward + encode_video, unmodified, transformers eager, fp32 CPU"})


if __name__ == "__main__":
    metrics()
    preprocess()
    tiny_core_video()

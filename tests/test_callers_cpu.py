"""CPU pinning of oracle/callers_oracle.py: against the fixtures written by the
REAL reference (tests/golden/make_golden_callers.py: evaluation/eval_ullava.py:validate, evaluation/tools.py,
SegToolBox, CLIPProcessor.pad_cv2) and against the third-party code the reference calls, where it is installed
(Pillow, torchvision.ops.box_iou, transformers' PIL-backed CLIP processor and logits warpers)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import callers_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        meta = json.load(f)
    return np.load(os.path.join(GOLD, name + ".npz")), meta


def test_intersection_union_counts_match_reference():
    z, meta = _load("callers_metrics")
    for i in range(meta["n_images"]):
        logits, gt = z[f"logits_{i}"], z[f"gt_{i}"]
        for m in range(logits.shape[0]):
            a_i, a_u, a_t = O.intersection_and_union((logits[m] > 0).astype(np.int32), gt[m], 2, 255)
            assert np.array_equal(np.concatenate([a_i, a_u, a_t]), z[f"counts_{i}"][m])


def test_validate_meters_match_reference_validate():
    """ciou / giou / prec@0.5 returned by the reference's validate() on the same synthetic dataset: bit-exact."""
    z, meta = _load("callers_metrics")
    counts = [z[f"counts_{i}"] for i in range(meta["n_images"])]
    hits = []
    for i in range(meta["n_images"]):
        iou = O.box_iou_diag(torch.from_numpy(z[f"pred_boxes_{i}"]), torch.from_numpy(z[f"gt_boxes_{i}"]))
        assert np.array_equal(iou.double().numpy(), z[f"box_iou_{i}"])
        hits.append((iou > 0.5).tolist())
    r = O.validate_meters(counts, hits)
    assert r["ciou"] == meta["ciou"] and r["giou"] == meta["giou"] and r["prec05"] == meta["prec05"], (r, meta)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_box_iou_matches_torchvision(dtype):
    ops = pytest.importorskip("torchvision.ops")
    g = torch.Generator().manual_seed(5)
    p = torch.rand((200, 4), generator=g)
    p[:, 2:] = p[:, :2] + torch.rand((200, 2), generator=g) * 0.5
    q = p + torch.randn((200, 4), generator=g) * 0.05
    p, q = p.to(dtype), q.to(dtype)
    ref = torch.diagonal(ops.box_iou(p * 1000, q * 1000))
    got = O.box_iou_diag(p, q)
    assert torch.equal(ref.float().nan_to_num(-1), got.float().nan_to_num(-1))


@pytest.mark.parametrize("h,w,oh,ow", [(480, 640, 768, 1024), (500, 333, 1024, 682), (640, 480, 448, 336),
                                       (97, 53, 31, 200), (300, 300, 300, 150), (64, 48, 64, 48), (1, 7, 5, 3)])
@pytest.mark.parametrize("bicubic", [False, True])
def test_pil_resize_restatement_is_bit_exact(h, w, oh, ow, bicubic):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC if bicubic else Image.BILINEAR))
    assert np.array_equal(O.pil_resize(img, oh, ow, bicubic), ref)


@pytest.mark.parametrize("bicubic", [False, True])
def test_library_tap_tables_equal_the_pillow_restatement(bicubic):
    """The host half of ullava_resize_u8 (C++ restatement of precompute_coeffs / normalize_coeffs_8bpc inside the
    library, no GPU involved) gives the oracle's tap tables bit for bit, for up- and down-scaling, odd sizes, 1 pixel."""
    import native
    for in_size, out_size in [(640, 1024), (480, 768), (333, 682), (1920, 597), (53, 200), (97, 31), (1, 5), (7, 3),
                              (300, 300), (336, 336), (1024, 64), (2, 1)]:
        bounds, taps = native.resample_coeffs(in_size, out_size, bicubic)
        ob, ok = O.pil_coeffs(in_size, out_size, bicubic)
        assert taps.shape == ok.shape, (in_size, out_size)
        assert np.array_equal(bounds, ob) and np.array_equal(taps, ok), (in_size, out_size)


def test_pil_resize_restatement_random_geometries():
    """Property test: for random small geometries (up / down / mixed scaling, degenerate 1-pixel sides) the restatement
    and the library's tap tables agree with Pillow bit for bit."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    Image = pytest.importorskip("PIL.Image")
    import native

    @hyp.settings(max_examples=60, deadline=None, derandomize=True)
    @hyp.given(st.integers(1, 48), st.integers(1, 48), st.integers(1, 80), st.integers(1, 80), st.booleans(),
               st.integers(0, 2 ** 31 - 1))
    def check(h, w, oh, ow, bicubic, seed):
        img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC if bicubic else Image.BILINEAR))
        assert np.array_equal(O.pil_resize(img, oh, ow, bicubic), ref)
        for a, b in ((w, ow), (h, oh)):
            lb, lk = native.resample_coeffs(a, b, bicubic)
            ob, ok = O.pil_coeffs(a, b, bicubic)
            assert np.array_equal(lb, ob) and np.array_equal(lk, ok)

    check()


def test_sam_preprocess_matches_reference_segtoolbox():
    z, meta = _load("callers_preprocess")
    for i, case in enumerate(meta["cases"]):
        img = z[f"img_{i}"]
        resized = O.sam_apply_image(img)
        assert list(resized.shape[:2]) == case["resized"]
        assert hashlib.sha256(np.ascontiguousarray(resized).tobytes()).hexdigest() == case["resized_sha256"]
        x = O.sam_preprocess(resized)
        assert list(x.shape) == case["sam_shape"]
        assert np.array_equal(x.reshape(-1)[::997], z[f"sam_sample_{i}"])
        assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == case["sam_sha256"]
        assert np.array_equal(O.pad_square_white(img), z[f"pad_{i}"])


@pytest.mark.parametrize("h,w", [(480, 640), (640, 480), (336, 336), (500, 333), (97, 53)])
def test_clip_preprocess_matches_hf_pil_processor(h, w):
    tr = pytest.importorskip("transformers")
    if not hasattr(tr, "CLIPImageProcessorPil"):
        pytest.skip("no PIL-backed CLIP processor in this transformers")
    proc = tr.CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = proc.preprocess(img, return_tensors="pt")["pixel_values"][0].numpy()
    assert np.array_equal(O.clip_preprocess(img, 336), ref)


@pytest.mark.parametrize("temperature,top_p,top_k", [(0.2, None, None), (0.2, 0.7, None), (1.0, 0.9, None),
                                                      (0.7, 0.3, None), (0.2, None, 50), (1.0, 0.9, 50), (1.3, 0.5, 7),
                                                      (1.0, None, 1), (1.0, 0.0, None), (1.0, 0.0, 50)])
def test_filtered_distribution_matches_hf_warpers(temperature, top_p, top_k):
    """The warper chain of HF generate(do_sample=True): temperature, top-k (GenerationConfig default 50), top-p."""
    lp = pytest.importorskip("transformers.generation.logits_process")
    g = torch.Generator().manual_seed(11)
    logits = torch.randn((4, 997), generator=g) * 3
    scores = lp.TemperatureLogitsWarper(temperature)(None, logits.clone())
    if top_k is not None:
        scores = lp.TopKLogitsWarper(top_k)(None, scores)
    if top_p is not None:
        scores = lp.TopPLogitsWarper(top_p)(None, scores)
    ref = scores.double().softmax(-1).numpy()
    for b in range(4):
        p = O.filtered_distribution(logits[b].numpy(), temperature, top_p, top_k)
        assert np.array_equal(p > 0, ref[b] > 0)
        assert np.abs(p - ref[b]).max() < 1e-6
        # inverse CDF: the drawn index is a kept token and its CDF interval contains u
        for u in (0.0, 0.25, 0.5, 0.999999):
            i = O.sample_inverse_cdf(logits[b].numpy(), temperature, top_p, u, top_k)
            cdf = np.cumsum(p)
            assert p[i] > 0 and cdf[i] > u * cdf[-1] - 1e-12 and (cdf[i] - p[i]) <= u * cdf[-1] + 1e-12


def test_video_branch_of_the_oracle_matches_reference():
    """encode_video + the video branch of embed_images_videos (models/ullava_core.py:160-180,248-269): logits of a
    batch mixing a video row and an image row, as the real reference computed them."""
    from oracle import ullava_oracle as U
    from oracle.synth import synth_normal, synth_state_dict
    from tests import configs as C
    z, meta = _load("tiny_core_video")
    _, core = _load("tiny_core")
    sd = synth_state_dict(core["shapes"], meta["seed"])
    cfg = dict(C.TINY_LLM)
    videos = synth_normal("videos", (1, 3, meta["frames"], 28, 28))
    images = synth_normal("images", (1, 3, 28, 28))
    vf = U.encode_video(sd, videos, cfg)
    assert np.abs(vf.numpy() - z["video_features"]).max() < 2e-4
    out = U.core_forward(sd, cfg, torch.from_numpy(z["ids"]), images, videos=videos)
    assert np.abs(out["logits"].numpy() - z["logits"]).max() < 2e-4


def test_host_mirrors_import_without_gpu_and_fail_loudly():
    """The evaluation / dataset mirrors import on a CPU box, and every entry point raises (no fallback) without CUDA."""
    import native
    from evaluation.tools import intersectionAndUnionGPU, bbox_iou, SegMeter, AverageMeter, Summary, dict_to_cuda  # noqa: F401
    from evaluation.eval_ullava import validate  # noqa: F401
    from dataset.processors.clip_processor import CLIPProcessor
    from dataset.tools.mask_toolbox import SegToolBox
    assert CLIPProcessor.resize_shape(480, 640, 336) == O.clip_resize_shape(480, 640, 336) == (336, 448)
    assert SegToolBox.get_preprocess_shape(480, 640, 1024) == O.sam_resize_shape(480, 640) == (768, 1024)
    assert SegToolBox.get_preprocess_shape(500, 333, 1024) == O.sam_resize_shape(500, 333)
    m = AverageMeter("x", ":6.3f", Summary.SUM)
    m.update(np.array([1.0, 2.0]), n=2)
    assert np.array_equal(m.sum, np.array([2.0, 4.0])) and m.count == 2
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises((native.NativeError, RuntimeError, AssertionError)):
        intersectionAndUnionGPU(torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32), 2)
    with pytest.raises((native.NativeError, RuntimeError, AssertionError)):
        SegMeter(0)

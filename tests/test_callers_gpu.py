"""GPU parity of the kernels on either side of the hot path, through the C ABI and
the host mirrors of the reference modules, against oracle/callers_oracle.py and the fixtures the REAL reference wrote
(tests/golden/callers_*.npz).  Integer / byte work and the exactly-rounded fp32 elementwise chains are compared
bit for bit; the sampling kernel is compared on the filtered distribution (1e-5) and on its inverse-CDF draw."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import callers_oracle as O
from tests.util_models import build_tiny_core, load_golden, oracle_inputs_core

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


# ---------------------------------------------------------------------------------------------------------------
# f2: evaluation metrics
# ---------------------------------------------------------------------------------------------------------------
def test_mask_iou_counts_match_reference_fixture(ctx):
    z, meta = load_golden("callers_metrics")
    for i in range(meta["n_images"]):
        logits = torch.from_numpy(z[f"logits_{i}"]).cuda()
        gt = torch.from_numpy(z[f"gt_{i}"]).cuda()
        ref = torch.from_numpy(z[f"counts_{i}"].astype(np.int32))
        assert torch.equal(ctx.mask_iou_counts(logits, gt).cpu(), ref)                       # fp32 logits, uint8 gt
        labels = (logits > 0).to(torch.int32)
        assert torch.equal(ctx.mask_iou_counts(labels, gt.to(torch.int32)).cpu(), ref)       # int32 / int32
        assert torch.equal(ctx.mask_iou_counts(labels.to(torch.uint8), gt).cpu(), ref)       # uint8 / uint8


@pytest.mark.parametrize("n,h,w", [(5, 336, 336), (1, 480, 640), (3, 37, 41), (10, 1, 1)])
def test_mask_iou_counts_vs_oracle(ctx, n, h, w):
    g = torch.Generator().manual_seed(n * 100 + h)
    logits = torch.randn((n, h, w), generator=g)
    gt = (torch.rand((n, h, w), generator=g) > 0.5).to(torch.uint8)
    gt[torch.rand((n, h, w), generator=g) > 0.93] = 255
    gt[0][torch.rand((h, w), generator=g) > 0.97] = 7          # a label outside {0, 1, ignore}: histc drops it
    got = ctx.mask_iou_counts(logits.cuda(), gt.cuda()).cpu().numpy()
    for m in range(n):
        a_i, a_u, a_t = O.intersection_and_union((logits[m] > 0).numpy().astype(np.int32), gt[m].numpy(), 2, 255)
        assert np.array_equal(got[m], np.concatenate([a_i, a_u, a_t]))


def test_seg_meter_reproduces_reference_validate(ctx):
    """The device meters over the reference's synthetic dataset give validate()'s ciou / giou / prec@0.5 exactly,
    whatever the batching of the updates."""
    from evaluation.tools import SegMeter, intersectionAndUnionGPU, bbox_iou
    z, meta = load_golden("callers_metrics")
    n = meta["n_images"]
    items = [dict(p=torch.from_numpy(z[f"logits_{i}"]).cuda(), g=torch.from_numpy(z[f"gt_{i}"]).cuda(),
                  pb=torch.from_numpy(z[f"pred_boxes_{i}"]).cuda(), gb=torch.from_numpy(z[f"gt_boxes_{i}"]).cuda())
             for i in range(n)]
    for chunks in ([[0], [1], [2], [3], [4]], [[0, 1, 2, 3, 4]], [[0, 1], [2, 3, 4]]):
        meter = SegMeter(0)
        for ch in chunks:
            meter.update([items[i]["p"] for i in ch], [items[i]["g"] for i in ch], [items[i]["pb"] for i in ch],
                         [items[i]["gb"] for i in ch])
        r = meter.result()
        assert r["ciou"] == meta["ciou"] and r["giou"] == meta["giou"] and r["prec05"] == meta["prec05"], (r, meta)
        assert r["images"] == n
    # the drop-in functions of evaluation/tools.py
    a_i, a_u, a_t = intersectionAndUnionGPU((items[0]["p"][0] > 0).int(), items[0]["g"][0].int(), 2, ignore_index=255)
    assert torch.equal(torch.cat([a_i, a_u, a_t]).cpu(), torch.from_numpy(z["counts_0"][0].astype(np.int32)))
    r = bbox_iou(items[0]["pb"][:1], items[0]["gb"][:1])
    assert r["miou"] == z["box_iou_0"][0] and r["num"] == 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_box_iou_diag_bit_exact(ctx, dtype):
    g = torch.Generator().manual_seed(5)
    p = torch.rand((333, 4), generator=g)
    p[:, 2:] = p[:, :2] + torch.rand((333, 2), generator=g) * 0.5
    q = p + torch.randn((333, 4), generator=g) * 0.05
    q[7] = p[7] + 2.0                                   # disjoint boxes: iou 0
    p, q = p.to(dtype), q.to(dtype)
    ref = O.box_iou_diag(p, q).float()
    meter = torch.zeros(3, dtype=torch.float64, device="cuda")
    got = ctx.box_iou_diag(p.cuda(), q.cuda(), meter=meter).cpu()
    assert torch.equal(got.nan_to_num(-1), ref.nan_to_num(-1))
    assert meter[:2].tolist() == [float((ref > 0.5).sum()), 333.0]


def test_validate_batched_right_padded_equals_per_image(ctx):
    """evaluation.eval_ullava.validate on the tiny model: a right-padded batch of 2 (different prompt lengths, the
    reference collator's pad_sequence layout) gives the same ciou / giou / prec@0.5 as the reference's batch-1 loop
    shape, and the meters match the oracle's restatement of validate() fed with the model's own masks."""
    from evaluation.eval_ullava import validate
    from tests import configs as C
    from tests.util_models import build_tiny_full, oracle_inputs_full
    dtype = torch.bfloat16
    model, sd, cfg = build_tiny_full(dtype)
    ids, images, images_sam, sizes, resizes = oracle_inputs_full()
    g = torch.Generator().manual_seed(9)
    items = []
    for b in range(2):
        row = ids[b] if b == 0 else ids[b][:-1]                       # second prompt one token shorter -> padded
        n_seg, n_loc = int((row == C.SEG_ID).sum()), int((row == C.LOC_ID).sum())
        gt = (torch.rand((n_seg,) + tuple(sizes[b]), generator=g) > 0.5).float()
        gt[torch.rand(gt.shape, generator=g) > 0.9] = 255.0
        box = torch.rand((n_loc, 2), generator=g) * 0.5
        items.append(dict(input_ids=row, labels=row.clone(), image=images[b], image_sam=images_sam[b], seg_mask=gt,
                          raw_size=sizes[b], resize=resizes[b], boxes=torch.cat([box, box + 0.3], 1)))

    def collate(instances):   # dataset/collators/base_collator.py:107-123 (GroundingCollator)
        pad = lambda key, v: torch.nn.utils.rnn.pad_sequence([i[key] for i in instances], batch_first=True, padding_value=v)
        input_ids = pad("input_ids", 0)
        return dict(input_ids=input_ids, labels=pad("labels", -100), attention_mask=input_ids.ne(0),
                    images=torch.stack([i["image"] for i in instances]),
                    images_sam=torch.stack([i["image_sam"] for i in instances]),
                    mask_list=[i["seg_mask"] for i in instances], size_list=[i["raw_size"] for i in instances],
                    resize_list=[i["resize"] for i in instances], bbox_list=[i["boxes"] for i in instances])

    one = validate(model, items, collate, dtype, batch_size=1, num_workers=0, verbose=False)
    two = validate(model, items, collate, dtype, batch_size=2, num_workers=0, verbose=False)
    assert all(np.isfinite(v) for v in one) and one[0] > 0
    # the padded batch takes the same kernels row for row; a pixel exactly at the threshold may still flip
    assert abs(one[0] - two[0]) < 0.5 and abs(one[1] - two[1]) < 0.5 and one[2] == two[2], (one, two)
    # oracle restatement of validate() on the masks / boxes the model produced
    counts, hits = [], []
    for it in items:
        from evaluation.tools import dict_to_cuda
        out = model(**dict_to_cuda(collate([it]), dtype), inference=True)
        pm = out["pred_masks"][0].float().cpu().numpy()
        gt = it["seg_mask"].numpy()
        counts.append(np.stack([np.concatenate(O.intersection_and_union((pm[m] > 0).astype(np.int32), gt[m], 2, 255))
                                for m in range(pm.shape[0])]))
        iou = O.box_iou_diag(out["pred_boxes"][0].cpu(), it["boxes"].to(dtype))
        hits.append((iou > 0.5).tolist())
    ref = O.validate_meters(counts, hits)
    assert one[0] == ref["ciou"] and one[1] == ref["giou"] and one[2] == ref["prec05"], (one, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_cross_entropy_kernel_matches_torch(ctx, dtype):
    """Shifted token cross-entropy (models/ullava_core.py:327-338) against torch.nn.functional.cross_entropy on the
    same logits: ignore_index rows, a strided logits view, and the all-ignored NaN case."""
    g = torch.Generator().manual_seed(12)
    B, T, V = 3, 37, 32011
    logits = (torch.randn((B, T, V), generator=g) * 3).to(dtype).cuda()
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[0, :9] = -100
    labels[2, 5::3] = -100
    labels = labels.cuda()
    ref = torch.nn.functional.cross_entropy(logits[:, :-1].float().reshape(-1, V), labels[:, 1:].reshape(-1))
    got = ctx.cross_entropy(logits, labels)
    assert abs(got.item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (got.item(), ref.item())
    view = logits[:, :20]                                              # batch stride != T * V
    ref_v = torch.nn.functional.cross_entropy(view[:, :-1].float().reshape(-1, V), labels[:, 1:20].reshape(-1))
    assert abs(ctx.cross_entropy(view, labels).item() - ref_v.item()) < 2e-5 * max(1.0, abs(ref_v.item()))
    assert torch.isnan(ctx.cross_entropy(logits, torch.full_like(labels, -100)))
    # through the module API: forward(labels=...) returns the loss of its own logits
    model, sd, cfg = build_tiny_core(torch.bfloat16)
    ids, images = oracle_inputs_core(2)
    out = model(input_ids=ids.cuda(), images=images.cuda().to(torch.bfloat16), labels=ids.cuda(), return_dict=True)
    ref_m = torch.nn.functional.cross_entropy(out.logits[:, :-1].float().reshape(-1, out.logits.shape[-1]),
                                              ids.cuda()[:, 1:].reshape(-1))
    assert abs(out.loss.item() - ref_m.item()) < 1e-4 * max(1.0, abs(ref_m.item()))


# ---------------------------------------------------------------------------------------------------------------
# f3: preprocessing
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,oh,ow", [(480, 640, 768, 1024), (500, 333, 1024, 682), (640, 480, 448, 336),
                                       (1080, 1920, 336, 597), (97, 53, 31, 200), (300, 300, 300, 150),
                                       (64, 48, 64, 48), (1, 7, 5, 3)])
@pytest.mark.parametrize("bicubic", [False, True])
def test_resize_u8_is_pillow_exact(ctx, h, w, oh, ow, bicubic):
    from PIL import Image
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC if bicubic else Image.BILINEAR))
    got = ctx.resize_u8(torch.from_numpy(img).cuda(), oh, ow, bicubic).cpu().numpy()
    assert np.array_equal(got, ref), f"max diff {np.abs(got.astype(int) - ref.astype(int)).max()}"
    if h * w < 400 * 400:
        assert np.array_equal(got, O.pil_resize(img, oh, ow, bicubic))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_seg_toolbox_matches_reference_fixture(ctx, dtype):
    from dataset.tools.mask_toolbox import SegToolBox
    z, meta = load_golden("callers_preprocess")
    tool = SegToolBox(dtype=dtype)
    assert SegToolBox().dtype == torch.float32        # the reference's processors return fp32
    for i, case in enumerate(meta["cases"]):
        img = z[f"img_{i}"]
        resized = tool.apply_image(img)
        assert hashlib.sha256(resized.cpu().numpy().tobytes()).hexdigest() == case["resized_sha256"]
        x, resize = tool(img)
        assert list(resize) == case["resized"] and list(x.shape) == case["sam_shape"]
        ref32 = O.sam_preprocess(O.sam_apply_image(img))
        assert hashlib.sha256(np.ascontiguousarray(ref32).tobytes()).hexdigest() == case["sam_sha256"]
        assert torch.equal(x.cpu(), torch.from_numpy(ref32).to(dtype))   # the reference casts the fp32 tensor (tools.py:55-67)
        # the reference's calling convention: CHW uint8 tensor in
        x2 = tool.preprocess(resized.permute(2, 0, 1))
        assert torch.equal(x2, x)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("h,w,pad", [(480, 640, False), (640, 480, True), (336, 336, False), (500, 333, True),
                                     (97, 53, False)])
def test_clip_processor_matches_oracle(ctx, dtype, h, w, pad):
    from dataset.processors.clip_processor import CLIPProcessor
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    proc = CLIPProcessor(aspect_ratio="pad" if pad else None, size=336, dtype=dtype)
    got = proc(img)
    ref = torch.from_numpy(O.clip_preprocess(img, 336, pad=pad)).to(dtype)
    assert got.shape == (3, 336, 336) and torch.equal(got.cpu(), ref)


# ---------------------------------------------------------------------------------------------------------------
# f4: video path + sampling
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bs,t,n,d", [(2, 8, 256, 1024), (1, 3, 4, 64), (3, 1, 7, 128)])
def test_video_pool_kernel(ctx, dtype, bs, t, n, d):
    g = torch.Generator().manual_seed(t * n)
    feats = torch.randn((bs, t, n, d), generator=g).to(dtype).cuda()
    got = ctx.video_pool(feats).float()
    f32 = feats.float()
    ref = torch.cat([f32.mean(dim=2), f32.mean(dim=1)], dim=1)
    assert got.shape == (bs, t + n, d)
    # fp32 accumulation + one rounding to the 16-bit type: half an ulp of the output (2^-9 bf16, 2^-12 fp16)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert bool(((got - ref).abs() <= ulp * ref.abs() + 1e-6).all())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_video_and_image_rows_match_reference_golden(ctx, dtype):
    """A batch mixing a <vid_beg> row and an <img_beg> row through UllavaCoreForCausalLM.forward(videos=...)."""
    from oracle.synth import synth_normal
    z, meta = load_golden("tiny_core_video")
    model, sd, cfg = build_tiny_core(dtype)
    videos = synth_normal("videos", (1, 3, meta["frames"], 28, 28)).cuda().to(dtype)
    images = synth_normal("images", (1, 3, 28, 28)).cuda().to(dtype)
    vf = model.encode_video(videos)
    assert np.abs(vf.float().cpu().numpy() - z["video_features"]).max() < (2e-2 if dtype == torch.float16 else 8e-2)
    out = model(input_ids=torch.from_numpy(z["ids"]).cuda(), images=images, videos=videos, return_dict=True)
    err = np.abs(out.logits.float().cpu().numpy() - z["logits"]).max()
    assert err < (1e-2 if dtype == torch.float16 else 8e-2), err


# ---------------------------------------------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V", [32011, 997, 64])
@pytest.mark.parametrize("temperature,top_p,top_k", [(0.2, None, 0), (0.2, 0.7, 0), (1.0, 0.9, 0), (0.7, 0.3, 0),
                                                      (1.5, 1.0, 0), (0.2, None, 50), (1.0, 0.9, 50), (1.3, 0.5, 7),
                                                      (1.0, None, 1), (1.0, 0.0, 0), (1.0, 0.0, 50)])
def test_sample_step_distribution_and_draw(ctx, V, temperature, top_p, top_k):
    B = 6
    g = torch.Generator().manual_seed(V + int(temperature * 10))
    logits = torch.randn((B, V), generator=g) * 3
    u = torch.rand((1, B), generator=g)
    u[0, 0], u[0, 1] = 0.0, 0.9999999
    probs = torch.empty((B, V), dtype=torch.float32, device="cuda")
    ids = torch.empty((B,), dtype=torch.int64, device="cuda")
    ctx.sample_step(logits.cuda(), temperature, top_p, u.cuda(), ids, probs_out=probs, top_k=top_k)
    probs, ids = probs.cpu().double().numpy(), ids.cpu().numpy()
    for b in range(B):
        ref = O.filtered_distribution(logits[b].numpy(), temperature, top_p, top_k)
        if top_k:
            assert int((probs[b] > 0).sum()) <= top_k
        # kept set: identical except for tokens whose ascending cumulative mass sits within rounding of the cut
        diff = np.nonzero((probs[b] > 0) != (ref > 0))[0]
        diff = np.array([t for t in diff if ref[t] > 1e-30], dtype=np.int64)   # fp32 exp underflows to 0, fp64 does not
        if len(diff):
            full = O.filtered_distribution(logits[b].numpy(), temperature, None, top_k)
            order = np.argsort(full, kind="stable")
            cum = np.cumsum(full[order])
            pos = {int(t): k for k, t in enumerate(order)}
            assert all(abs(cum[pos[int(t)]] - (1 - top_p)) < 1e-5 for t in diff), diff
        else:
            assert np.abs(probs[b] - ref).max() < 1e-5
        assert abs(probs[b].sum() - 1) < 1e-4
        # the draw: inverse CDF of the kernel's own distribution, in vocabulary order
        cdf = np.cumsum(probs[b])
        i, t = int(ids[b]), float(u[0, b]) * cdf[-1]
        assert probs[b][i] > 0 and cdf[i] - probs[b][i] <= t + 1e-5 and t <= cdf[i] + 1e-5, (i, t, cdf[i], probs[b][i])
        if not len(diff):
            j = O.sample_inverse_cdf(logits[b].numpy(), temperature, top_p, float(u[0, b]), top_k)
            if j != i:   # only when u * Z falls within fp32 rounding of a CDF step
                assert abs(i - j) <= max(2, V // 1000) or min(abs(t - cdf[i]), abs(t - (cdf[i] - probs[b][i]))) < 1e-5


def test_sample_step_statistics_and_bookkeeping(ctx):
    """20 000 draws of one small distribution follow it (5 sigma per bin); eos / pad / sequence bookkeeping as in
    ullava_greedy_step."""
    V, N = 50, 20000
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn((1, V), generator=g) * 2).expand(N, V).contiguous()
    u = torch.rand((1, N), generator=g)
    ids = torch.empty((N,), dtype=torch.int64, device="cuda")
    ctx.sample_step(logits.cuda(), 0.8, 0.9, u.cuda(), ids)
    p = O.filtered_distribution(logits[0].numpy(), 0.8, 0.9)
    freq = np.bincount(ids.cpu().numpy(), minlength=V) / N
    assert np.all(np.abs(freq - p) <= 5 * np.sqrt(p * (1 - p) / N) + 1e-9)
    assert np.all(freq[p == 0] == 0)
    # bookkeeping: row 1 already finished -> pad; a row drawing eos becomes finished; seqs[:, pos + 1], ++pos
    B, H = 3, 16
    lg = torch.full((B, V), -30.0)
    lg[0, 5] = lg[1, 6] = lg[2, 7] = 30.0
    pos = torch.tensor([4], dtype=torch.int32, device="cuda")
    uni = torch.full((8, B), 0.5, device="cuda")
    cur = torch.zeros((B,), dtype=torch.int64, device="cuda")
    seqs = torch.zeros((B, 8), dtype=torch.int64, device="cuda")
    fin = torch.tensor([0, 1, 0], dtype=torch.uint8, device="cuda")
    final_h = torch.randn((B, H), device="cuda").to(torch.bfloat16)
    hid = torch.zeros((B, 7, H), dtype=torch.bfloat16, device="cuda")
    ctx.sample_step(lg.cuda(), 1.0, None, uni, cur, seqs, final_h, hid, fin, eos_id=7, pad_id=99, pos_dev=pos)
    assert cur.tolist() == [5, 99, 7] and seqs[:, 5].tolist() == [5, 99, 7] and fin.tolist() == [0, 1, 1]
    assert int(pos) == 5 and torch.equal(hid[:, 4], final_h)


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_generate_sampling_on_device(ctx, dtype):
    """do_sample generate (HF warper chain: temperature, top-k = 50 by default, top-p): reproducible under a seeded
    generator, equal to greedy when the temperature is tiny or top_k = 1 / top_p = 0, and the graph-replayed loop agrees
    with the eagerly launched one and with the stopping_criteria variant given the same uniforms."""
    model, sd, cfg = build_tiny_core(dtype)
    ids, images = oracle_inputs_core(2)
    ids, images = ids.cuda(), images.cuda().to(dtype)
    greedy = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=False)
    cold = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1e-4,
                          generator=torch.Generator("cuda").manual_seed(1))
    assert torch.equal(greedy, cold)
    a = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0, top_p=0.9,
                       generator=torch.Generator("cuda").manual_seed(7))
    b = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0, top_p=0.9,
                       generator=torch.Generator("cuda").manual_seed(7))
    c = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0, top_p=0.9,
                       generator=torch.Generator("cuda").manual_seed(8))
    assert torch.equal(a, b) and a.shape == (2, ids.shape[1] + 8)
    assert torch.equal(a[:, :ids.shape[1]], ids) and int(a.max()) < model.config.vocab_size
    assert not torch.equal(a, c)
    for kw in (dict(top_k=1), dict(top_p=0.0)):          # only the most likely token survives the filter
        one = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0,
                             generator=torch.Generator("cuda").manual_seed(3), **kw)
        assert torch.equal(one, greedy), kw
    model.use_cuda_graph = False
    d = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0, top_p=0.9,
                       generator=torch.Generator("cuda").manual_seed(7))
    model.use_cuda_graph = True
    e = model.generate(input_ids=ids, images=images, max_new_tokens=8, do_sample=True, temperature=1.0, top_p=0.9,
                       generator=torch.Generator("cuda").manual_seed(7), stopping_criteria=[lambda s, sc: False])
    assert torch.equal(a, d) and torch.equal(a, e)
    # evaluate()'s default (temperature = 0.2) also runs on the device path
    out = model.generate(input_ids=ids, images=images, max_new_tokens=4, do_sample=True, temperature=0.2,
                         output_hidden_states=True, return_dict_in_generate=True)
    assert out.sequences.shape == (2, ids.shape[1] + 4) and out.hidden_states[-1][-1].shape[1] == ids.shape[1] + 3

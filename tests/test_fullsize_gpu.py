"""Parity at BASELINE.json's FULL sizes (configs c2 / c3 / c4), where the fp32 CPU oracle of the whole model would
take minutes: size-independent properties of the product path, plus the oracle on the slice it finishes in seconds.

  c2  ViT-L/14-336, 23 layers, batch 32: every image's features are bit-identical to the same image encoded in a
      batch of two (batch invariance), and image 0 matches the CPU oracle (23 full-width layers, fp32).
  c3  LLaMA-7B prefill, 576 visual + 32 text tokens, batch 16: batch invariance, causality (editing the prompt tail
      leaves every earlier position bit-identical), and KV-cached decode == teacher-forced prefill on the sequence it
      generated (token ids wherever the top-2 margin exceeds the logit tolerance).
  c4  full u-LLaVA-7B pipeline, batch 8, 64 greedy tokens + masks: run-to-run determinism (bit-identical ids and mask
      logits), batch invariance of the token ids, mask logits of batch-8 and batch-2 runs equal up to 16-bit rounding.
Synthetic weights of the real architecture (bench.build_model, seed 0), the bench's synthetic inputs."""
import pytest
import torch

import bench
from oracle import ullava_oracle as O

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DT = torch.bfloat16


@pytest.fixture(scope="module")
def full(ctx):
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev, DT)
    yield model
    del model
    torch.cuda.empty_cache()


def _inputs(n):
    ids, img, sam = bench.make_inputs(0, n, DT)
    return ids.cuda(), img.cuda(), sam.cuda()


def test_c2_vit_batch32_invariance_and_oracle(full):
    _, img, _ = _inputs(32)
    feats = full.llm.encode_image(img)
    assert feats.shape == (32, bench.N_PATCH, 1024)
    # an image's features do not depend on the rest of the batch or on its slot in it: bit-identical in a batch of 2
    # (same kernels; a batch of ONE has < 1024 token rows and takes the CTA-per-row LayerNorm, whose fp32 statistics
    # are summed in another order -> compared within the 16-bit rounding it can cause)
    for b in (0, 17, 30):
        pair = full.llm.encode_image(img[b:b + 2])
        assert torch.equal(pair, feats[b:b + 2]), f"images {b}, {b + 1}: batch-32 features differ from the batch-2 run"
    alone = full.llm.encode_image(img[31:32])[0].float()
    rel = (alone - feats[31].float()).norm() / feats[31].float().norm()
    assert rel < 1e-2, f"image 31 alone vs in the batch: relative difference {rel:.4f}"
    # oracle: the same 23 layers in fp32 on the host, on the 16-bit weights the device holds
    sd = {k: v.detach().float().cpu() for k, v in full.llm.vision_encoder.state_dict().items()}
    ref = O.clip_vit_hidden(sd, "", img[:1].float().cpu(), dict(bench.VISION), -2)[:, 1:]
    err = (feats[0].float().cpu() - ref[0]).norm() / ref[0].norm()
    assert err < 2e-2, f"relative error of the 23-layer CLIP features vs the fp32 oracle: {err:.4f}"   # bf16 activations


def test_c3_llama_prefill_b16_invariance_causality_and_cache(full):
    ids, img, _ = _inputs(16)
    out = full.llm(input_ids=ids, images=img, return_dict=True)
    logits = out.logits
    assert logits.shape == (16, bench.P_LEN, bench.VOCAB)
    pair = full.llm(input_ids=ids[5:7], images=img[5:7], return_dict=True).logits
    assert torch.equal(pair, logits[5:7]), "samples 5, 6: batch-16 logits differ from the batch-2 run"
    del pair
    # causality: a different prompt tail changes nothing before it
    ids2 = ids.clone()
    ids2[:, -10:] = torch.randint(3, 32000, (16, 10), device=ids.device)
    logits2 = full.llm(input_ids=ids2, images=img, return_dict=True).logits
    cut = bench.P_LEN - 10
    assert torch.equal(logits2[:, :cut], logits[:, :cut])
    assert not torch.equal(logits2[:, cut:], logits[:, cut:])
    del logits2
    # KV-cached greedy decode vs teacher-forced prefill of the generated sequence
    new = 6
    seqs = full.llm.generate(input_ids=ids, images=img, max_new_tokens=new, do_sample=False)
    assert seqs.shape == (16, bench.P_LEN + new) and torch.equal(seqs[:, :bench.P_LEN], ids)
    tf = full.llm(input_ids=seqs, images=img, return_dict=True).logits[:, bench.P_LEN - 1:-1].float()
    top2 = tf.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    agree = tf.argmax(-1) == seqs[:, bench.P_LEN:]
    tol = 8e-2   # bf16 bar of DESIGN.md section 2 (8 x the 1e-2 fp16 bar)
    assert bool(agree[margin > 2 * tol].all()), "decode step and prefill disagree on a token with a clear margin"
    print(f"c3: decode == prefill argmax on {float(agree.float().mean()):.3f} of the steps "
          f"({int((margin > 2 * tol).sum())} of {margin.numel()} with a margin above {2 * tol})")


def test_c4_full_pipeline_b8_determinism_and_batch_invariance(full):
    ids, img, sam = _inputs(8)
    sizes, resizes = [(bench.IMG, bench.IMG)] * 8, [(bench.SAM_IMG, bench.SAM_IMG)] * 8
    # same execution configuration for both batch sizes: the decode steps on the decode lane of the SM partition (the
    # stream-K split of the weight-streaming GEMMs, hence their fp32 summation order, follows the lane's SM count)
    full.overlap_min_batch = 1
    run = lambda sl: full.evaluate(sam[sl], img[sl], ids[sl], sizes[sl], resizes[sl], max_new_tokens=64, temperature=0)
    seq_a, masks_a, boxes_a = run(slice(0, 8))
    seq_b, masks_b, boxes_b = run(slice(0, 8))
    assert seq_a.shape == (8, bench.P_LEN + 64)
    assert torch.equal(seq_a, seq_b)
    assert all(m.shape[0] >= 1 and m.shape[1:] == (bench.IMG, bench.IMG) for m in masks_a)
    assert all(torch.equal(x, y) for x, y in zip(masks_a, masks_b)), "mask logits are not reproducible run to run"
    seq_2, masks_2, _ = run(slice(3, 5))
    assert torch.equal(seq_2, seq_a[3:5]), "samples 3, 4: batch-8 token ids differ from the batch-2 run"
    for j in range(2):
        assert masks_2[j].shape == masks_a[3 + j].shape
        # The mask decoder sees 6 x n_prompts token rows: 48 rows take the tensor-core GEMM, 12 rows the weight-streaming
        # one, so the logits differ by 16-bit rounding.  Random weights give noise-like masks (no coherent region, a
        # fixed share of pixels within rounding of 0), hence the relative-error bar and the exact sign check outside
        # the band instead of north_star's IoU >= 0.999 (which the tiny-model tests apply to the oracle's masks).
        x, y = masks_2[j].float(), masks_a[3 + j].float()
        rel = ((x - y).norm() / y.norm()).item()
        assert rel < 2e-2, f"mask logits batch-8 vs batch-2: relative difference {rel:.4f}"
        band = 4 * (x - y).abs().mean().item() + 1e-6
        clear = y.abs() > 8 * band
        assert bool(((x > 0) == (y > 0))[clear].all()) and float(clear.float().mean()) > 0.5


def _oracle_cfg():
    cfg = dict(bench.LLM)
    cfg["rope_theta"] = 10000.0
    return cfg


SAM_VITH = dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=[7, 15, 23, 31], window_size=14,
                patch_size=16)


@pytest.fixture(scope="module")
def oracle_run(full):
    """The oracle restatement of the WHOLE path at full size (LLaMA-7B, ViT-L/14-336, SAM ViT-H), evaluated in fp32 ON
    THE GPU on the very weights the product model holds (the CPU would need tens of minutes): B = 2 bench prompts,
    prefill logits, 8 greedy tokens with their margins, SAM embeddings, low-res and final mask logits.

    Yardstick: the same restatement evaluated in bf16 and in fp16 (plain torch ops on 16-bit tensors, i.e. what the
    reference's own HF path computes in that dtype: 16-bit GEMM outputs, fp32 softmax / RMSNorm statistics).  A 32-layer
    random-init decoder amplifies rounding noise, so the distance between the reference's 16-bit path and its fp32 path
    is what a faithful 16-bit implementation can be held to at this size; north_star's 1e-2 (fp16) is met on the tiny
    and single-layer fixtures (tests/test_models_gpu.py) and is not attainable by ANY 16-bit evaluation here."""
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False      # plain fp32 arithmetic: the oracle is the checker
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd32 = {k: v.detach().float() for k, v in full.state_dict().items()}
        ids, img, sam = _inputs(2)
        cfg = _oracle_cfg()
        new = 8
        pre = O.core_forward(sd32, cfg, ids, img.float(), prefix="llm.")
        o_seqs, o_hid, margins = O.greedy_generate(sd32, cfg, ids, img.float(), new, prefix="llm.")
        emb = O.sam_image_encoder(sd32, "visual_model.", sam.float(), SAM_VITH)
        sizes, resizes = [(bench.IMG, bench.IMG)] * 2, [(bench.SAM_IMG, bench.SAM_IMG)] * 2
        scfg = dict(seg_token_idx=bench.SEG_ID, loc_token_idx=bench.LOC_ID)
        pm, pb, low = O.masks_from_hidden(sd32, scfg, o_seqs, o_hid, emb, sizes, resizes)
        out = dict(ids=ids, img=img, sam=sam, new=new, last_logits=pre["logits"][:, -1].clone(),
                   last_hidden=pre["last_hidden"].clone(), seqs=o_seqs, hid=o_hid, margins=margins, emb=emb, pm=pm,
                   low=low, sizes=sizes, resizes=resizes, yard={})
        del pre
        for dt in (torch.bfloat16, torch.float16):
            sd16 = {k: v.to(dt) for k, v in sd32.items() if k.startswith("llm.")}
            y = O.core_forward(sd16, cfg, ids, img.to(dt), prefix="llm.")
            out["yard"][dt] = dict(
                logits=(y["logits"][:, -1].float() - out["last_logits"]).abs().max().item(),
                logits_rms=(y["logits"][:, -1].float() - out["last_logits"]).pow(2).mean().sqrt().item(),
                hidden=(y["last_hidden"].float() - out["last_hidden"]).abs().max().item())
            del sd16, y
        del sd32
        torch.cuda.empty_cache()
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _iou(a, b):
    a, b = a > 0, b > 0
    u = (a | b).sum().item()
    return 1.0 if u == 0 else (a & b).sum().item() / u


def _logit_check(name, dt, logits, hidden, r):
    """last-position prefill logits / post-norm hidden states of the product against the fp32 oracle, held to the
    distance of the reference's own path in that dtype (yardstick) from the same oracle"""
    err = (logits.float() - r["last_logits"]).abs().max().item()
    rms = (logits.float() - r["last_logits"]).pow(2).mean().sqrt().item()
    herr = (hidden.float() - r["last_hidden"]).abs().max().item()
    y = r["yard"][dt]
    print(f"{name} vs fp32 oracle: last-position logits max-abs {err:.4f} (rms {rms:.4f}), hidden max-abs {herr:.4f}; "
          f"the reference arithmetic in the same dtype: logits {y['logits']:.4f} (rms {y['logits_rms']:.4f}), hidden "
          f"{y['hidden']:.4f}; logit std {r['last_logits'].std().item():.3f}")
    assert rms <= 1.25 * y["logits_rms"] + 1e-3, (rms, y)
    assert err <= 1.5 * y["logits"] + 1e-2, (err, y)
    assert herr <= 1.5 * y["hidden"] + 1e-2, (herr, y)
    return max(err, y["logits"])


def test_c4_full_size_bf16_vs_fp32_oracle(full, oracle_run):
    """The benchmarked configuration itself (full u-LLaVA-7B, bf16, evaluate() through the graph-replayed decode loop)
    against the fp32 oracle: greedy ids (exact wherever the oracle's top-2 margin exceeds twice the logit error of a
    bf16 evaluation), last-position prefill logits, hidden states, SAM embeddings, mask logits, thresholded masks."""
    from tests.util_models import greedy_walk
    r = oracle_run
    ids, img, sam, new = r["ids"], r["img"], r["sam"], r["new"]
    full.overlap_min_batch = 1               # the configuration the bench runs: decode on the SM-partition lane
    seqs, masks, _ = full.evaluate(sam, img, ids, r["sizes"], r["resizes"], max_new_tokens=new, temperature=0)
    assert full.llm.graph_kernel_launches() > 0
    out = full.llm(input_ids=ids, images=img, return_dict=True, logits_to_keep=1, logits_fp32=True,
                   _return_last_hidden=True)
    tol = _logit_check("c4 bf16", torch.bfloat16, out.logits[:, 0], out.hidden_states[-1], r)
    print(f"  oracle margins {[round(float(x), 3) for x in r['margins'].flatten()]}")
    checked = 0
    for b in range(2):
        exact, prefix = greedy_walk(seqs[b].cpu(), r["seqs"][b].cpu(), r["margins"][b].cpu(), bench.P_LEN, 2 * tol)
        checked += exact
        print(f"  sample {b}: {exact} ids asserted exactly (margin > {2 * tol:.3f}), common prefix {prefix} of {new}")
    # SAM ViT-H embeddings (32 blocks) and the heads, teacher-forced with the ORACLE's ids / hidden states so that the
    # comparison does not depend on the greedy ids above
    emb = full.get_visual_embs(sam)
    e_rel = ((emb.float() - r["emb"]).norm() / r["emb"].norm()).item()
    assert e_rel < 3e-2, f"SAM ViT-H embeddings: relative error {e_rel:.4f}"
    tm, _, _ = full._decode_heads(r["seqs"], r["hid"].to(DT), emb, r["sizes"], r["resizes"])
    ious = []
    for i in range(2):
        got, ref = tm[i].float(), r["pm"][i]
        assert got.shape == ref.shape and got.shape[0] >= 1
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 6e-2, f"mask logits vs fp32 oracle: relative error {rel:.4f}"
        band = ref.abs() > 0.1 * ref.abs().max()
        agree = ((got > 0) == (ref > 0))[band].float().mean().item()
        assert agree == 1.0, f"mask pixels outside the rounding band disagree ({agree:.6f})"
        ious.append(_iou(got, ref))
    print(f"  SAM embeddings rel err {e_rel:.4f}; mask IoU vs oracle (noise-like random-weight masks) {ious}")
    assert min(ious) > 0.9


def test_c4_full_size_fp16_vs_fp32_oracle(full, oracle_run):
    """The same weights cast to fp16 (north_star states its logit bar for fp16): prefill logits / hidden states against
    the fp32 oracle and the fp16 yardstick, greedy ids where the margin is clear."""
    import copy
    from tests.util_models import greedy_walk
    r = oracle_run
    ids, img, new = r["ids"], r["img"], r["new"]
    tower, stack = full.llm._tower, full.llm._stack          # packed copies / captured graphs are not deep-copyable
    full.llm._tower = full.llm._stack = None
    try:
        llm16 = copy.deepcopy(full.llm).to(torch.float16)
    finally:
        full.llm._tower, full.llm._stack = tower, stack
    try:
        out = llm16(input_ids=ids, images=img.to(torch.float16), return_dict=True, logits_to_keep=1, logits_fp32=True,
                    _return_last_hidden=True)
        tol = _logit_check("c4 fp16", torch.float16, out.logits[:, 0], out.hidden_states[-1], r)
        seqs = llm16.generate(input_ids=ids, images=img.to(torch.float16), max_new_tokens=new, do_sample=False)
        checked = 0
        for b in range(2):
            exact, prefix = greedy_walk(seqs[b].cpu(), r["seqs"][b].cpu(), r["margins"][b].cpu(), bench.P_LEN, 2 * tol)
            checked += exact
            print(f"  sample {b}: {exact} ids asserted exactly (margin > {2 * tol:.3f}), common prefix {prefix} of {new}")
        assert checked >= 4, "fp16: too few generated tokens had a clear margin"
    finally:
        del llm16
        torch.cuda.empty_cache()

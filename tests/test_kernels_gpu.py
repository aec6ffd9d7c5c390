"""GPU parity of every primitive kernel behind the C ABI against a plain torch fp32 reference of the
same op (floating-point kernels; tolerances written next to each check)."""

import pytest
import torch

import native
from native import EPI_NONE, EPI_RELU, EPI_GELU, EPI_QUICK_GELU, EPI_SILU_MUL

pytestmark = pytest.mark.gpu
DT = [torch.bfloat16, torch.float16]


def _rand(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


def _tol(dtype):
    # one rounding of the output to the 16-bit type: 2^-8 (bf16) / 2^-11 (fp16) relative
    return (1.0 / 128, 2e-2) if dtype == torch.bfloat16 else (1.0 / 1024, 4e-3)


def _check(out, ref, dtype, what, extra_abs=0.0):
    rtol, atol = _tol(dtype)
    out = out.float()
    err = (out - ref).abs()
    lim = atol + extra_abs + rtol * ref.abs()
    bad = (err > lim)
    assert not bool(bad.any()), f"{what}: max err {err.max().item():.4g} (ref max {ref.abs().max().item():.4g}), " \
                                f"{int(bad.sum())} / {bad.numel()} out of tolerance"


def _act_ref(x, epi):
    if epi == EPI_RELU:
        return torch.relu(x)
    if epi == EPI_GELU:
        return torch.nn.functional.gelu(x)
    if epi == EPI_QUICK_GELU:
        return x * torch.sigmoid(1.702 * x)
    return x


# ------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("bn", [16, 32, 64, 128, 256])
def test_gemm_single_tile_each_bn(ctx, dtype, bn):
    a = _rand((128, 64), dtype, seed=1)
    w = _rand((bn, 64), dtype, seed=2)
    out = ctx.gemm(a, w, force_bn=bn, no_swap=True)
    _check(out, a.float() @ w.float().t(), dtype, f"gemm 128x{bn}x64")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (300, 200, 136), (1000, 1032, 4096), (577 * 3, 1024, 1024),
                                   (129, 257 * 8, 72), (4096, 4096, 1024), (64, 64, 2048)])
def test_gemm_shapes(ctx, dtype, M, N, K):
    a = _rand((M, K), dtype, 1.0, seed=3)
    w = _rand((N, K), dtype, K ** -0.5, seed=4)
    out = ctx.gemm(a, w)
    _check(out, a.float() @ w.float().t(), dtype, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K,epi", [(4096, 4096, 1024, EPI_NONE), (5000, 2000, 136, EPI_GELU),
                                       (9728, 1024, 4096, EPI_NONE), (4933, 2312, 1288, EPI_QUICK_GELU),
                                       (37 * 256 + 1, 1280, 64, EPI_RELU)])
def test_gemm_cta_pair_kernel(ctx, dtype, M, N, K, epi):
    """Shapes with >= 148 tiles of 256 x 256 run on CTA pairs (tcgen05 cta_group::2): ragged M (second CTA of a pair
    partly or fully out of range), ragged N and K, every fused epilogue, in-place residual."""
    a = _rand((M, K), dtype, seed=41)
    w = _rand((N, K), dtype, K ** -0.5, seed=42)
    b = _rand((N,), dtype, seed=43)
    r = _rand((M, N), dtype, seed=44)
    ref = _act_ref(a.float() @ w.float().t() + b.float(), epi) + r.float()
    out = ctx.gemm(a, w, bias=b, residual=r, epilogue=epi)
    _check(out, ref, dtype, f"pair gemm {M}x{N}x{K} epi {epi}")
    single = ctx.gemm(a, w, bias=b, residual=r, epilogue=epi, force_bn=256)   # single-CTA kernel, same K order
    assert torch.equal(out, single), "CTA-pair and single-CTA kernels must round identically"
    r2 = r.clone()
    ctx.gemm(a, w, bias=b, residual=r2, epilogue=epi, out=r2)
    assert torch.equal(r2, out)


@pytest.mark.parametrize("dtype", DT)
def test_gemm_cta_pair_silu_mul(ctx, dtype):
    M, K, F = 4864, 512, 2048
    a = _rand((M, K), dtype, seed=45)
    wg = _rand((F, K), dtype, K ** -0.5, seed=46)
    wu = _rand((F, K), dtype, K ** -0.5, seed=47)
    packed = torch.stack([wg.view(F // 16, 16, K), wu.view(F // 16, 16, K)], dim=1).reshape(2 * F, K).contiguous()
    out = ctx.gemm(a, packed, epilogue=EPI_SILU_MUL)
    g, u = a.float() @ wg.float().t(), a.float() @ wu.float().t()
    _check(out, torch.nn.functional.silu(g) * u, dtype, "pair silu_mul")
    assert torch.equal(out, ctx.gemm(a, packed, epilogue=EPI_SILU_MUL, force_bn=256))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("epi", [EPI_NONE, EPI_RELU, EPI_GELU, EPI_QUICK_GELU])
def test_gemm_bias_act_residual(ctx, dtype, epi):
    M, N, K = 700, 1024, 512
    a = _rand((M, K), dtype, seed=5)
    w = _rand((N, K), dtype, K ** -0.5, seed=6)
    b = _rand((N,), dtype, seed=7)
    r = _rand((M, N), dtype, seed=8)
    out = ctx.gemm(a, w, bias=b, residual=r, epilogue=epi)
    ref = _act_ref(a.float() @ w.float().t() + b.float(), epi) + r.float()
    _check(out, ref, dtype, f"gemm epi {epi}")
    # in-place residual (D aliases residual), as used by the o_proj / down_proj / fc2 calls
    r2 = r.clone()
    ctx.gemm(a, w, bias=b, residual=r2, epilogue=epi, out=r2)
    _check(r2, ref, dtype, f"gemm epi {epi} in-place")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M", [8, 32, 200])
def test_gemm_silu_mul(ctx, dtype, M):
    K, F = 512, 1024  # packed weight has 2F rows: per 32-row chunk 16 gate rows then 16 up rows
    a = _rand((M, K), dtype, seed=9)
    wg = _rand((F, K), dtype, K ** -0.5, seed=10)
    wu = _rand((F, K), dtype, K ** -0.5, seed=11)
    packed = torch.stack([wg.view(F // 16, 16, K), wu.view(F // 16, 16, K)], dim=1).reshape(2 * F, K).contiguous()
    out = ctx.gemm(a, packed, epilogue=EPI_SILU_MUL)
    g = a.float() @ wg.float().t()
    u = a.float() @ wu.float().t()
    _check(out, torch.nn.functional.silu(g) * u, dtype, f"silu_mul M={M}")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K", [(1, 4096, 4096), (8, 12288, 4096), (32, 4096, 11008), (16, 1024, 256), (5, 256, 2048)])
def test_gemm_small_m_swap_ab(ctx, dtype, M, N, K):
    a = _rand((M, K), dtype, seed=12)
    w = _rand((N, K), dtype, K ** -0.5, seed=13)
    b = _rand((N,), dtype, seed=14)
    r = _rand((M, N), dtype, seed=15)
    out = ctx.gemm(a, w, bias=b, residual=r)
    ref = a.float() @ w.float().t() + b.float() + r.float()
    _check(out, ref, dtype, f"swap gemm {M}x{N}x{K}")
    out2 = ctx.gemm(a, w, bias=b, residual=r, no_swap=True)
    _check(out2, ref, dtype, f"no-swap gemm {M}x{N}x{K}")


@pytest.mark.parametrize("pdl", [True, False])
@pytest.mark.parametrize("M,N,K", [(32, 4096, 4096), (3, 300, 200), (17, 1000, 72), (32, 128, 64), (1, 128, 8192),
                                   (9, 22016, 4096), (32, 32011, 512)])
def test_gemm_stream_k_partition(ctx, pdl, M, N, K):
    """Weight-streaming kernel: ragged N / K, a single unit, one tile cut across many CTAs (K = 8192, N = 128),
    many tiles per CTA; repeated launches prove that the tile counters return to zero."""
    dtype = torch.bfloat16
    a = _rand((M, K), dtype, seed=21)
    w = _rand((N, K), dtype, K ** -0.5, seed=22)
    b = _rand((N,), dtype, seed=23)
    ref = torch.relu(a.float() @ w.float().t() + b.float())
    ctx.set_pdl(pdl)
    try:
        for it in range(3):
            out = ctx.gemm(a, w, bias=b, epilogue=EPI_RELU)
            _check(out, ref, dtype, f"stream gemm {M}x{N}x{K} launch {it}")
    finally:
        ctx.set_pdl(True)


def test_gemm_stream_after_norm_chain(ctx):
    """The decode chain rmsnorm -> GEMM -> GEMM(silu*mul) -> GEMM(+residual in place) back to back: with programmatic
    dependent launch each GEMM prefetches weights under its predecessor and must still see its finished output."""
    dtype = torch.bfloat16
    B, H, F = 32, 1024, 2048
    x = _rand((B, H), dtype, seed=24)
    g = (_rand((H,), dtype, seed=25).float() * 0.1 + 1).to(dtype)
    w1 = _rand((H, H), dtype, H ** -0.5, seed=26)
    wg = _rand((F, H), dtype, H ** -0.5, seed=27)
    wu = _rand((F, H), dtype, H ** -0.5, seed=28)
    wgu = torch.stack([wg.view(F // 16, 16, H), wu.view(F // 16, 16, H)], dim=1).reshape(2 * F, H).contiguous()
    wd = _rand((H, F), dtype, F ** -0.5, seed=29)
    outs = []
    for _ in range(4):
        hid = x.clone()
        xn = ctx.rmsnorm(hid, g, 1e-6)
        h1 = ctx.gemm(xn, w1)
        act = ctx.gemm(h1, wgu, epilogue=EPI_SILU_MUL)
        ctx.gemm(act, wd, residual=hid, out=hid)
        outs.append(hid)
    xf = x.float()
    xn_ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * g.float()).to(dtype).float()
    h1_ref = (xn_ref @ w1.float().t()).to(dtype).float()
    act_ref = (torch.nn.functional.silu(h1_ref @ wg.float().t()) * (h1_ref @ wu.float().t())).to(dtype).float()
    ref = act_ref @ wd.float().t() + xf
    for o in outs:
        _check(o, ref, dtype, "norm->gemm chain", extra_abs=2e-2)
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("tiles", [0, 12, 256])
@pytest.mark.parametrize("nn,nk", [(4096, 4096), (300, 200), (128, 64), (22016, 4096)])
def test_gemm_stream_next_weight_hint(ctx, tiles, nn, nk):
    """ullava_gemm_next_weight is a pure hint: the L2 prefetch of the following GEMM's weights (ragged shapes, fewer
    units than SMs, more tiles asked for than the next CTA's range holds) never changes a bit of either product."""
    dtype = torch.bfloat16
    M, N, K = 32, 4096, 1024
    a = _rand((M, K), dtype, seed=31)
    w = _rand((N, K), dtype, K ** -0.5, seed=32)
    wn = _rand((nn, nk), dtype, nk ** -0.5, seed=33)
    a2 = _rand((M, nk), dtype, seed=34)
    base = ctx.gemm(a, w)
    base2 = ctx.gemm(a2, wn)
    _check(base, a.float() @ w.float().t(), dtype, "stream gemm")
    ctx.set_weight_prefetch(tiles)
    try:
        for _ in range(2):
            ctx.gemm_next_weight(wn)
            out = ctx.gemm(a, w)
            out2 = ctx.gemm(a2, wn)   # the hint was consumed by the call above: this launch prefetches nothing
            assert torch.equal(out, base) and torch.equal(out2, base2)
        ctx.gemm_next_weight(wn)
        big = ctx.gemm(_rand((200, K), dtype, seed=35), w)   # large-M path drops the hint
        assert big.shape == (200, N)
        with pytest.raises(RuntimeError):
            ctx.gemm_next_weight(wn[:, 1:])                   # misaligned / non-contiguous rows are rejected
    finally:
        ctx.set_weight_prefetch(12)
        ctx.gemm_next_weight(None)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M", [4, 70])
def test_gemm_lm_head_fp32_out(ctx, dtype, M):
    N, K = 32011, 512  # vocab of the reference tokenizer: odd N, fp32 logits
    a = _rand((M, K), dtype, seed=16)
    w = _rand((N, K), dtype, K ** -0.5, seed=17)
    out = ctx.gemm(a, w, out_f32=True)
    assert out.dtype == torch.float32 and out.shape == (M, N)
    ref = a.float() @ w.float().t()
    assert (out - ref).abs().max().item() < 2e-3  # fp32 accumulate + fp32 store: only summation order differs


@pytest.mark.parametrize("dtype", DT)
def test_gemm_split_k_large_m(ctx, dtype):
    M, N, K = 256, 256, 4096
    a = _rand((M, K), dtype, seed=18)
    w = _rand((N, K), dtype, K ** -0.5, seed=19)
    b = _rand((N,), dtype, seed=20)
    out = ctx.gemm(a, w, bias=b, epilogue=EPI_RELU, force_splits=4)
    _check(out, torch.relu(a.float() @ w.float().t() + b.float()), dtype, "split-k")


def test_gemm_strided_operands(ctx):
    dtype = torch.bfloat16
    big = _rand((300, 1536), dtype, seed=21)
    a = big[:, 256:512]  # lda = 1536
    w = _rand((128, 256), dtype, 1 / 16, seed=22)
    outbuf = torch.zeros((300, 512), dtype=dtype, device="cuda")
    ctx.gemm(a, w, out=outbuf[:, 128:256])
    _check(outbuf[:, 128:256], a.float() @ w.float().t(), dtype, "strided gemm")
    assert float(outbuf[:, :128].abs().max()) == 0 and float(outbuf[:, 256:].abs().max()) == 0


def test_gemm_rejects_bad_arguments(ctx):
    a = _rand((16, 60), torch.bfloat16)  # K=60 -> lda not a multiple of 8
    w = _rand((16, 60), torch.bfloat16)
    with pytest.raises(native.NativeError):
        ctx.gemm(a, w)


# ------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,cols", [(577, 1024), (6, 256), (4096, 64), (33, 4096), (4100, 1280), (1031, 1024),
                                       (2048, 256), (1025, 4096)])   # >= 1024 rows: warp-per-row kernel
def test_layernorm(ctx, dtype, rows, cols):
    x = _rand((rows, cols), dtype, 2.0, seed=30) + 0.5
    w = _rand((cols,), dtype, seed=31)
    b = _rand((cols,), dtype, seed=32)
    out = ctx.layernorm(x, w, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-5)
    _check(out, ref, dtype, "layernorm")
    out = ctx.layernorm(x, w, b, 1e-6, act=EPI_GELU)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-6))
    _check(out, ref, dtype, "layernorm+gelu")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,cols", [(608, 4096), (8, 4096), (5, 128), (1216, 4096), (3000, 128)])
def test_rmsnorm(ctx, dtype, rows, cols):
    x = _rand((rows, cols), dtype, 3.0, seed=33)
    w = _rand((cols,), dtype, seed=34)
    out = ctx.rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    _check(out, ref, dtype, "rmsnorm", extra_abs=1e-2)


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, causal, q_pos0, scale):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * scale
    if causal:
        sq, sk = s.shape[-2:]
        qi = torch.arange(sq, device=s.device)[:, None] + q_pos0
        kj = torch.arange(sk, device=s.device)[None, :]
        s = s.masked_fill(kj > qi, float("-inf"))
    return (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,Sq,Sk,D,causal,pos0", [
    (2, 16, 577, 577, 64, False, 0),     # CLIP ViT-L/14-336
    (2, 4, 608, 608, 128, True, 0),      # LLaMA prefill
    (1, 2, 100, 164, 128, True, 64),     # chunked prefill against a longer cache
    (3, 8, 6, 4096, 16, False, 0),       # SAM token->image
    (1, 8, 4096, 6, 16, False, 0),       # SAM image->token
    (2, 8, 6, 6, 32, False, 0),          # SAM token self attention
    (1, 2, 64, 64, 64, True, 0), (1, 1, 1, 70, 64, False, 0),
])
def test_flash_attention(ctx, dtype, B, H, Sq, Sk, D, causal, pos0):
    # q/k/v as strided views of packed buffers, like the packed-QKV GEMM output
    qkv = _rand((B, max(Sq, Sk), 3, H, D), dtype, seed=40)
    q, k, v = qkv[:, :Sq, 0], qkv[:, :Sk, 1], qkv[:, :Sk, 2]
    out = ctx.attention(q, k, v, causal=causal, q_pos0=pos0)
    ref = _attn_ref(q, k, v, causal, pos0, D ** -0.5)
    _check(out, ref, dtype, f"attention {B,H,Sq,Sk,D,causal}")


@pytest.fixture
def legacy_attention(ctx):
    ctx.set_attention_impl(1)
    yield ctx
    ctx.set_attention_impl(0)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,Sq,Sk,D,causal,pos0", [
    (2, 16, 577, 577, 64, False, 0), (2, 4, 608, 608, 128, True, 0), (1, 2, 100, 164, 128, True, 64),
])
def test_flash_attention_warp_level_kernel(legacy_attention, dtype, B, H, Sq, Sk, D, causal, pos0):
    """The mma.sync kernels stay selectable (ullava_set_attention_impl) for A/B measurements; keep them correct."""
    qkv = _rand((B, max(Sq, Sk), 3, H, D), dtype, seed=41)
    q, k, v = qkv[:, :Sq, 0], qkv[:, :Sk, 1], qkv[:, :Sk, 2]
    out = legacy_attention.attention(q, k, v, causal=causal, q_pos0=pos0)
    _check(out, _attn_ref(q, k, v, causal, pos0, D ** -0.5), dtype, f"legacy attention {B,H,Sq,Sk,D,causal}")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,Sq,Sk,D,causal,pos0", [
    (1, 1, 128, 128, 64, False, 0),      # exactly one tile, one 128-byte slab
    (1, 1, 128, 128, 128, False, 0),     # two slabs
    (1, 1, 128, 128, 80, False, 0),      # 64-column slab + 16-column SWIZZLE_32B tail
    (1, 2, 128, 384, 64, False, 0),      # three key tiles: ring + online softmax
    (2, 3, 300, 1000, 80, False, 0),     # ragged rows / keys
    (1, 2, 640, 640, 128, True, 0),      # causal, diagonal tiles
    (2, 2, 33, 2049, 128, True, 2016),   # chunked prefill deep into a cache
    (1, 1, 256, 4096, 64, False, 0),     # long key sequence: lazy rescale over 32 tiles
])
def test_flash_attention_tcgen05_tiles(ctx, dtype, B, H, Sq, Sk, D, causal, pos0):
    qkv = _rand((B, max(Sq, Sk), 3, H, D), dtype, seed=42)
    q, k, v = qkv[:, :Sq, 0], qkv[:, :Sk, 1], qkv[:, :Sk, 2]
    ctx.set_attention_impl(2)
    try:
        out = ctx.attention(q, k, v, causal=causal, q_pos0=pos0)
    finally:
        ctx.set_attention_impl(0)
    _check(out, _attn_ref(q, k, v, causal, pos0, D ** -0.5), dtype, f"tcgen05 attention {B,H,Sq,Sk,D,causal}")


def test_flash_attention_rescale_growing_maximum(ctx):
    """Scores that keep growing along the key axis force the O / row-sum rescale on every tile."""
    B, H, S, D = 1, 2, 1024, 64
    q = _rand((B, S, H, D), torch.bfloat16, seed=43)
    k = _rand((B, S, H, D), torch.bfloat16, seed=44)
    v = _rand((B, S, H, D), torch.bfloat16, seed=45)
    ramp = torch.linspace(0.2, 6.0, S, device="cuda")[None, :, None, None]
    k = (k.float() * 0.1 + q.float().mean(1, keepdim=True) * ramp).to(torch.bfloat16)
    out = ctx.attention(q, k, v)
    _check(out, _attn_ref(q, k, v, False, 0, D ** -0.5), torch.bfloat16, "growing-max attention")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("D", [64, 80, 128])
@pytest.mark.parametrize("Sk", [100, 200, 300, 500, 897, 2048])
def test_flash_attention_maximum_handed_from_tile_to_tile(ctx, dtype, D, Sk):
    """The two softmax warpgroups take alternating key tiles and hand the row maximum on (fmha_sm100.cu, ALT): keys whose
    scores climb by ~12 log2 units per tile force the lazy rescale on (nearly) every tile, in both groups; every row
    also has tiles whose scores lie far below the maximum in use (partial sums converted late), and 1 ... 16 tiles cover
    'the other group never ran' and both parities of the last tile."""
    B, H, Sq = 2, 3, 200
    g = torch.Generator(device="cuda").manual_seed(43 + Sk + D)
    q = torch.randn((B, Sq, H, D), generator=g, device="cuda")
    k = torch.randn((B, Sk, H, D), generator=g, device="cuda")
    v = torch.randn((B, Sk, H, D), generator=g, device="cuda")
    # one strong common direction: q.k grows with the key index for half of the rows and falls for the other half
    u = torch.randn((D,), generator=g, device="cuda")
    u = u / u.norm()
    ramp = torch.linspace(0.0, 1.0, Sk, device="cuda") * (Sk / 128.0) * 12.0 * 0.6931 * D ** 0.5
    sign = torch.where(torch.arange(Sq, device="cuda") % 2 == 0, 1.0, -1.0)
    q = (q + sign[None, :, None, None] * 1.0 * u).to(dtype)
    k = (k + ramp[None, :, None, None] * u).to(dtype)
    v = v.to(dtype)
    ctx.set_attention_impl(2)
    try:
        out = ctx.attention(q, k, v, causal=False, q_pos0=0)
    finally:
        ctx.set_attention_impl(0)
    ref = _attn_ref(q, k, v, False, 0, D ** -0.5)
    assert torch.isfinite(out.float()).all()
    _check(out, ref, dtype, f"handed-on maximum {Sk, D}")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,S,D,causal", [
    (7, 9, 577, 64, False),      # 315 items of 5 tiles on 148 persistent CTAs: every CTA crosses item boundaries
    (3, 16, 1400, 128, True),    # causal: items of 1 ... 11 tiles, odd and even tile counts follow each other
    (2, 3, 700, 80, False),      # fewer items than SMs: one item per CTA through the same loop
    (5, 31, 130, 64, True),      # two-tile items whose second tile is almost empty
])
def test_flash_attention_persistent_items(ctx, dtype, B, H, S, D, causal):
    """The bias-free kernels run one CTA per SM over (query tile, head, batch) items with the K / V ring, the S buffers
    and the softmax alternation carried across items (fmha_sm100.cu, fm_item): check the item boundaries."""
    qkv = _rand((B, S, 3, H, D), dtype, seed=52)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ctx.set_attention_impl(2)
    try:
        out = ctx.attention(q, k, v, causal=causal)
        again = ctx.attention(q, k, v, causal=causal)
    finally:
        ctx.set_attention_impl(0)
    _check(out, _attn_ref(q, k, v, causal, 0, D ** -0.5), dtype, f"persistent attention {B,H,S,D,causal}")
    assert torch.equal(out, again)


@pytest.mark.parametrize("D,Sk,causal", [(64, 4096, False), (128, 1664, True), (80, 1000, False)])
def test_flash_attention_is_bit_reproducible(ctx, D, Sk, causal):
    """The pipeline has nine asynchronous parties per CTA (TMA, two MMA issuers, eight softmax warps); a missed
    dependency shows up as run-to-run differences long before it shows up as a tolerance failure."""
    qkv = _rand((2, Sk, 3, 4, D), torch.bfloat16, seed=49)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    first = ctx.attention(q, k, v, causal=causal).clone()
    for _ in range(12):
        assert torch.equal(ctx.attention(q, k, v, causal=causal), first)


def _relpos_ref(q, k, v, rel_h, rel_w, S, scale):
    """segment_anything/modeling/image_encoder.py:196-260 + add_decomposed_rel_pos (:355-392), fp32."""
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))           # [B, H, N, D]
    attn = qf @ kf.transpose(-1, -2) * scale
    idx = torch.arange(S, device=q.device)
    rel = idx[:, None] - idx[None, :] + (S - 1)
    Rh, Rw = rel_h.float()[rel], rel_w.float()[rel]                            # [S, S, D]
    B, H, N, D = qf.shape
    r_q = qf.reshape(B, H, S, S, D)
    bh = torch.einsum("bnhwc,hkc->bnhwk", r_q, Rh)
    bw = torch.einsum("bnhwc,wkc->bnhwk", r_q, Rw)
    attn = (attn.view(B, H, S, S, S, S) + bh[..., :, None] + bw[..., None, :]).view(B, H, N, N)
    return (torch.softmax(attn, -1) @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("packed_tables", [False, True])
@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,S,D", [(3, 2, 14, 80), (1, 2, 64, 80), (2, 2, 9, 64), (50, 16, 14, 80)])
def test_attention_relpos(ctx, dtype, impl, packed_tables, B, H, S, D):
    N = S * S
    qkv = _rand((B, N, 3, H, D), dtype, seed=46)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    rel_h = _rand((2 * S - 1, D), dtype, scale=0.5, seed=47)
    rel_w = _rand((2 * S - 1, D), dtype, scale=0.5, seed=48)
    if packed_tables:   # [Rh; Rw] in one buffer, as the model packs them (enables the single-pass window kernel)
        both = torch.cat([rel_h, rel_w]).contiguous()
        rel_h, rel_w = both[:2 * S - 1], both[2 * S - 1:]
    ctx.set_attention_impl(impl)
    try:
        out = ctx.attention_relpos(q, k, v, rel_h, rel_w, S)
        # scattered output rows (window_unpartition): reverse the row order, drop every 7th row
        rows = torch.arange(B * N, device="cuda", dtype=torch.int32).flip(0).contiguous()
        rows[::7] = -1
        scat = torch.zeros((B * N, H, D), dtype=dtype, device="cuda")
        ctx.attention_relpos(q, k, v, rel_h, rel_w, S, out=scat.view(B, N, H, D), o_row_map=rows)
    finally:
        ctx.set_attention_impl(0)
    ref = _relpos_ref(q, k, v, rel_h, rel_w, S, D ** -0.5)
    _check(out, ref, dtype, f"relpos attention impl={impl} {B,H,S,D}")
    keep = rows >= 0
    exp = torch.zeros_like(scat)
    exp[rows[keep].long()] = out.reshape(B * N, H, D)[keep]
    assert torch.equal(scat, exp)


@pytest.mark.parametrize("S", [14, 64])
def test_attention_relpos_is_bit_reproducible(ctx, S):
    N, D = S * S, 80
    qkv = _rand((3, N, 3, 4, D), torch.bfloat16, seed=50)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    both = _rand((2 * (2 * S - 1), D), torch.bfloat16, scale=0.5, seed=51)
    rel_h, rel_w = both[:2 * S - 1], both[2 * S - 1:]
    first = ctx.attention_relpos(q, k, v, rel_h, rel_w, S).clone()
    for _ in range(12):
        assert torch.equal(ctx.attention_relpos(q, k, v, rel_h, rel_w, S), first)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,D,ctx_len,max_seq", [(2, 32, 128, 609, 672), (32, 32, 128, 672, 672), (3, 4, 64, 5, 16),
                                                   (1, 2, 16, 33, 40), (2, 2, 32, 1, 8)])
def test_decode_attention(ctx, dtype, B, H, D, ctx_len, max_seq):
    kc = _rand((B, H, max_seq, D), dtype, seed=41)
    vc = _rand((B, H, max_seq, D), dtype, seed=42)
    qkv = _rand((B, 3 * H * D), dtype, seed=43)
    out = ctx.attention_decode(qkv, kc, vc, ctx_len)
    q = qkv[:, :H * D].view(B, 1, H, D)
    ref = _attn_ref(q, kc[:, :, :ctx_len].permute(0, 2, 1, 3), vc[:, :, :ctx_len].permute(0, 2, 1, 3), False, 0,
                    D ** -0.5)
    _check(out.view(B, 1, H, D), ref, dtype, "decode attention")


# ------------------------------------------------------------------------------------------
# rope + kv cache, glue
# ------------------------------------------------------------------------------------------
def _rope_tables(max_pos, hd, theta=10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    fr = torch.arange(max_pos).float()[:, None] * inv[None, :]
    return fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,D,ctx_len,max_seq", [(2, 32, 128, 609, 672), (32, 32, 128, 672, 672), (3, 4, 64, 5, 16),
                                                   (2, 2, 16, 1, 8), (1, 2, 32, 8, 8)])
def test_decode_attention_fused_rope_matches_two_kernels(ctx, dtype, B, H, D, ctx_len, max_seq):
    """ullava_attention_decode_rope == ullava_rope_kvcache (one new token) followed by ullava_attention_decode: same
    cache contents and the same output, with the position given by value and through device memory (graph replay)."""
    g = torch.Generator(device="cuda").manual_seed(ctx_len * 7 + D)
    qkv = torch.randn((B, 3 * H * D), generator=g, device="cuda").to(dtype)
    kc = torch.randn((B, H, max_seq, D), generator=g, device="cuda").to(dtype)
    vc = torch.randn((B, H, max_seq, D), generator=g, device="cuda").to(dtype)
    pos = ctx_len - 1
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, device="cuda").float() / D))
    fr = torch.arange(max_seq, device="cuda").float()[:, None] * inv[None, :]
    cos, sin = fr.cos().contiguous(), fr.sin().contiguous()
    qkv_a, kc_a, vc_a = qkv.clone(), kc.clone(), vc.clone()
    ctx.rope_kvcache(qkv_a, kc_a, vc_a, B, 1, pos, cos, sin)
    ref = ctx.attention_decode(qkv_a[:, :H * D], kc_a, vc_a, ctx_len)
    for use_dev in (False, True):
        kc_b, vc_b = kc.clone(), vc.clone()
        pos_dev = torch.tensor([pos], dtype=torch.int32, device="cuda") if use_dev else None
        out = ctx.attention_decode_rope(qkv, kc_b, vc_b, 0 if use_dev else ctx_len, cos, sin, pos_dev=pos_dev)
        assert torch.equal(kc_b, kc_a) and torch.equal(vc_b, vc_a), "fused kernel wrote a different cache row"
        assert torch.equal(out, ref), (out.float() - ref.float()).abs().max().item()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,S,H,D,pos0", [(2, 37, 4, 128, 0), (3, 1, 2, 64, 11), (1, 5, 2, 16, 3)])
def test_rope_kvcache(ctx, dtype, B, S, H, D, pos0):
    max_seq = pos0 + S + 3
    cos, sin = _rope_tables(max_seq, D)
    qkv = _rand((B * S, 3 * H * D), dtype, seed=50)
    orig = qkv.clone()
    kc = torch.zeros((B, H, max_seq, D), dtype=dtype, device="cuda")
    vc = torch.zeros_like(kc)
    ctx.rope_kvcache(qkv, kc, vc, B, S, pos0, cos, sin)
    o = orig.float().view(B, S, 3, H, D)
    c = torch.cat([cos[pos0:pos0 + S], cos[pos0:pos0 + S]], -1)[None, :, None, :]
    s = torch.cat([sin[pos0:pos0 + S], sin[pos0:pos0 + S]], -1)[None, :, None, :]

    def rot(x):
        x1, x2 = x[..., :D // 2], x[..., D // 2:]
        return x * c + torch.cat([-x2, x1], -1) * s

    _check(qkv.view(B, S, 3, H, D)[:, :, 0], rot(o[:, :, 0]), dtype, "rope q")
    _check(kc[:, :, pos0:pos0 + S].permute(0, 2, 1, 3), rot(o[:, :, 1]), dtype, "rope k -> cache")
    assert torch.equal(vc[:, :, pos0:pos0 + S].permute(0, 2, 1, 3), orig.view(B, S, 3, H, D)[:, :, 2])
    assert torch.equal(qkv.view(B, S, 3, H, D)[:, :, 2], orig.view(B, S, 3, H, D)[:, :, 2])
    assert float(kc[:, :, :pos0].abs().sum()) == 0 and float(kc[:, :, pos0 + S:].abs().sum()) == 0


@pytest.mark.parametrize("dtype", DT)
def test_im2col_matches_conv2d(ctx, dtype):
    B, img, patch, out_ch = 2, 28, 14, 64
    k = 3 * patch * patch
    k_pad = (k + 63) // 64 * 64
    px = _rand((B, 3, img, img), dtype, seed=60)
    wconv = _rand((out_ch, 3, patch, patch), dtype, k ** -0.5, seed=61)
    col = ctx.vit_im2col(px, patch, k_pad)
    assert float(col[:, k:].abs().max()) == 0
    wp = torch.zeros((out_ch, k_pad), dtype=dtype, device="cuda")
    wp[:, :k] = wconv.view(out_ch, -1)
    out = ctx.gemm(col, wp)
    ref = torch.nn.functional.conv2d(px.float(), wconv.float(), stride=patch).flatten(2).transpose(1, 2).reshape(-1, out_ch)
    _check(out, ref, dtype, "patchify")


@pytest.mark.parametrize("dtype", DT)
def test_glue_kernels_bit_exact(ctx, dtype):
    # index / copy work: bit-exact
    B, L, dim, vocab, n_patch = 3, 20, 64, 50, 6
    table = _rand((vocab, dim), dtype, seed=70)
    ids = torch.randint(0, vocab, (B, L), device="cuda")
    emb = ctx.embed_gather(ids, table).view(B, L, dim)
    assert torch.equal(emb, table[ids])
    feats = _rand((B, n_patch, dim), dtype, seed=71)
    start = torch.tensor([2, -1, 10], dtype=torch.int32, device="cuda")
    ref = emb.clone()
    ref[0, 3:3 + n_patch] = feats[0]
    ref[2, 11:11 + n_patch] = feats[2]
    ctx.splice_rows(emb, feats, start)
    assert torch.equal(emb, ref)
    pe = _rand((B * 4, dim), dtype, seed=72)
    cls = _rand((dim,), dtype, seed=73)
    pos = _rand((5, dim), dtype, seed=74)
    x = ctx.vit_assemble(pe, cls, pos, B)
    refx = (torch.cat([cls.float().expand(B, 1, dim), pe.float().view(B, 4, dim)], 1) + pos.float()).to(dtype)
    assert torch.equal(x, refx)


def test_argmax_first_index(ctx):
    logits = torch.randn((7, 32011), device="cuda")
    logits[3, 100] = 50.0
    logits[3, 20000] = 50.0
    out = ctx.argmax(logits)
    assert torch.equal(out, logits.argmax(-1)) or out[3].item() == 100
    assert out[3].item() == 100


# ------------------------------------------------------------------------------------------
# SAM post-process
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("inp,orig", [((1024, 1024), (336, 336)), ((768, 1024), (480, 640)), ((1024, 683), (500, 333))])
def test_sam_postprocess(ctx, dtype, inp, orig):
    n = 3
    masks = _rand((n, 4, 256, 256), dtype, 4.0, seed=80)
    out, bits = ctx.sam_postprocess(masks, 4 * 256 * 256, n, 256, 1024, inp, orig, pack_bits=True)
    m0 = masks[:, 0:1].float()
    ref = torch.nn.functional.interpolate(m0, (1024, 1024), mode="bilinear", align_corners=False)
    ref = ref[..., :inp[0], :inp[1]]
    ref = torch.nn.functional.interpolate(ref, orig, mode="bilinear", align_corners=False)[:, 0]
    assert (out - ref).abs().max().item() < 1e-4  # fp32 both sides; only FMA contraction differs
    # packed threshold bits agree wherever the logit is not within rounding noise of 0
    flat = (out.reshape(n, -1) > 0)
    idx = torch.arange(flat.shape[1], device="cuda")
    unpacked = ((bits[:, idx // 32] >> (idx % 32)) & 1).bool()
    assert torch.equal(unpacked, flat)

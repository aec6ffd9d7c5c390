"""ctypes binding of libullava_sm100.so (C ABI declared in include/ullava_sm100.h).

This module is the ONLY place where the host-side mirror of the reference API
(u-llava_b200/models/*) touches native code.  PyTorch is used for device memory and
streams only: every wrapper passes raw device pointers + shapes to the C ABI.

There is no CPU or torch fallback: if the shared library is missing or the device is
not sm_100, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libullava_sm100.so")

ABI_VERSION = 4   # ULLAVA_ABI_VERSION of include/ullava_sm100.h (struct layouts and signatures mirrored below)
BF16, F16, F32 = 0, 1, 2
EPI_NONE, EPI_RELU, EPI_GELU, EPI_QUICK_GELU, EPI_SILU_MUL = 0, 1, 2, 3, 4
SAM_N_WEIGHTS = 121

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [("A", _vp), ("lda", _i64), ("B", _vp), ("ldb", _i64), ("D", _vp), ("ldd", _i64),
                ("bias", _vp), ("residual", _vp), ("ldr", _i64),
                ("M", _i32), ("N", _i32), ("K", _i32), ("dtype", _i32), ("out_f32", _i32),
                ("epilogue", _i32), ("force_bn", _i32), ("force_splits", _i32), ("no_swap", _i32)]


class AttnArgs(C.Structure):
    _fields_ = [("q", _vp), ("q_bs", _i64), ("q_rs", _i64), ("q_hs", _i64),
                ("k", _vp), ("k_bs", _i64), ("k_rs", _i64), ("k_hs", _i64),
                ("v", _vp), ("v_bs", _i64), ("v_rs", _i64), ("v_hs", _i64),
                ("o", _vp), ("o_bs", _i64), ("o_rs", _i64), ("o_hs", _i64),
                ("batch", _i32), ("heads", _i32), ("seq_q", _i32), ("seq_k", _i32), ("head_dim", _i32),
                ("causal", _i32), ("q_pos0", _i32), ("scale", _f32), ("dtype", _i32)]


class SamDecoderArgs(C.Structure):
    _fields_ = [("weights", C.POINTER(_vp)), ("n_weights", _i32),
                ("image_embeddings", _vp), ("prompt_image", _vp), ("image_pe", _vp), ("text_embeds", _vp),
                ("n_prompts", _i32), ("low_res_masks", _vp), ("iou_pred", _vp),
                ("scratch", _vp), ("scratch_bytes", _sz), ("dtype", _i32)]


class VitArgs(C.Structure):
    _fields_ = [("weights", C.POINTER(_vp)), ("n_weights", _i32), ("pixels", _vp), ("out", _vp),
                ("scratch", _vp), ("scratch_bytes", _sz),
                ("batch", _i32), ("img", _i32), ("patch", _i32), ("hidden", _i32), ("heads", _i32),
                ("ffn", _i32), ("layers_used", _i32), ("k_pad", _i32), ("act", _i32), ("eps", _f32),
                ("dtype", _i32)]


class SamEncoderArgs(C.Structure):
    _fields_ = [("weights", C.POINTER(_vp)), ("n_weights", _i32), ("pixels", _vp), ("out", _vp),
                ("scratch", _vp), ("scratch_bytes", _sz), ("win_rows", _vp), ("unwin_rows", _vp),
                ("batch", _i32), ("img", _i32), ("patch", _i32), ("embed_dim", _i32), ("depth", _i32),
                ("heads", _i32), ("window", _i32), ("out_chans", _i32), ("global_mask", C.c_uint64),
                ("eps", _f32), ("dtype", _i32), ("block_begin", _i32), ("block_end", _i32)]


class LlamaArgs(C.Structure):
    _fields_ = [("weights", C.POINTER(_vp)), ("n_weights", _i32), ("hidden", _vp), ("final_out", _vp),
                ("all_hidden", _vp), ("k_cache", _vp), ("v_cache", _vp), ("scratch", _vp), ("scratch_bytes", _sz),
                ("batch", _i32), ("seq", _i32), ("pos0", _i32), ("max_seq", _i32),
                ("layers", _i32), ("hidden_size", _i32), ("heads", _i32), ("head_dim", _i32), ("ffn", _i32),
                ("eps", _f32), ("rope_cos", _vp), ("rope_sin", _vp), ("dtype", _i32),
                ("chain_program", _vp), ("chain_bytes", _sz), ("pos_offset", _vp)]


class DecodeArgs(C.Structure):
    _fields_ = [("llama", LlamaArgs), ("pos_dev", _vp), ("embed_table", _vp), ("vocab", _i32), ("lm_head", _vp),
                ("cur_ids", _vp), ("logits", _vp), ("seqs", _vp), ("seqs_ld", _i64), ("hid_buf", _vp),
                ("hid_bs", _i64), ("finished", _vp), ("eos_id", _i32), ("pad_id", _i32),
                ("uniforms", _vp), ("uniforms_ld", _i64), ("temperature", _f32), ("top_p", _f32), ("top_k", _i32)]


# name -> (restype, argtypes); must list every symbol declared in include/ullava_sm100.h
_SIGNATURES = {
    "ullava_abi_version": (_i32, []),
    "ullava_last_error": (C.c_char_p, []),
    "ullava_create": (_i32, [_i32, C.POINTER(_vp)]),
    "ullava_destroy": (_i32, [_vp]),
    "ullava_set_workspace": (_i32, [_vp, _vp, _sz]),
    "ullava_launch_count": (_i64, [_vp]),
    "ullava_partition_create": (_i32, [_i32, _i32, _i32, _i32, C.POINTER(_vp)]),
    "ullava_partition_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_vp), C.POINTER(_vp)]),
    "ullava_partition_destroy": (_i32, [_vp]),
    "ullava_set_sm_limit": (_i32, [_vp, _i32]),
    "ullava_sm_count": (_i32, [_vp]),
    "ullava_profile_begin": (_i32, [_vp]),
    "ullava_profile_end": (_i32, [_vp, C.POINTER(C.c_double)]),
    "ullava_gemm": (_i32, [_vp, C.POINTER(GemmArgs), _vp]),
    "ullava_layernorm": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _i32, _i32, _vp]),
    "ullava_rmsnorm": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _f32, _i32, _vp]),
    "ullava_attention": (_i32, [_vp, C.POINTER(AttnArgs), _vp]),
    "ullava_attention_relpos": (_i32, [_vp, C.POINTER(AttnArgs), _vp, _vp, _i32, _vp, _vp]),
    "ullava_set_attention_impl": (_i32, [_vp, _i32]),
    "ullava_set_pdl": (_i32, [_vp, _i32]),
    "ullava_gemm_next_weight": (_i32, [_vp, _vp, _i32, _i32, _i64]),
    "ullava_set_weight_prefetch": (_i32, [_vp, _i32]),
    "ullava_attention_decode": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _i32, _i32, _i32,
                                       _f32, _i32, _vp]),
    "ullava_attention_decode_rope": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp,
                                            _i32, _vp, _vp, _vp, _f32, _i32, _vp]),
    "ullava_rope_kvcache": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp,
                                   _i32, _vp]),
    "ullava_vit_im2col": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ullava_vit_assemble": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ullava_embed_gather": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ullava_splice_rows": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ullava_copy_rows": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp]),
    "ullava_argmax": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp]),
    "ullava_video_pool": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ullava_sam_mask_decoder": (_i32, [_vp, C.POINTER(SamDecoderArgs), _vp]),
    "ullava_sam_mask_decoder_scratch_bytes": (_sz, [_i32]),
    "ullava_sam_postprocess": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                      _vp]),
    "ullava_vit_forward": (_i32, [_vp, C.POINTER(VitArgs), _vp]),
    "ullava_vit_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "ullava_sam_encoder_forward": (_i32, [_vp, C.POINTER(SamEncoderArgs), _vp]),
    "ullava_sam_encoder_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "ullava_llama_forward": (_i32, [_vp, C.POINTER(LlamaArgs), _vp]),
    "ullava_llama_scratch_bytes": (_sz, [_i32, _i32, _i32]),
    "ullava_llama_decode_step": (_i32, [_vp, C.POINTER(DecodeArgs), _vp]),
    "ullava_debug_chain_trace": (_i32, [_vp, _vp]),
    "ullava_debug_fmha_trace": (_i32, [_vp, _vp]),
    "ullava_llama_chain_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ullava_llama_chain_prepare": (_i32, [_vp, C.POINTER(DecodeArgs)]),
    "ullava_greedy_step": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _i32, _i32,
                                  _vp, _vp]),
    "ullava_sample_step": (_i32, [_vp, _vp, _i64, _i32, _i32, _f32, _f32, _i32, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64,
                                  _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "ullava_mask_iou_counts": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i64, _i32, _vp, _vp]),
    "ullava_seg_meter_update": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "ullava_box_iou_diag": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "ullava_cross_entropy_scratch_bytes": (_sz, [_i32, _i32]),
    "ullava_cross_entropy": (_i32, [_vp, _vp, _i32, _i32, _i64, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _sz,
                                    _vp]),
    "ullava_resample_coeffs": (_i32, [_i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32), _sz]),
    "ullava_resize_u8_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ullava_resize_u8": (_i32, [_vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _sz, _vp]),
    "ullava_clip_preprocess": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(_f32), C.POINTER(_f32),
                                      C.c_double, _vp, _i32, _vp]),
    "ullava_sam_preprocess": (_i32, [_vp, _vp, _i32, _i32, _i32, C.POINTER(_f32), C.POINTER(_f32), _vp, _i32, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def build_library(force: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libullava_sm100.so (nvcc, sm_100a only)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["bash", script], check=True, capture_output=True)
    return LIB_PATH


def load_library():
    """dlopen the C ABI; raises (never falls back) when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / torch fallback for the u-LLaVA sm_100a path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.ullava_abi_version() != ABI_VERSION:
            raise RuntimeError("libullava_sm100.so ABI version mismatch")
        _lib = lib
        return lib


def exported_symbols() -> Sequence[str]:
    return tuple(_SIGNATURES.keys())


def resample_coeffs(in_size: int, out_size: int, bicubic: bool):
    """Host-only: (bounds [out, 2], taps [out, ksize]) of one Pillow resampling pass as the device kernels use them."""
    import numpy as np
    lib = load_library()
    ksize = lib.ullava_resample_coeffs(in_size, out_size, int(bicubic), None, None, 0)
    if ksize <= 0:
        raise ValueError("bad resampling geometry")
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    taps = np.zeros((out_size, ksize), dtype=np.int32)
    lib.ullava_resample_coeffs(in_size, out_size, int(bicubic), bounds.ctypes.data_as(C.POINTER(_i32)),
                               taps.ctypes.data_as(C.POINTER(_i32)), taps.size)
    return bounds, taps


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    raise TypeError(f"the sm_100a path stores activations/weights in bf16 or fp16, got {dt}")


def px_dtype_code(dt: torch.dtype) -> int:
    """pixel tensors of the preprocessing kernels: fp32 (the reference processors' output) or the model dtype"""
    return F32 if dt == torch.float32 else dtype_code(dt)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class NativeError(RuntimeError):
    pass


class Context:
    """Handle of ullava_create: one per (device, lane).  Owns a torch-allocated scratch workspace.  A context serves
    one stream at a time (workspace, stream-K counters and the prefetch hint are per context); lane 0 is the default,
    Partition (below) creates one more context per concurrent lane."""

    _by_device = {}

    def __init__(self, device: int, workspace_bytes: int = 256 << 20):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("CUDA device required: the u-LLaVA B200 path has no CPU fallback")
        h = _vp()
        st = self.lib.ullava_create(int(device), C.byref(h))
        if st != 0:
            raise NativeError(f"ullava_create failed ({st}): {self.lib.ullava_last_error().decode()}")
        self.handle = h
        self.device = torch.device("cuda", device)
        self.workspace = torch.empty(workspace_bytes, dtype=torch.uint8, device=self.device)
        self._chk(self.lib.ullava_set_workspace(self.handle, self.workspace.data_ptr(), workspace_bytes))

    @classmethod
    def get(cls, device=None, lane: int = 0) -> "Context":
        if device is None:
            device = torch.cuda.current_device()
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        ctx = cls._by_device.get((device, lane))
        if ctx is None:
            ctx = cls(device)
            cls._by_device[(device, lane)] = ctx
        return ctx

    def set_sm_limit(self, sms: int):
        """Size persistent grids / stream-K splits for `sms` SMs (0 = the whole device): the SM count of the partition
        lane this context launches into."""
        self._chk(self.lib.ullava_set_sm_limit(self.handle, int(sms)))

    def sm_count(self) -> int:
        return int(self.lib.ullava_sm_count(self.handle))

    def _chk(self, st: int):
        if st != 0:
            raise NativeError(f"libullava_sm100 error {st}: {self.lib.ullava_last_error().decode()}")

    @classmethod
    def total_launches(cls, device=None) -> int:
        """Kernels enqueued by every lane's context of `device`."""
        if isinstance(device, torch.device):
            device = device.index
        return sum(c.launch_count() for (d, _), c in cls._by_device.items() if device is None or d == device)

    def launch_count(self) -> int:
        return int(self.lib.ullava_launch_count(self.handle))

    # ---- per-kernel-class CUDA-event profile (bench.py roofline) ---------------------------
    PROF_CLASSES = ("gemm_tensor", "gemm_stream", "attn_prefill", "attn_decode", "norm", "glue", "sam", "other")

    def profile_begin(self):
        self._chk(self.lib.ullava_profile_begin(self.handle))

    def profile_end(self) -> dict:
        buf = (C.c_double * (len(self.PROF_CLASSES) * 4))()
        self._chk(self.lib.ullava_profile_end(self.handle, buf))
        return {name: dict(ms=buf[4 * i], flops=buf[4 * i + 1], bytes=buf[4 * i + 2], launches=int(buf[4 * i + 3]))
                for i, name in enumerate(self.PROF_CLASSES)}

    # ---- primitives ------------------------------------------------------------------
    def gemm(self, a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, epilogue=EPI_NONE, out=None,
             out_f32=False, force_bn=0, force_splits=0, no_swap=False) -> torch.Tensor:
        """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) + residual   (nn.Linear semantics)."""
        assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
        assert a.stride(1) == 1 and w.stride(1) == 1
        M, K = a.shape
        N = w.shape[0]
        n_out = N // 2 if epilogue == EPI_SILU_MUL else N
        if out is None:
            out = torch.empty((M, n_out), dtype=torch.float32 if out_f32 else a.dtype, device=a.device)
        assert out.stride(1) == 1
        g = GemmArgs()
        g.A, g.lda, g.B, g.ldb = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
        g.D, g.ldd = out.data_ptr(), out.stride(0)
        g.bias = _ptr(bias)
        g.residual = _ptr(residual)
        g.ldr = residual.stride(0) if residual is not None else 0
        g.M, g.N, g.K = M, N, K
        g.dtype = dtype_code(a.dtype)
        g.out_f32 = 1 if out.dtype == torch.float32 else 0
        g.epilogue, g.force_bn, g.force_splits, g.no_swap = epilogue, force_bn, force_splits, int(no_swap)
        self._chk(self.lib.ullava_gemm(self.handle, C.byref(g), _stream()))
        return out

    def layernorm(self, x, weight, bias, eps, act=EPI_NONE, out=None):
        x2 = x.reshape(-1, x.shape[-1])
        if out is None:
            out = torch.empty_like(x2)
        o2 = out.reshape(-1, x.shape[-1])
        self._chk(self.lib.ullava_layernorm(self.handle, x2.data_ptr(), x2.stride(0), weight.data_ptr(),
                                            bias.data_ptr(), o2.data_ptr(), o2.stride(0), x2.shape[0], x2.shape[1],
                                            float(eps), act, dtype_code(x.dtype), _stream()))
        return out.reshape(x.shape)

    def rmsnorm(self, x, weight, eps, out=None):
        x2 = x.reshape(-1, x.shape[-1])
        if out is None:
            out = torch.empty_like(x2)
        o2 = out.reshape(-1, x.shape[-1])
        self._chk(self.lib.ullava_rmsnorm(self.handle, x2.data_ptr(), x2.stride(0), weight.data_ptr(), o2.data_ptr(),
                                          o2.stride(0), x2.shape[0], x2.shape[1], float(eps), dtype_code(x.dtype),
                                          _stream()))
        return out.reshape(x.shape)

    def attention(self, q, k, v, causal=False, q_pos0=0, scale=None, out=None):
        """q,k,v: [B, S, H, D] views (last dim contiguous); returns [B, Sq, H, D]."""
        B, Sq, H, D = q.shape
        Sk = k.shape[1]
        if out is None:
            out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
        a = AttnArgs()
        for name, t in (("q", q), ("k", k), ("v", v), ("o", out)):
            assert t.stride(3) == 1
            setattr(a, name, t.data_ptr())
            setattr(a, name + "_bs", t.stride(0))
            setattr(a, name + "_rs", t.stride(1))
            setattr(a, name + "_hs", t.stride(2))
        a.batch, a.heads, a.seq_q, a.seq_k, a.head_dim = B, H, Sq, Sk, D
        a.causal, a.q_pos0 = int(causal), int(q_pos0)
        a.scale = float(scale if scale is not None else D ** -0.5)
        a.dtype = dtype_code(q.dtype)
        self._chk(self.lib.ullava_attention(self.handle, C.byref(a), _stream()))
        return out

    def set_pdl(self, enabled: bool):
        """Programmatic dependent launch for the decode (M <= 32) GEMMs: weight prefetch under the previous kernel."""
        self._chk(self.lib.ullava_set_pdl(self.handle, int(bool(enabled))))

    def gemm_next_weight(self, w: Optional[torch.Tensor]):
        """Hint: the next M <= 32 gemm() pulls the head of `w` ([N, K] nn.Linear weight of the GEMM after it) into L2."""
        if w is None:
            self._chk(self.lib.ullava_gemm_next_weight(self.handle, None, 0, 0, 0))
        else:
            assert w.dim() == 2 and w.stride(1) == 1
            self._chk(self.lib.ullava_gemm_next_weight(self.handle, w.data_ptr(), w.shape[0], w.shape[1], w.stride(0)))

    def set_weight_prefetch(self, tiles_per_sm: int):
        """16 KB weight tiles per SM each decode GEMM pulls ahead for its successor (0 = off)."""
        self._chk(self.lib.ullava_set_weight_prefetch(self.handle, int(tiles_per_sm)))

    def set_attention_impl(self, impl: int):
        """0 = per shape (default), 1 = warp-level mma.sync kernels only, 2 = tcgen05/TMEM wherever compiled."""
        self._chk(self.lib.ullava_set_attention_impl(self.handle, int(impl)))

    def attention_relpos(self, q, k, v, rel_h, rel_w, grid_side, scale=None, out=None, o_row_map=None):
        """SAM ViT attention with decomposed rel-pos bias.  q,k,v: [B, S*S, H, D] views; rel_h/rel_w: [2S-1, D]."""
        B, Sq, H, D = q.shape
        if out is None:
            out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
        a = AttnArgs()
        for name, t in (("q", q), ("k", k), ("v", v), ("o", out)):
            assert t.stride(3) == 1
            setattr(a, name, t.data_ptr())
            setattr(a, name + "_bs", t.stride(0))
            setattr(a, name + "_rs", t.stride(1))
            setattr(a, name + "_hs", t.stride(2))
        a.batch, a.heads, a.seq_q, a.seq_k, a.head_dim = B, H, Sq, k.shape[1], D
        a.causal, a.q_pos0 = 0, 0
        a.scale = float(scale if scale is not None else D ** -0.5)
        a.dtype = dtype_code(q.dtype)
        assert rel_h.is_contiguous() and rel_w.is_contiguous()
        self._chk(self.lib.ullava_attention_relpos(self.handle, C.byref(a), rel_h.data_ptr(), rel_w.data_ptr(),
                                                   int(grid_side), _ptr(o_row_map), _stream()))
        return out

    def attention_decode(self, q, k_cache, v_cache, ctx_len, scale=None, out=None):
        """q: [B, H*D] (row stride arbitrary); caches: [B, H, max_seq, D]; returns [B, H*D]."""
        B, H, _, D = k_cache.shape
        if out is None:
            out = torch.empty((B, H * D), dtype=q.dtype, device=q.device)
        self._chk(self.lib.ullava_attention_decode(
            self.handle, q.data_ptr(), q.stride(0), k_cache.data_ptr(), v_cache.data_ptr(), k_cache.stride(0),
            k_cache.stride(1), out.data_ptr(), out.stride(0), B, H, D, int(ctx_len),
            float(scale if scale is not None else D ** -0.5), dtype_code(q.dtype), _stream()))
        return out

    def attention_decode_rope(self, qkv, k_cache, v_cache, ctx_len, cos, sin, pos_dev=None, scale=None, out=None,
                              pos_offset=None):
        """Fused decode step: RoPE(q, k) of the new token, KV-cache write at row ctx_len - 1, single-query attention.
        qkv: [B, 3 * H * D] packed; caches [B, H, max_seq, D]; cos / sin fp32 [max_seq, D / 2]."""
        B, H, max_seq, D = k_cache.shape
        if out is None:
            out = torch.empty((B, H * D), dtype=qkv.dtype, device=qkv.device)
        self._chk(self.lib.ullava_attention_decode_rope(
            self.handle, qkv.data_ptr(), qkv.stride(0), k_cache.data_ptr(), v_cache.data_ptr(), k_cache.stride(0),
            k_cache.stride(1), out.data_ptr(), out.stride(0), B, H, D, int(ctx_len), _ptr(pos_dev), max_seq,
            cos.data_ptr(), sin.data_ptr(), _ptr(pos_offset), float(scale if scale is not None else D ** -0.5),
            dtype_code(qkv.dtype), _stream()))
        return out

    def rope_kvcache(self, qkv, k_cache, v_cache, batch, seq, pos0, cos, sin):
        B, H, _, D = k_cache.shape
        self._chk(self.lib.ullava_rope_kvcache(
            self.handle, qkv.data_ptr(), qkv.stride(0), k_cache.data_ptr(), v_cache.data_ptr(), k_cache.stride(0),
            k_cache.stride(1), batch, seq, H, D, int(pos0), cos.data_ptr(), sin.data_ptr(), dtype_code(qkv.dtype),
            _stream()))

    def vit_im2col(self, pixels, patch, k_pad):
        B, _, img, _ = pixels.shape
        g = img // patch
        out = torch.empty((B * g * g, k_pad), dtype=pixels.dtype, device=pixels.device)
        self._chk(self.lib.ullava_vit_im2col(self.handle, pixels.data_ptr(), out.data_ptr(), B, img, patch, k_pad,
                                             dtype_code(pixels.dtype), _stream()))
        return out

    def vit_assemble(self, patch_embeds, cls, pos, batch):
        n_p = patch_embeds.shape[0] // batch
        dim = patch_embeds.shape[1]
        out = torch.empty((batch, n_p + 1, dim), dtype=patch_embeds.dtype, device=patch_embeds.device)
        self._chk(self.lib.ullava_vit_assemble(self.handle, patch_embeds.data_ptr(), cls.data_ptr(), pos.data_ptr(),
                                               out.data_ptr(), batch, n_p, dim, dtype_code(out.dtype), _stream()))
        return out

    def embed_gather(self, ids, table, out=None):
        rows = ids.numel()
        dim = table.shape[1]
        if out is None:
            out = torch.empty((rows, dim), dtype=table.dtype, device=table.device)
        ids = ids.reshape(-1).contiguous()
        assert ids.dtype == torch.int64
        self._chk(self.lib.ullava_embed_gather(self.handle, ids.data_ptr(), table.data_ptr(), out.data_ptr(), rows,
                                               dim, table.shape[0], dtype_code(table.dtype), _stream()))
        return out

    def splice_rows(self, embeds, feats, start):
        B, L, dim = embeds.shape
        n_patch = feats.shape[1]
        assert start.dtype == torch.int32 and start.is_cuda
        self._chk(self.lib.ullava_splice_rows(self.handle, embeds.data_ptr(), feats.data_ptr(), start.data_ptr(), B, L,
                                              n_patch, dim, dtype_code(embeds.dtype), _stream()))

    def copy_rows(self, src, dst, batch, rows, cols, src_bs, src_rs, dst_bs, dst_rs):
        self._chk(self.lib.ullava_copy_rows(self.handle, src.data_ptr(), src_bs, src_rs, dst.data_ptr(), dst_bs,
                                            dst_rs, batch, rows, cols, dtype_code(src.dtype), _stream()))

    def argmax(self, logits, out=None):
        rows, cols = logits.shape
        assert logits.dtype == torch.float32 and logits.stride(1) == 1
        if out is None:
            out = torch.empty((rows,), dtype=torch.int64, device=logits.device)
        self._chk(self.lib.ullava_argmax(self.handle, logits.data_ptr(), logits.stride(0), out.data_ptr(), rows, cols,
                                         _stream()))
        return out

    # ---- stage-level --------------------------------------------------------------------
    @staticmethod
    def pointer_table(tensors: Sequence[torch.Tensor]):
        arr = (_vp * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = t.data_ptr()
        return arr

    def sam_mask_decoder(self, weight_table, n_weights, image_embeddings, prompt_image, image_pe, text_embeds,
                         want_iou=True):
        n = text_embeds.shape[0]
        dt = text_embeds.dtype
        dev = text_embeds.device
        masks = torch.empty((n, 4, 256, 256), dtype=dt, device=dev)
        iou = torch.empty((n, 4), dtype=dt, device=dev) if want_iou else None
        sb = int(self.lib.ullava_sam_mask_decoder_scratch_bytes(n))
        scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
        a = SamDecoderArgs()
        a.weights, a.n_weights = weight_table, n_weights
        a.image_embeddings, a.prompt_image = image_embeddings.data_ptr(), prompt_image.data_ptr()
        a.image_pe, a.text_embeds, a.n_prompts = image_pe.data_ptr(), text_embeds.data_ptr(), n
        a.low_res_masks, a.iou_pred = masks.data_ptr(), _ptr(iou)
        a.scratch, a.scratch_bytes, a.dtype = scratch.data_ptr(), sb, dtype_code(dt)
        self._chk(self.lib.ullava_sam_mask_decoder(self.handle, C.byref(a), _stream()))
        return masks, iou

    def sam_postprocess(self, masks, mask_stride, n, low_res, img_size, input_size, original_size, pack_bits=False):
        out_h, out_w = int(original_size[0]), int(original_size[1])
        out = torch.empty((n, out_h, out_w), dtype=torch.float32, device=masks.device)
        bits = None
        if pack_bits:
            bits = torch.zeros((n, (out_h * out_w + 31) // 32), dtype=torch.int32, device=masks.device)
        self._chk(self.lib.ullava_sam_postprocess(self.handle, masks.data_ptr(), int(mask_stride), out.data_ptr(),
                                                  _ptr(bits), n, low_res, img_size, int(input_size[0]),
                                                  int(input_size[1]), out_h, out_w, dtype_code(masks.dtype),
                                                  _stream()))
        return out, bits

    def vit_forward(self, weight_table, n_weights, pixels, cfg: dict, scratch=None):
        B = pixels.shape[0]
        g = cfg["img"] // cfg["patch"]
        out = torch.empty((B, g * g, cfg["hidden"]), dtype=pixels.dtype, device=pixels.device)
        sb = int(self.lib.ullava_vit_scratch_bytes(B, cfg["img"], cfg["patch"], cfg["hidden"], cfg["ffn"],
                                                   cfg["k_pad"]))
        if scratch is None or scratch.numel() < sb:
            scratch = torch.empty(sb, dtype=torch.uint8, device=pixels.device)
        a = VitArgs()
        a.weights, a.n_weights, a.pixels, a.out = weight_table, n_weights, pixels.data_ptr(), out.data_ptr()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        a.batch, a.img, a.patch, a.hidden, a.heads = B, cfg["img"], cfg["patch"], cfg["hidden"], cfg["heads"]
        a.ffn, a.layers_used, a.k_pad, a.act, a.eps = cfg["ffn"], cfg["layers_used"], cfg["k_pad"], cfg["act"], cfg["eps"]
        a.dtype = dtype_code(pixels.dtype)
        self._chk(self.lib.ullava_vit_forward(self.handle, C.byref(a), _stream()))
        return out

    def sam_encoder_forward(self, weight_table, n_weights, pixels, cfg: dict, win_rows, unwin_rows, scratch=None,
                            blocks=None):
        """SAM ViT image encoder: pixels [B,3,img,img] -> [B,out_chans,g,g] (NCHW).  blocks = (begin, end): only that
        block range (patch embedding with begin == 0, neck + output with end == depth; the token state stays in
        `scratch` between the calls); the output is returned (and allocated) by the call that ends at depth."""
        B = pixels.shape[0]
        g = cfg["img"] // cfg["patch"]
        b0, b1 = blocks if blocks is not None else (0, cfg["depth"])
        last = b1 >= cfg["depth"]
        out = torch.empty((B, cfg["out_chans"], g, g), dtype=pixels.dtype, device=pixels.device) if last else None
        sb = int(self.lib.ullava_sam_encoder_scratch_bytes(B, cfg["img"], cfg["patch"], cfg["embed_dim"],
                                                           cfg["window"], cfg["out_chans"]))
        if scratch is None or scratch.numel() < sb:
            scratch = torch.empty(sb, dtype=torch.uint8, device=pixels.device)
        a = SamEncoderArgs()
        a.weights, a.n_weights, a.pixels = weight_table, n_weights, pixels.data_ptr()
        a.out = out.data_ptr() if last else scratch.data_ptr()   # not written unless the range ends at depth
        a.block_begin, a.block_end = int(b0), int(b1)
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        a.win_rows, a.unwin_rows = _ptr(win_rows), _ptr(unwin_rows)
        a.batch, a.img, a.patch, a.embed_dim = B, cfg["img"], cfg["patch"], cfg["embed_dim"]
        a.depth, a.heads, a.window, a.out_chans = cfg["depth"], cfg["heads"], cfg["window"], cfg["out_chans"]
        a.global_mask, a.eps, a.dtype = cfg["global_mask"], cfg["eps"], dtype_code(pixels.dtype)
        self._chk(self.lib.ullava_sam_encoder_forward(self.handle, C.byref(a), _stream()))
        return out, scratch

    def llama_chain_bytes(self, layers: int, hidden: int, ffn: int, vocab: int) -> int:
        return int(self.lib.ullava_llama_chain_bytes(int(layers), int(hidden), int(ffn), int(vocab)))

    def llama_chain_prepare(self, args: "DecodeArgs"):
        """Builds the decode-layer chain program into args.llama.chain_program (synchronous; not under capture)."""
        self._chk(self.lib.ullava_llama_chain_prepare(self.handle, C.byref(args)))

    def llama_decode_step(self, args: "DecodeArgs"):
        """One greedy decode step with the position in device memory (CUDA-graph replayable)."""
        self._chk(self.lib.ullava_llama_decode_step(self.handle, C.byref(args), _stream()))

    def greedy_step(self, logits, cur_ids, seqs, final_h, hid_buf, finished, eos_id, pad_id, pos_dev):
        """next = argmax(logits) with eos/pad handling; cur_ids <- next; seqs[:, pos+1] <- next;
        hid_buf[:, pos] <- final_h; ++pos (pos read from / written to device memory)."""
        rows, cols = logits.shape
        self._chk(self.lib.ullava_greedy_step(
            self.handle, logits.data_ptr(), logits.stride(0), rows, cols, cur_ids.data_ptr(), _ptr(seqs),
            seqs.stride(0) if seqs is not None else 0, _ptr(final_h), _ptr(hid_buf),
            hid_buf.stride(0) if hid_buf is not None else 0, final_h.shape[-1] if final_h is not None else 8,
            _ptr(finished), int(eos_id), int(pad_id), pos_dev.data_ptr(), _stream()))

    def video_pool(self, feats: torch.Tensor) -> torch.Tensor:
        """feats [bs, T, N, D] -> [bs, T + N, D]: temporal means then spatial means (encode_video)."""
        bs, t, n, d = feats.shape
        feats = feats.contiguous()
        out = torch.empty((bs, t + n, d), dtype=feats.dtype, device=feats.device)
        self._chk(self.lib.ullava_video_pool(self.handle, feats.data_ptr(), out.data_ptr(), bs, t, n, d,
                                             dtype_code(feats.dtype), _stream()))
        return out

    def sample_step(self, logits, temperature, top_p, uniforms, cur_ids, seqs=None, final_h=None, hid_buf=None,
                    finished=None, eos_id=-1, pad_id=0, pos_dev=None, probs_out=None, top_k=0):
        """Temperature / top-k / top-p draw per row by inverse CDF with the caller's uniforms[pos, b]; bookkeeping as
        greedy_step.  probs_out (optional [rows, cols] fp32) receives the filtered, renormalised distribution.
        top_k = 0 / top_p = None switch the filters off (HF's default top_k = 50 is applied by generate())."""
        rows, cols = logits.shape
        assert uniforms.dtype == torch.float32 and uniforms.dim() == 2 and uniforms.stride(1) == 1
        self._chk(self.lib.ullava_sample_step(
            self.handle, logits.data_ptr(), logits.stride(0), rows, cols, float(temperature),
            float(top_p if top_p is not None else 1.0), int(top_k or 0), uniforms.data_ptr(), uniforms.stride(0),
            cur_ids.data_ptr(),
            _ptr(seqs), seqs.stride(0) if seqs is not None else 0, _ptr(final_h), _ptr(hid_buf),
            hid_buf.stride(0) if hid_buf is not None else 0, final_h.shape[-1] if final_h is not None else 8,
            _ptr(finished), int(eos_id), int(pad_id), _ptr(pos_dev), _ptr(probs_out), _stream()))

    # ---- evaluation metrics (evaluation/tools.py, evaluation/eval_ullava.py) -------------------
    _KIND = {torch.float32: 0, torch.int32: 1, torch.uint8: 2}

    def mask_iou_counts(self, pred: torch.Tensor, target: torch.Tensor, ignore_index: int = 255, out=None):
        """pred [n, ...] fp32 logits (label = logit > 0) or int32 / uint8 labels; target [n, ...] int32 / uint8.
        Returns int32 [n, 6] = area_intersection[0:2], area_union[0:2], area_target[0:2] (K = 2)."""
        assert pred.shape == target.shape and pred.is_contiguous() and target.is_contiguous()
        assert target.dtype in (torch.int32, torch.uint8), target.dtype
        n = pred.shape[0]
        hw = pred[0].numel() if n else 1
        if out is None:
            out = torch.empty((n, 6), dtype=torch.int32, device=pred.device)
        self._chk(self.lib.ullava_mask_iou_counts(self.handle, pred.data_ptr(), self._KIND[pred.dtype],
                                                  target.data_ptr(), self._KIND[target.dtype], n, hw,
                                                  int(ignore_index), out.data_ptr(), _stream()))
        return out

    def seg_meter_update(self, counts: torch.Tensor, offsets: torch.Tensor, state: torch.Tensor):
        """state (8 fp64 on the device) += validate()'s per-image sums; offsets int32 [n_images + 1]."""
        assert counts.dtype == torch.int32 and offsets.dtype == torch.int32 and state.dtype == torch.float64
        assert state.numel() >= 8
        self._chk(self.lib.ullava_seg_meter_update(self.handle, counts.data_ptr(), offsets.data_ptr(),
                                                   offsets.numel() - 1, state.data_ptr(), _stream()))

    def box_iou_diag(self, pred: torch.Tensor, gt: torch.Tensor, meter: Optional[torch.Tensor] = None):
        """iou[i] = torchvision box_iou(pred[i] * 1000, gt[i] * 1000) (xyxy); meter (3 fp64): hits(>0.5), boxes."""
        assert pred.shape == gt.shape and pred.shape[-1] == 4 and pred.dtype == gt.dtype
        pred, gt = pred.contiguous(), gt.contiguous()
        n = pred.numel() // 4
        dt = F32 if pred.dtype == torch.float32 else dtype_code(pred.dtype)
        iou = torch.empty((n,), dtype=torch.float32, device=pred.device)
        self._chk(self.lib.ullava_box_iou_diag(self.handle, pred.data_ptr(), gt.data_ptr(), n, dt, iou.data_ptr(),
                                               _ptr(meter), _stream()))
        return iou

    def cross_entropy(self, logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
        """Shifted token cross-entropy (CrossEntropyLoss over logits[:, :-1] vs labels[:, 1:], mean over valid labels).
        logits [B, T, V] fp32 / bf16 / fp16 with unit stride on V; labels int64 [B, T].  Returns a 0-dim fp32 tensor."""
        B, T, V = logits.shape
        assert logits.stride(2) == 1 and labels.dtype == torch.int64 and labels.shape[0] == B and labels.shape[1] >= T
        labels = labels if labels.stride(1) == 1 else labels.contiguous()
        out = torch.empty((2,), dtype=torch.float32, device=logits.device)
        nbytes = int(self.lib.ullava_cross_entropy_scratch_bytes(B, T))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
        f32 = logits.dtype == torch.float32
        self._chk(self.lib.ullava_cross_entropy(self.handle, logits.data_ptr(), int(f32), 0 if f32 else dtype_code(logits.dtype),
                                                logits.stride(1), logits.stride(0), labels.data_ptr(), labels.stride(0),
                                                B, T, V, int(ignore_index), out.data_ptr(), scratch.data_ptr(), nbytes,
                                                _stream()))
        return out[0]

    # ---- image preprocessing (dataset/processors/clip_processor.py, dataset/tools/mask_toolbox.py) -----------
    def resize_u8(self, img: torch.Tensor, out_h: int, out_w: int, bicubic: bool) -> torch.Tensor:
        """PIL.Image.resize of a uint8 [H, W, 3] device image (BICUBIC or BILINEAR), bit-exact with Pillow."""
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        out = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=img.device)
        nbytes = int(self.lib.ullava_resize_u8_scratch_bytes(h, w, out_h, out_w))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=img.device)
        self._chk(self.lib.ullava_resize_u8(self.handle, img.data_ptr(), h, w, out.data_ptr(), out_h, out_w,
                                            1 if bicubic else 0, scratch.data_ptr(), nbytes, _stream()))
        return out

    def clip_preprocess(self, img: torch.Tensor, top: int, left: int, size: int, mean, std, dtype,
                        rescale: float = 1 / 255) -> torch.Tensor:
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_contiguous()
        out = torch.empty((3, size, size), dtype=dtype, device=img.device)
        m, s = (_f32 * 3)(*[float(v) for v in mean]), (_f32 * 3)(*[float(v) for v in std])
        self._chk(self.lib.ullava_clip_preprocess(self.handle, img.data_ptr(), img.shape[0], img.shape[1], top, left,
                                                  size, m, s, float(rescale), out.data_ptr(), px_dtype_code(dtype),
                                                  _stream()))
        return out

    def sam_preprocess(self, img: torch.Tensor, sam_size: int, mean, std, dtype) -> torch.Tensor:
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_contiguous()
        out = torch.empty((3, sam_size, sam_size), dtype=dtype, device=img.device)
        m, s = (_f32 * 3)(*[float(v) for v in mean]), (_f32 * 3)(*[float(v) for v in std])
        self._chk(self.lib.ullava_sam_preprocess(self.handle, img.data_ptr(), img.shape[0], img.shape[1], sam_size, m,
                                                 s, out.data_ptr(), px_dtype_code(dtype), _stream()))
        return out

    def fill_llama_args(self, a: "LlamaArgs", weight_table, n_weights, hidden, k_cache, v_cache, scratch, batch, seq,
                        pos0, cfg: dict, rope_cos, rope_sin, final_out=None, all_hidden=None, pos_offset=None):
        a.weights, a.n_weights = weight_table, n_weights
        a.hidden, a.final_out, a.all_hidden = hidden.data_ptr(), _ptr(final_out), _ptr(all_hidden)
        a.k_cache, a.v_cache = k_cache.data_ptr(), v_cache.data_ptr()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        a.batch, a.seq, a.pos0, a.max_seq = batch, seq, pos0, k_cache.shape[3]
        a.layers, a.hidden_size, a.heads, a.head_dim, a.ffn = (cfg["layers"], cfg["hidden"], cfg["heads"],
                                                               cfg["head_dim"], cfg["ffn"])
        a.eps = cfg["eps"]
        a.rope_cos, a.rope_sin = rope_cos.data_ptr(), rope_sin.data_ptr()
        a.dtype = dtype_code(hidden.dtype)
        a.pos_offset = _ptr(pos_offset)
        return a

    def llama_forward(self, weight_table, n_weights, hidden, k_cache, v_cache, scratch, batch, seq, pos0, cfg: dict,
                      rope_cos, rope_sin, final_out=None, all_hidden=None, pos_offset=None):
        a = LlamaArgs()
        a.pos_offset = _ptr(pos_offset)
        a.weights, a.n_weights = weight_table, n_weights
        a.hidden, a.final_out, a.all_hidden = hidden.data_ptr(), _ptr(final_out), _ptr(all_hidden)
        a.k_cache, a.v_cache = k_cache.data_ptr(), v_cache.data_ptr()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        a.batch, a.seq, a.pos0, a.max_seq = batch, seq, pos0, k_cache.shape[3]
        a.layers, a.hidden_size, a.heads, a.head_dim, a.ffn = (cfg["layers"], cfg["hidden"], cfg["heads"],
                                                               cfg["head_dim"], cfg["ffn"])
        a.eps = cfg["eps"]
        a.rope_cos, a.rope_sin = rope_cos.data_ptr(), rope_sin.data_ptr()
        a.dtype = dtype_code(hidden.dtype)
        self._chk(self.lib.ullava_llama_forward(self.handle, C.byref(a), _stream()))

    def llama_scratch_bytes(self, rows, hidden, ffn) -> int:
        return int(self.lib.ullava_llama_scratch_bytes(rows, hidden, ffn))


class Partition:
    """ullava_partition: two torch streams on disjoint SM sets of one device plus one Context per lane, sized to the
    lane (lane A: `sms_a` SMs, high priority -- the latency-sensitive decode steps; lane B: the rest)."""

    _by_key = {}

    def __init__(self, device: int, sms_a: int):
        lib = load_library()
        h = _vp()
        st = lib.ullava_partition_create(int(device), int(sms_a), -1, 0, C.byref(h))
        if st != 0:
            raise NativeError(f"ullava_partition_create failed ({st}): {lib.ullava_last_error().decode()}")
        self.lib, self.handle, self.device = lib, h, device
        a, b, sa, sb = _i32(), _i32(), _vp(), _vp()
        lib.ullava_partition_info(h, C.byref(a), C.byref(b), C.byref(sa), C.byref(sb))
        self.sms = (a.value, b.value)
        dev = torch.device("cuda", device)
        self.streams = (torch.cuda.ExternalStream(sa.value, device=dev), torch.cuda.ExternalStream(sb.value, device=dev))
        self.ctx = (Context.get(device, lane=1), Context.get(device, lane=2))
        self.ctx[0].set_sm_limit(self.sms[0])
        self.ctx[1].set_sm_limit(self.sms[1])

    @classmethod
    def current(cls, device):
        """The partition already created on `device` (there is at most one), or None."""
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        for (d, _), q in cls._by_key.items():
            if d == device:
                return q
        return None

    @classmethod
    def get(cls, device, sms_a: int) -> "Partition":
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        key = (device, sms_a)
        p = cls._by_key.get(key)
        if p is None:
            for (d, _), q in cls._by_key.items():
                if d == device:
                    raise NativeError("one SM partition per device (its lanes own contexts 1 and 2)")
            p = cls(device, sms_a)
            cls._by_key[key] = p
        return p

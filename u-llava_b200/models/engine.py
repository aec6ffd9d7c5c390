"""Host-side glue between the reference-shaped nn.Module tree and libullava_sm100.so.

The nn.Modules (LlamaModel, CLIPVisionModel, nn.Linear ...) are *parameter containers* that keep
the reference's attribute tree / state_dict keys / from_pretrained format; their forward() is never
called.  This module packs their parameters once into the layouts the kernels want (fused QKV,
gate/up interleaved by 16 rows, zero-padded im2col patch weight), keeps pointer tables for the C ABI
and owns the caller-side buffers (KV cache, activations, scratch).

Everything here is host logic + pointer plumbing; all arithmetic happens in the C ABI calls.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

try:  # importable both as models.engine (drop-in layout) and standalone
    import native
except ImportError:  # pragma: no cover
    from .. import native  # type: ignore


def _sig(params, modules=()) -> Tuple:
    """Identity of everything a packed copy was made from: parameter storage / in-place version / dtype / device, plus
    -- for the linear layers whose weights are fused into packed copies -- the module objects and their LoRA state
    (merge_and_unload() swaps the modules, merge_adapter() / set_adapter() / disable_adapter() flip flags without
    touching a version counter)."""
    return (tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in params),
            tuple((id(m),) + _lora_state(m) for m in modules))


def _lora_state(lin) -> Tuple:
    if not hasattr(lin, "lora_A"):
        return ()
    active = getattr(lin, "active_adapters", None)
    if active is None or callable(active):
        active = getattr(lin, "active_adapter", None)
    if isinstance(active, str):
        active = (active,)
    return (bool(getattr(lin, "merged", False)), bool(getattr(lin, "disable_adapters", False)), tuple(active or ()))


def effective_weight(lin) -> torch.Tensor:
    """The [out, in] matrix a (possibly LoRA-wrapped) nn.Linear applies at inference time.

    The reference wraps `model.llm` in `PeftModel` when lora_r > 0 (inference_ullava.py:42-43, eval_ullava.py:137-138;
    targets q_proj / v_proj by default, train_ullava.py:42,88-113): the injected peft layers keep the base weight in
    `.weight` (peft 0.4, the reference's pin) or `.base_layer.weight` (peft >= 0.6) and add
    scaling * lora_B(lora_A(x)) in their forward.  The kernels read packed weight copies, so the adapters are folded
    in at pack time:  W_eff = W + sum_active scaling * B @ A  (exactly what peft's own merge() computes).  Anything this
    does not understand (DoRA, embedding adapters, adapters with a bias) raises instead of being dropped silently."""
    base = getattr(lin, "base_layer", lin)
    w = base.weight.detach()
    if not hasattr(lin, "lora_A"):
        return w
    merged, disabled, active = _lora_state(lin)
    if merged or disabled:
        return w            # merged: the delta already lives in W; disabled: the base model is what runs
    if len(getattr(lin, "lora_embedding_A", {}) or {}) or len(getattr(lin, "lora_magnitude_vector", {}) or {}):
        raise NotImplementedError("only plain LoRA adapters on nn.Linear are folded into the packed weights")
    wf = None
    for name in active:
        if name not in lin.lora_A:
            continue
        A, Bm = lin.lora_A[name], lin.lora_B[name]
        if getattr(A, "bias", None) is not None or getattr(Bm, "bias", None) is not None:
            raise NotImplementedError("LoRA adapters with a bias are not supported on the B200 path")
        r = lin.r[name] if isinstance(getattr(lin, "r", None), dict) else A.weight.shape[0]
        if r <= 0:
            continue
        delta = (Bm.weight.detach().float() @ A.weight.detach().float()) * float(lin.scaling[name])
        if getattr(lin, "fan_in_fan_out", False):
            delta = delta.t()
        wf = (w.float() if wf is None else wf) + delta.to(w.device)
    return w if wf is None else wf.to(w.dtype)


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[F,K],[F,K] -> [2F,K] where every 32-row chunk = 16 gate rows followed by the 16 matching up rows
    (layout consumed by the SILU_MUL GEMM epilogue)."""
    F_, K = gate.shape
    assert F_ % 16 == 0
    return torch.stack([gate.view(F_ // 16, 16, K), up.view(F_ // 16, 16, K)], dim=1).reshape(2 * F_, K).contiguous()


class KVCache:
    """KV cache owned by the caller side: k, v [layers, B, heads, max_seq, head_dim]; `length` tokens valid.
    Returned as `past_key_values` by UllavaCoreForCausalLM.forward(use_cache=True)."""

    def __init__(self, layers, batch, heads, max_seq, head_dim, dtype, device):
        self.k = torch.empty((layers, batch, heads, max_seq, head_dim), dtype=dtype, device=device)
        self.v = torch.empty_like(self.k)
        self.length = 0
        self.max_seq = max_seq
        self.batch = batch

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.length

    def __len__(self):
        return self.k.shape[0]

    def __bool__(self):
        return True

    def to_legacy(self):
        """Tuple of (k, v) [B, heads, length, hd] per layer, HF legacy format (views)."""
        return tuple((self.k[l, :, :, : self.length], self.v[l, :, :, : self.length]) for l in range(self.k.shape[0]))


class VisionTower:
    """CLIP ViT weights packed for ullava_vit_forward."""

    def __init__(self, vision_encoder, hidden_layer: int):
        self.module = vision_encoder
        self.hidden_layer = hidden_layer
        self._sig = None

    def _pack(self):
        vm = self.module.vision_model
        cfg = self.module.config
        n_layers = cfg.num_hidden_layers
        idx = self.hidden_layer if self.hidden_layer >= 0 else n_layers + 1 + self.hidden_layer
        if not (0 <= idx <= n_layers):
            raise ValueError(f"vision_hidden_layer {self.hidden_layer} out of range for {n_layers} layers")
        act = {"quick_gelu": native.EPI_QUICK_GELU, "gelu": native.EPI_GELU}.get(cfg.hidden_act)
        if act is None:
            raise NotImplementedError(f"CLIP hidden_act {cfg.hidden_act!r}")
        pw = vm.embeddings.patch_embedding.weight
        H = pw.shape[0]
        k = pw[0].numel()
        k_pad = (k + 63) // 64 * 64
        wp = torch.zeros((H, k_pad), dtype=pw.dtype, device=pw.device)
        wp[:, :k] = pw.detach().reshape(H, k)
        tensors = [wp, vm.embeddings.class_embedding.detach().contiguous(),
                   vm.embeddings.position_embedding.weight.detach().contiguous(),
                   vm.pre_layrnorm.weight.detach(), vm.pre_layrnorm.bias.detach()]
        for l in range(idx):
            lay = vm.encoder.layers[l]
            a = lay.self_attn
            tensors += [lay.layer_norm1.weight.detach(), lay.layer_norm1.bias.detach(),
                        torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).detach().contiguous(),
                        torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).detach().contiguous(),
                        a.out_proj.weight.detach().contiguous(), a.out_proj.bias.detach(),
                        lay.layer_norm2.weight.detach(), lay.layer_norm2.bias.detach(),
                        lay.mlp.fc1.weight.detach().contiguous(), lay.mlp.fc1.bias.detach(),
                        lay.mlp.fc2.weight.detach().contiguous(), lay.mlp.fc2.bias.detach()]
        self.tensors = tensors  # keep alive
        self.table = native.Context.pointer_table(tensors)
        self.cfg = dict(img=cfg.image_size, patch=cfg.patch_size, hidden=H, heads=cfg.num_attention_heads,
                        ffn=cfg.intermediate_size, layers_used=idx, k_pad=k_pad, act=act,
                        eps=float(cfg.layer_norm_eps))

    def ensure(self):
        sig = _sig(self.module.parameters())
        if sig != self._sig:
            if any(hasattr(m, "lora_A") for m in self.module.modules()):
                raise NotImplementedError("LoRA adapters on the CLIP tower are outside the u-LLaVA path "
                                          "(train_ullava.py:88-113 excludes vision_encoder)")
            self._pack()
            self._sig = sig

    def invalidate(self):
        self._sig = None

    def __call__(self, ctx: "native.Context", pixels: torch.Tensor) -> torch.Tensor:
        self.ensure()
        dt = self.tensors[0].dtype
        pixels = pixels.to(dt).contiguous()
        if pixels.shape[-1] != self.cfg["img"] or pixels.shape[-2] != self.cfg["img"]:
            raise ValueError(f"Input image size ({pixels.shape[-2]}*{pixels.shape[-1]}) doesn't match model "
                             f"({self.cfg['img']}*{self.cfg['img']}).")
        return ctx.vit_forward(self.table, len(self.tensors), pixels, self.cfg)


class LlamaStack:
    """LLaMA decoder weights packed for ullava_llama_forward + lm_head / embedding plumbing."""

    def __init__(self, llama_model, lm_head):
        self.model = llama_model
        self.lm_head = lm_head
        self._sig = None
        self._scratch = None
        self._rope = None
        self._session = None
        self.graph_launches = 0  # kernels launched through CUDA-graph replays (not seen by ullava_launch_count)

    def decode_session(self, ctx, batch: int, max_seq: int, embed_table, lm_head, keep_hidden: bool) -> "DecodeSession":
        self.ensure()
        key = (batch, max_seq, keep_hidden, embed_table.data_ptr(), lm_head.data_ptr(), self._sig)
        if self._session is None or self._session.key != key:
            self._session = None  # free the old KV cache before allocating the new one
            self._session = DecodeSession(self, ctx, batch, max_seq, embed_table, lm_head, keep_hidden)
        return self._session

    def _pack(self):
        cfg = self.model.config
        heads = cfg.num_attention_heads
        kvh = getattr(cfg, "num_key_value_heads", heads) or heads
        if kvh != heads:
            raise NotImplementedError("grouped-query attention is outside the u-LLaVA (LLaMA-7B) path")
        if getattr(cfg, "attention_bias", False) or getattr(cfg, "mlp_bias", False):
            raise NotImplementedError("LLaMA with biases is outside the u-LLaVA path")
        tensors = []
        ew = effective_weight   # LoRA adapters (PeftModel-wrapped llm) are folded into the packed copies
        for lay in self.model.layers:
            a, m = lay.self_attn, lay.mlp
            tensors += [lay.input_layernorm.weight.detach(),
                        torch.cat([ew(a.q_proj), ew(a.k_proj), ew(a.v_proj)], 0).contiguous(),
                        ew(a.o_proj).contiguous(),
                        lay.post_attention_layernorm.weight.detach(),
                        interleave_gate_up(ew(m.gate_proj), ew(m.up_proj)),
                        ew(m.down_proj).contiguous()]
        tensors.append(self.model.norm.weight.detach())
        self.tensors = tensors
        self.head_w = ew(self.lm_head).contiguous()          # == lm_head.weight itself when there is no adapter
        emb = self.model.embed_tokens
        if hasattr(emb, "lora_embedding_A") and len(emb.lora_embedding_A):
            raise NotImplementedError("LoRA adapters on embed_tokens are not supported on the B200 path")
        self.embed_w = getattr(emb, "base_layer", emb).weight.detach()
        self.table = native.Context.pointer_table(tensors)
        H = cfg.hidden_size
        self.cfg = dict(layers=len(self.model.layers), hidden=H, heads=heads, head_dim=H // heads,
                        ffn=cfg.intermediate_size, eps=float(cfg.rms_norm_eps))
        rp = getattr(cfg, "rope_parameters", None) or {}
        self.theta = float(rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0)))
        if rp.get("rope_type", "default") != "default":
            raise NotImplementedError("only default RoPE is on the u-LLaVA path")
        self.dtype = tensors[1].dtype
        self.device = tensors[1].device
        self._rope = None

    def _linears(self):
        out = [self.lm_head]
        for lay in self.model.layers:
            a, m = lay.self_attn, lay.mlp
            out += [a.q_proj, a.k_proj, a.v_proj, a.o_proj, m.gate_proj, m.up_proj, m.down_proj]
        return out

    def ensure(self):
        sig = _sig(list(self.model.parameters()) + list(self.lm_head.parameters()), self._linears())
        if sig != self._sig:
            self._pack()
            self._sig = sig
            self._session = None  # decode session (KV cache, captured graph) belongs to the old packed copies

    def invalidate(self):
        """Force a re-pack at the next call (for edits the signature cannot see, e.g. `p.data.copy_()` of a weight)."""
        self._sig = None

    def rope_tables(self, max_pos: int):
        if self._rope is None or self._rope[0].shape[0] < max_pos:
            hd = self.cfg["head_dim"]
            n = max(max_pos, 2048)
            # LlamaRotaryEmbedding (hf:models/llama/modeling_llama.py:74-135), fp32 on the host side
            inv = 1.0 / (self.theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            fr = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
            self._rope = (fr.cos().to(self.device).contiguous(), fr.sin().to(self.device).contiguous())
        return self._rope

    def new_cache(self, batch: int, max_seq: int) -> KVCache:
        self.ensure()
        c = self.cfg
        return KVCache(c["layers"], batch, c["heads"], max_seq, c["head_dim"], self.dtype, self.device)

    def scratch(self, ctx, rows: int) -> torch.Tensor:
        need = ctx.llama_scratch_bytes(rows, self.cfg["hidden"], self.cfg["ffn"])
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != self.device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._scratch

    def run(self, ctx, hidden: torch.Tensor, cache: KVCache, batch: int, seq: int, want_all_hidden=False,
            pos_offset=None):
        """hidden [batch*seq, H] is updated in place to the last layer's output; returns
        (final_norm_out [batch*seq, H], all_hidden [layers, batch*seq, H] or None)."""
        self.ensure()
        pos0 = cache.length
        if pos0 + seq > cache.max_seq:
            raise ValueError(f"KV cache too small: {pos0}+{seq} > {cache.max_seq}")
        cos, sin = self.rope_tables(cache.max_seq)
        final = torch.empty_like(hidden)
        allh = None
        if want_all_hidden:
            allh = torch.empty((self.cfg["layers"],) + tuple(hidden.shape), dtype=hidden.dtype, device=hidden.device)
        ctx.llama_forward(self.table, len(self.tensors), hidden, cache.k, cache.v, self.scratch(ctx, batch * seq),
                          batch, seq, pos0, self.cfg, cos, sin, final_out=final, all_hidden=allh,
                          pos_offset=pos_offset if seq == 1 else None)
        cache.length = pos0 + seq
        return final, allh


class DecodeSession:
    """Persistent device state of the greedy decode loop for one (batch, max_seq) shape: KV cache, step
    buffers, the device-side position and the CUDA graph of one decode step (ullava_llama_decode_step reads
    the position from device memory, so ONE captured graph is replayed for every step and every later
    generate() call of the same shape)."""

    def __init__(self, stack: "LlamaStack", ctx, batch: int, max_seq: int, embed_table: torch.Tensor,
                 lm_head: torch.Tensor, keep_hidden: bool):
        stack.ensure()
        self.stack, self.ctx, self.batch, self.max_seq = stack, ctx, batch, max_seq
        dev, dt = stack.device, stack.dtype
        H, V = stack.cfg["hidden"], lm_head.shape[0]
        self.key = (batch, max_seq, keep_hidden, embed_table.data_ptr(), lm_head.data_ptr(), stack._sig)
        self.cache = stack.new_cache(batch, max_seq)
        self.hidden = torch.empty((batch, H), dtype=dt, device=dev)
        self.final = torch.empty((batch, H), dtype=dt, device=dev)
        self.logits = torch.empty((batch, V), dtype=torch.float32, device=dev)
        self.cur_ids = torch.zeros((batch,), dtype=torch.int64, device=dev)
        self.pos = torch.zeros((1,), dtype=torch.int32, device=dev)
        # per-sample KV / RoPE position = pos + pos_offset[b] (<= 0): right-padded prompts of different lengths
        self.pos_offset = torch.zeros((batch,), dtype=torch.int32, device=dev)
        self.finished = torch.zeros((batch,), dtype=torch.uint8, device=dev)
        self.seqs = torch.zeros((batch, max_seq), dtype=torch.int64, device=dev)
        self.hid_buf = torch.empty((batch, max_seq - 1, H), dtype=dt, device=dev) if keep_hidden else None
        self.scratch = torch.empty(ctx.llama_scratch_bytes(batch, H, stack.cfg["ffn"]), dtype=torch.uint8, device=dev)
        self.embed_table, self.lm_head = embed_table, lm_head
        # decode-layer chains (csrc/gemm_chain_sm100.cu): per step context, the argument block with its own chain
        # program (the program holds TMA descriptors of this session's buffers and is partitioned over the context's SMs).
        # Bit-identical to the kernel-per-GEMM step, but measured SLOWER on B200 (profiles/r02_decode_chain.md: 163-176 vs
        # 154 us per layer), so it is opt-in: ULLAVA_DECODE_CHAIN=1.
        self.use_chain = batch <= 32 and os.environ.get("ULLAVA_DECODE_CHAIN", "0") == "1"
        self.chain_args = {}      # id(step context) -> (DecodeArgs, program buffer)
        self.graphs = {}          # id(step context) -> (CUDAGraph, kernel nodes): persistent grids are sized per context
        self.graph_nodes = 0
        self.eos_id, self.pad_id = -1, 0
        self.sampling = None      # None = greedy, else (temperature, top_p, top_k)
        self.uniforms = None      # [max_seq, batch] fp32: the draw of row b at position pos (ullava_sample_step)
        self.args = None

    def _build_args(self):
        cos, sin = self.stack.rope_tables(self.max_seq)
        a = native.DecodeArgs()
        self.ctx.fill_llama_args(a.llama, self.stack.table, len(self.stack.tensors), self.hidden, self.cache.k,
                                 self.cache.v, self.scratch, self.batch, 1, 0, self.stack.cfg, cos, sin,
                                 final_out=self.final, pos_offset=self.pos_offset)
        a.pos_dev = self.pos.data_ptr()
        a.embed_table, a.vocab, a.lm_head = self.embed_table.data_ptr(), self.lm_head.shape[0], self.lm_head.data_ptr()
        a.cur_ids, a.logits = self.cur_ids.data_ptr(), self.logits.data_ptr()
        a.seqs, a.seqs_ld = self.seqs.data_ptr(), self.seqs.stride(0)
        a.hid_buf = self.hid_buf.data_ptr() if self.hid_buf is not None else None
        a.hid_bs = self.hid_buf.stride(0) if self.hid_buf is not None else 0
        a.finished, a.eos_id, a.pad_id = self.finished.data_ptr(), self.eos_id, self.pad_id
        if self.sampling is not None:
            a.uniforms, a.uniforms_ld = self.uniforms.data_ptr(), self.uniforms.stride(0)
            a.temperature, a.top_p, a.top_k = float(self.sampling[0]), float(self.sampling[1]), int(self.sampling[2])
        self.args = a

    def begin(self, input_ids: torch.Tensor, eos_id, pad_id: int, sampling=None, generator=None, lengths=None):
        """sampling: None (greedy) or (temperature, top_p, top_k); the uniforms of every position are drawn here, once per
        generate() call, from `generator` (torch's default CUDA generator when None).  lengths (optional, [B] on the
        device): valid prompt lengths of a right-padded batch; sample b then continues at position lengths[b]."""
        if lengths is None:
            self.pos_offset.zero_()
        else:
            self.pos_offset.copy_((lengths - input_ids.shape[1]).to(torch.int32))
        eos = -1 if eos_id is None else int(eos_id)
        if sampling is not None:
            sampling = (float(sampling[0]), float(sampling[1]) if sampling[1] is not None else 1.0,
                        int(sampling[2]) if len(sampling) > 2 and sampling[2] else 0)
            if self.uniforms is None:
                self.uniforms = torch.empty((self.max_seq, self.batch), dtype=torch.float32, device=self.stack.device)
            self.uniforms.uniform_(0.0, 1.0, generator=generator)
        if self.args is None or eos != self.eos_id or int(pad_id) != self.pad_id or sampling != self.sampling:
            self.eos_id, self.pad_id, self.sampling = eos, int(pad_id), sampling
            self._build_args()
            self.graphs = {}   # eos / pad ids and the sampling parameters are baked into the captured kernel arguments
            self.chain_args = {}
        P = input_ids.shape[1]
        self.cache.length = 0
        self.finished.zero_()
        self.seqs.fill_(pad_id)
        self.seqs[:, :P] = input_ids

    def first_token(self, last_final: torch.Tensor, prompt_len: int):
        """Greedy token after the prefill: lm_head on the last prompt position + the step bookkeeping at pos = P-1."""
        self.ctx.gemm(last_final, self.lm_head, out=self.logits)
        self.pos.fill_(prompt_len - 1)
        if self.sampling is not None:
            self.ctx.sample_step(self.logits, self.sampling[0], self.sampling[1], self.uniforms, self.cur_ids, self.seqs,
                                 last_final, self.hid_buf, self.finished, self.eos_id, self.pad_id, self.pos,
                                 top_k=self.sampling[2])
        else:
            self.ctx.greedy_step(self.logits, self.cur_ids, self.seqs, last_final, self.hid_buf, self.finished,
                                 self.eos_id, self.pad_id, self.pos)

    def steps(self, n: int, use_graph: bool = True, ctx=None) -> int:
        """Runs n decode steps on the current stream; returns the number of native kernels launched through graph
        REPLAYS (eager launches are counted by ullava_launch_count; the phantom increments made while capturing are
        cancelled).  ctx: the context the steps are launched through (default: the session's) -- a partition lane's
        context when the decode loop runs on an SM partition (native.Partition); one graph is kept per context."""
        if n <= 0:
            return 0
        ctx = ctx or self.ctx
        args = self._step_args(ctx)
        replayed = 0
        done = 0
        entry = self.graphs.get(id(ctx))
        if use_graph and entry is None:
            ctx.llama_decode_step(args)  # eager warm-up step (also a real step)
            done = 1
            if n > 1:
                c0 = ctx.launch_count()
                g = torch.cuda.CUDAGraph()
                cur = torch.cuda.current_stream()
                # on a partition lane the capture has to happen on the lane's own (green-context) stream, so that the
                # kernel nodes stay confined to the lane's SMs; the legacy default stream cannot capture at all
                on_lane = cur != torch.cuda.default_stream(cur.device)
                with torch.cuda.graph(g, **(dict(stream=cur) if on_lane else {})):
                    ctx.llama_decode_step(args)
                nodes = ctx.launch_count() - c0
                replayed -= nodes  # capture bumped the native counter without launching anything
                entry = self.graphs[id(ctx)] = (g, nodes)
                self.graph_nodes = nodes
        for _ in range(n - done):
            if use_graph and entry is not None:
                entry[0].replay()
                replayed += entry[1]
            else:
                ctx.llama_decode_step(args)
        self.cache.length += n
        return replayed

    def _step_args(self, ctx):
        """The argument block of ullava_llama_decode_step for steps launched through `ctx`: with use_chain a copy of
        self.args that carries the chain program built for this context (its SM count fixes the stream-K partition)."""
        if not self.use_chain:
            return self.args
        entry = self.chain_args.get(id(ctx))
        if entry is None:
            a = native.DecodeArgs.from_buffer_copy(self.args)
            c = self.stack.cfg
            nbytes = ctx.llama_chain_bytes(c["layers"], c["hidden"], c["ffn"], self.lm_head.shape[0])
            prog = torch.empty(nbytes, dtype=torch.uint8, device=self.stack.device)
            a.llama.chain_program, a.llama.chain_bytes = prog.data_ptr(), nbytes
            torch.cuda.synchronize(self.stack.device)      # the program is written with a synchronous copy
            ctx.llama_chain_prepare(a)
            entry = self.chain_args[id(ctx)] = (a, prog)
        return entry[0]

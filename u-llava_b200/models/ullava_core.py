"""B200-native u-LLaVA core: CLIP ViT -> projector -> LLaMA, behind the reference's module API.

Mirrors /root/reference/models/ullava_core.py (class names, constructor, attribute tree, state_dict
keys, forward/generate signatures and return containers) so that inference_ullava_core.py and
UllavaForCausalLM use it unchanged.  The HuggingFace modules instantiated here hold parameters only:
every forward pass goes through libullava_sm100.so (see engine.py / native.py).  There is no torch
fallback: without a CUDA sm_100 device the forward raises.

Reference behaviour kept (file:line in /root/reference):
  * encode_image selects hidden_states[vision_hidden_layer] and drops CLS      (models/ullava_core.py:146-158)
  * image features are projected and spliced after <img_beg>                  (:182-277)
  * start/end token count assertion                                           (:209-211)
  * a [B,1] input_ids call with past_key_values is a decode step (no vision)  (:188-189)
  * forward returns CausalLMOutputWithPast(loss, logits, past_key_values, hidden_states, attentions) (:279-355)
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch
import torch.nn as nn
from transformers import (AutoConfig, AutoModelForCausalLM, CLIPVisionConfig, CLIPVisionModel, LlamaConfig,
                          LlamaForCausalLM, LlamaModel)
from transformers.modeling_outputs import CausalLMOutputWithPast

try:
    from utils.registry import registry  # present when dropped into the reference tree
except Exception:  # standalone use
    class _Registry:
        mapping = {"model_name_mapping": {}}

        @classmethod
        def register_model(cls, name):
            def wrap(model_cls):
                cls.mapping["model_name_mapping"][name] = model_cls
                return model_cls
            return wrap

    registry = _Registry()

import native
from .engine import KVCache, LlamaStack, VisionTower


class UllavaCoreConfig(LlamaConfig):
    model_type = "ullava_core"
    is_composition = True

    def __init__(self, vision_config=None, vision_hidden_layer=-1, projector_type='mlp',
                 projector_from_scratch=True, mm_token_ids=None, **kwargs):
        super().__init__(**kwargs)
        self.vision_hidden_layer = vision_hidden_layer
        self.mm_token_ids = mm_token_ids
        self.projector_type = projector_type
        self.projector_from_scratch = projector_from_scratch
        if isinstance(vision_config, CLIPVisionConfig):
            vision_config = vision_config.to_dict()
        self.vision_config = CLIPVisionConfig(**vision_config) if vision_config else {}

    def to_dict(self):
        output = copy.deepcopy(self.__dict__)
        output["vision_config"] = self.vision_config.to_dict() if self.vision_config else {}
        output["model_type"] = self.__class__.model_type
        return output


class GenerateOutput(dict):
    """Minimal stand-in for transformers' GenerateDecoderOnlyOutput: attribute + key access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


@registry.register_model('ullava_core')
class UllavaCoreForCausalLM(LlamaForCausalLM):
    config_class = UllavaCoreConfig

    def __init__(self, config: UllavaCoreConfig):
        super(LlamaForCausalLM, self).__init__(config)
        self.config = config
        self.model = LlamaModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.vision_encoder = CLIPVisionModel(config.vision_config)
        self.vision_projector = self.build_vision_projector(config.vision_config.hidden_size, config.hidden_size,
                                                            config.projector_type)
        self.vision_hidden_layer = config.vision_hidden_layer
        self.projector_from_scratch = config.projector_from_scratch
        self.mm_token_ids = config.mm_token_ids
        self.post_init()
        self._tower = None
        self._stack = None
        self.timeline = None  # optional list of (stage, torch.cuda.Event), see UllavaForCausalLM._mark
        self.use_cuda_graph = True  # greedy decode loop replays one captured step graph (engine.DecodeSession)

    def graph_kernel_launches(self) -> int:
        """Native kernels launched through CUDA-graph replays (they bypass ullava_launch_count)."""
        return self._stack.graph_launches if self._stack is not None else 0

    def _mark(self, name: str):
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((name, ev))

    @staticmethod
    def build_vision_projector(in_dim, hidden_dim, name='mlp'):
        if name == 'mlp':
            return nn.Linear(in_dim, hidden_dim)
        if name == 'mlp2x':
            return nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, hidden_dim))
        raise NotImplementedError

    def get_input_embeddings(self) -> nn.Module:
        return self.model.embed_tokens

    def get_output_embeddings(self) -> nn.Module:
        return self.lm_head

    def init_mm_tokens(self, tokenizer, mm_tokens):
        mm_token_ids = {k: tokenizer.convert_tokens_to_ids(v) for k, v in mm_tokens.items()}
        self.config.mm_token_ids = mm_token_ids
        self.mm_token_ids = self.config.mm_token_ids

    # ---- native plumbing -------------------------------------------------------------------
    def _ctx(self) -> "native.Context":
        dev = getattr(self.lm_head, "base_layer", self.lm_head).weight.device
        if dev.type != "cuda":
            raise RuntimeError("UllavaCoreForCausalLM (B200 build) runs on a CUDA sm_100 device only; "
                               "call .cuda() first -- there is no CPU fallback")
        return native.Context.get(dev)

    def _engine(self):
        if self._tower is None or self._tower.module is not self.vision_encoder or \
                self._tower.hidden_layer != self.vision_hidden_layer:
            self._tower = VisionTower(self.vision_encoder, self.vision_hidden_layer)
        if self._stack is None or self._stack.model is not self.model or self._stack.lm_head is not self.lm_head:
            self._stack = LlamaStack(self.model, self.lm_head)
        return self._tower, self._stack

    def invalidate_packed(self):
        """Drop the packed weight copies (fused QKV, interleaved gate/up, LoRA-merged matrices, decode session).  They
        are re-made automatically when a parameter is replaced, updated through autograd-visible in-place ops, or
        when LoRA adapters are merged / switched; call this after edits that leave no trace (`p.data.copy_(...)`)."""
        if self._tower is not None:
            self._tower.invalidate()
        if self._stack is not None:
            self._stack.invalidate()

    def _project(self, ctx, feats2d: torch.Tensor) -> torch.Tensor:
        vp = self.vision_projector
        if isinstance(vp, nn.Linear):
            return ctx.gemm(feats2d, vp.weight.detach(), bias=vp.bias.detach())
        h = ctx.gemm(feats2d, vp[0].weight.detach(), bias=vp[0].bias.detach(), epilogue=native.EPI_GELU)
        return ctx.gemm(h, vp[2].weight.detach(), bias=vp[2].bias.detach())

    def encode_image(self, image_tensors):
        """[bs,3,H,W] -> [bs, num_patches, vision_hidden] (CLS removed)."""
        with torch.no_grad():
            tower, _ = self._engine()
            return tower(self._ctx(), image_tensors)

    def encode_video(self, video_clip_tensors):
        """[bs, C, T, H, W] -> temporal+spatial pooled features [bs, T + num_patches, hidden]
        (reference :160-180).  The frames go through the native ViT, the two mean-poolings through ullava_video_pool."""
        bs, c, t, h, w = video_clip_tensors.shape
        frames = video_clip_tensors.permute(0, 2, 1, 3, 4).reshape(bs * t, c, h, w)
        feats = self.encode_image(frames).view(bs, t, -1, self.vision_encoder.config.hidden_size)
        return self._ctx().video_pool(feats)

    def embed_images_videos(self, input_ids=None, images=None, videos=None):
        if input_ids.shape[1] == 1:
            return input_ids, None
        ctx = self._ctx()
        _, stack = self._engine()
        stack.ensure()
        B, L = input_ids.shape
        table = stack.embed_w
        H = table.shape[1]
        embeds = ctx.embed_gather(input_ids, table).view(B, L, H)
        ids = self.mm_token_ids
        # one host round trip for the control decisions of the whole batch (the reference syncs per sample)
        flags = torch.stack([(input_ids == ids["IMG_START"]).sum(1), (input_ids == ids["IMG_END"]).sum(1),
                             (input_ids == ids["VID_START"]).sum(1), (input_ids == ids["VID_END"]).sum(1),
                             (input_ids == ids["IMG_START"]).int().argmax(1),
                             (input_ids == ids["VID_START"]).int().argmax(1)], 1).cpu()
        for b in range(B):
            n_is, n_ie, n_vs, n_ve = (int(v) for v in flags[b, :4])
            assert n_is == n_ie and n_vs == n_ve, \
                "Number of image start and end tokens should be the same. {0} vs {1}, {2} vs {3}".format(
                    n_is, n_ie, n_vs, n_ve)
        img_rows = [b for b in range(B) if int(flags[b, 0]) > 0]
        vid_rows = [b for b in range(B) if int(flags[b, 0]) == 0 and int(flags[b, 2]) > 0]
        if img_rows:
            feats = self.encode_image(images)  # [n_img, P, Hv]
            n_patch = feats.shape[1]
            proj = self._project(ctx, feats.reshape(-1, feats.shape[-1])).view(feats.shape[0], n_patch, H)
            start = torch.full((B,), -1, dtype=torch.int32)
            feat_rows = torch.zeros((B, n_patch, H), dtype=proj.dtype, device=proj.device) if len(img_rows) != B else None
            for i, b in enumerate(img_rows):
                start[b] = int(flags[b, 4])
                if int(start[b]) + 1 + n_patch > L:
                    raise ValueError("image patch tokens do not fit in the sequence")
                if feat_rows is not None:
                    feat_rows[b] = proj[i]
            ctx.splice_rows(embeds, proj if feat_rows is None else feat_rows, start.to(embeds.device))
        if vid_rows:
            vfeats = self.encode_video(videos)
            n_fp = vfeats.shape[1]
            vproj = self._project(ctx, vfeats.reshape(-1, vfeats.shape[-1])).view(vfeats.shape[0], n_fp, H)
            start = torch.full((B,), -1, dtype=torch.int32)
            rows = torch.zeros((B, n_fp, H), dtype=vproj.dtype, device=vproj.device)
            for i, b in enumerate(vid_rows):
                start[b] = int(flags[b, 5])
                rows[b] = vproj[i]
            ctx.splice_rows(embeds, rows, start.to(embeds.device))
        return None, embeds

    @staticmethod
    def _check_mask(attention_mask, B, L):
        """Right padding is harmless under a causal mask (valid positions never see pad tokens);
        anything else (left padding / holes) is not on the reference's inference path."""
        if attention_mask is None:
            return
        m = attention_mask.bool()
        if m.shape[-1] != L:
            return  # generation-time mask covering past + current tokens
        if bool(m.all()):
            return
        valid = m.int().sum(1)
        expect = torch.arange(L, device=m.device)[None, :] < valid[:, None]
        if not bool((m == expect).all()):
            raise NotImplementedError("only right-padded attention masks are supported on the B200 path")

    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None, past_key_values=None,
                inputs_embeds: Optional[torch.FloatTensor] = None, labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, images: Optional[torch.FloatTensor] = None,
                videos: Optional[torch.FloatTensor] = None, return_dict: Optional[bool] = None,
                logits_to_keep: int = 0, logits_fp32: bool = False, **kwargs):
        output_attentions = output_attentions if output_attentions is not None else self.config.output_attentions
        output_hidden_states = (output_hidden_states if output_hidden_states is not None
                                else self.config.output_hidden_states)
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash kernels")
        # position_ids are implied by the KV-cache length (equal-length / right-padded batches)
        only_last_hidden = bool(kwargs.pop("_return_last_hidden", False))
        caller_embeds = inputs_embeds is not None
        ctx = self._ctx()
        _, stack = self._engine()
        with torch.no_grad():
            if inputs_embeds is None:
                input_ids, inputs_embeds = self.embed_images_videos(input_ids, images, videos)
                if inputs_embeds is None:  # decode step: [B,1] token ids
                    stack.ensure()
                    inputs_embeds = ctx.embed_gather(input_ids, stack.embed_w).view(input_ids.shape[0], 1, -1)
            B, L, H = inputs_embeds.shape
            self._check_mask(attention_mask, B, L)
            if past_key_values is not None and not isinstance(past_key_values, KVCache):
                raise TypeError("past_key_values must be the KVCache returned by a previous forward(use_cache=True)")
            if past_key_values is not None:
                cache = past_key_values
            else:
                max_seq = L if not use_cache else max(L + 64, int(getattr(self.config, "native_max_seq", 0) or 0))
                cache = stack.new_cache(B, max_seq)
            hidden = inputs_embeds.reshape(B * L, H)
            if caller_embeds:
                hidden = hidden.to(stack.dtype).clone()  # the stack updates its input in place
            final, allh = stack.run(ctx, hidden, cache, B, L, want_all_hidden=bool(output_hidden_states))
            w = stack.head_w   # lm_head.weight with any LoRA adapter folded in (engine.effective_weight)
            if logits_to_keep:
                last = final.view(B, L, H)[:, -logits_to_keep:].reshape(-1, H)
                logits = ctx.gemm(last, w, out_f32=logits_fp32).view(B, logits_to_keep, -1)
            else:
                logits = ctx.gemm(final, w, out_f32=logits_fp32).view(B, L, -1)
            hs = None
            if output_hidden_states:
                hs = tuple(allh[l].view(B, L, H) for l in range(allh.shape[0])) + (final.view(B, L, H),)
            elif only_last_hidden:
                hs = (final.view(B, L, H),)

        loss = None
        if labels is not None:
            # reference :327-338 (CrossEntropyLoss over the shifted logits; inference callers ignore it)
            if logits.shape[1] == labels.shape[1] and logits.shape[1] > 1:
                loss = ctx.cross_entropy(logits, labels.to(logits.device))
            else:
                raise ValueError("labels must cover the positions the logits were computed for")

        pkv = cache if use_cache else None
        if not return_dict:
            out = (logits, pkv, hs)
            out = tuple(o for o in out if o is not None)
            return (loss,) + out if loss is not None else out
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=pkv, hidden_states=hs,
                                      attentions=None)

    # ---- generation --------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids=None, images=None, videos=None, max_new_tokens=32, num_beams=1, top_p=None,
                 do_sample=False, temperature=1.0, output_hidden_states=False, return_dict_in_generate=False,
                 no_repeat_ngram_size=None, stopping_criteria=None, eos_token_id=None, pad_token_id=None,
                 attention_mask=None, use_cache=True, generator=None, top_k=None, **kwargs):
        """Greedy / sampling generation with a KV cache (replaces HF GenerationMixin.generate for the
        path used by UllavaForCausalLM.evaluate, models/ullava.py:349-365, and inference_ullava_core.py:73-80).

        With output_hidden_states=True and return_dict_in_generate=True, `hidden_states[-1][-1]` is the
        post-final-norm hidden state of every processed position, [B, T-1, H] -- exactly what the
        reference reads when its checkpoints run with use_cache=False.

        Sampling (do_sample and temperature > 0) applies HF's warper chain: temperature, top-k, top-p.  top_k=None
        takes generation_config.top_k, i.e. the GenerationConfig default of 50 that the reference's
        `generate(do_sample=True, temperature=0.2, top_p=None)` call runs with; top_k=0 switches it off.
        top_p=None or 1 switches the nucleus filter off, top_p=0 keeps the most likely token only (HF:
        min_tokens_to_keep=1).

        stopping_criteria (HF StoppingCriteria callables, e.g. models.tools.KeywordsStoppingCriteria as passed by
        inference_ullava.py:88-102) are evaluated WITHOUT leaving the graph-replayed device loop: the loop runs in
        chunks of 8 tokens and the criteria are then called on every new prefix in order, exactly as HF would have
        called them token by token; the result is trimmed to the first hit."""
        if num_beams != 1:
            raise NotImplementedError("beam search is outside the u-LLaVA path (num_beams=1 everywhere)")
        if no_repeat_ngram_size:
            raise NotImplementedError("no_repeat_ngram_size is not supported on the B200 path")
        ctx = self._ctx()
        _, stack = self._engine()
        stack.ensure()
        B, P = input_ids.shape
        if eos_token_id is None:
            eos_token_id = getattr(self.config, "eos_token_id", None)
        if isinstance(eos_token_id, (list, tuple)):
            eos_token_id = eos_token_id[0] if eos_token_id else None
        if pad_token_id is None:
            pad_token_id = getattr(self.config, "pad_token_id", None)
            if pad_token_id is None:
                pad_token_id = eos_token_id if eos_token_id is not None else 0
        self._check_mask(attention_mask, B, P)
        # Right-padded prompts of different lengths in one batch (the reference generates one prompt at a time): the
        # prefill runs on the padded block, every sample then continues right after ITS last valid token.  Returned
        # sequences keep the HF layout [prompt | pads | new tokens]; hidden_states[-1][-1][:, j] is the state that
        # predicted token j + 1 for every generated token (column P - 1 holds the state of the last valid position).
        lengths = None
        if attention_mask is not None and attention_mask.shape[-1] == P and not bool(attention_mask.bool().all()):
            lengths = attention_mask.bool().sum(1).to(torch.int32)
            if int(lengths.min()) < 1:
                raise ValueError("every prompt needs at least one valid token")
        sampling = None
        if do_sample and temperature and temperature > 0:
            if top_p is not None and not (0.0 <= float(top_p) <= 1.0):
                raise ValueError(f"`top_p` has to be a float in [0, 1], but is {top_p}")   # TopPLogitsWarper
            if top_k is None:
                gk = getattr(getattr(self, "generation_config", None), "top_k", None)
                top_k = 50 if gk is None else int(gk)
            if int(top_k) < 0:
                raise ValueError(f"`top_k` has to be a non-negative integer, but is {top_k}")
            sampling = (float(temperature), top_p, int(top_k))
        if max_new_tokens < 1:
            seqs = input_ids.clone()
            return GenerateOutput(sequences=seqs, hidden_states=None, past_key_values=None) \
                if return_dict_in_generate else seqs
        after_prefill = kwargs.pop("_after_prefill", None)   # callable run once the prefill has been enqueued
        decode_lane = kwargs.pop("_decode_lane", None)       # (native.Context, torch stream) of an SM partition lane
        crit = []
        if stopping_criteria is not None:
            crit = list(stopping_criteria) if isinstance(stopping_criteria, (list, tuple)) or \
                hasattr(stopping_criteria, "__iter__") else [stopping_criteria]
        return self._generate_device(ctx, stack, input_ids, images, videos, max_new_tokens, eos_token_id,
                                     pad_token_id, output_hidden_states, return_dict_in_generate,
                                     sampling=sampling, generator=generator, lengths=lengths, criteria=crit,
                                     after_prefill=after_prefill, decode_lane=decode_lane)

    CRITERIA_CHUNK = 8   # decode steps between two host checks (EOS of every row / stopping criteria)

    def _generate_device(self, ctx, stack, input_ids, images, videos, max_new_tokens, eos_token_id, pad_token_id,
                         output_hidden_states, return_dict_in_generate, sampling=None, generator=None,
                         lengths=None, criteria=(), after_prefill=None, decode_lane=None):
        """Greedy (or, with `sampling` = (temperature, top_p, top_k), sampling) loop with all per-step state on the
        device: prefill, then max_new_tokens-1 replays of ONE captured decode-step graph (the position is read from
        device memory).  The host only synchronises every CRITERIA_CHUNK steps, and only when an eos id or stopping
        criteria are set.

        decode_lane = (context, stream) of one lane of an SM partition (native.Partition): the decode steps are then
        launched through that context into that stream (the prefill stays on the caller's stream and the whole
        machine); after_prefill() is called once the prefill is enqueued -- UllavaForCausalLM.evaluate uses the pair to
        run the SAM image encoder on the other lane while the HBM-bound decode steps run on this one."""
        B, P = input_ids.shape
        H = self.config.hidden_size
        T = P + max_new_tokens
        table, w = stack.embed_w, stack.head_w
        sess = stack.decode_session(ctx, B, T, table, w, bool(output_hidden_states))
        sess.begin(input_ids, eos_token_id, pad_token_id, sampling=sampling, generator=generator, lengths=lengths)
        _, embeds = self.embed_images_videos(input_ids, images, videos)
        self._mark("vit_projector_splice")
        final, _ = stack.run(ctx, embeds.view(B * P, H), sess.cache, B, P)
        if sess.hid_buf is not None:
            ctx.copy_rows(final, sess.hid_buf, B, P, H, P * H, H, sess.hid_buf.stride(0), H)
        self._mark("prefill")
        if after_prefill is not None:
            after_prefill()
        if lengths is None:
            last = final.view(B, P, H)[:, -1].contiguous()
        else:   # last VALID position of every row; first_token() also stores it as the state of column P - 1
            last = final.view(B, P, H)[torch.arange(B, device=final.device), (lengths - 1).long()].contiguous()
        sess.first_token(last, P)
        step_ctx = None
        caller_stream = torch.cuda.current_stream()
        if decode_lane is not None:
            step_ctx, lane_stream = decode_lane
            lane_stream.wait_stream(caller_stream)
            torch.cuda.set_stream(lane_stream)
        try:
            n_tokens = self._decode_loop(sess, P, max_new_tokens, eos_token_id, criteria, step_ctx, stack)
        finally:
            if decode_lane is not None:
                torch.cuda.set_stream(caller_stream)
                caller_stream.wait_stream(decode_lane[1])
        return self._finish_generate(sess, P, n_tokens, eos_token_id, output_hidden_states, return_dict_in_generate)

    def _decode_loop(self, sess, P, max_new_tokens, eos_token_id, criteria, step_ctx, stack) -> int:

        def first_hit(n_from, n_to):
            """Calls the criteria on the prefixes holding n_from+1 .. n_to generated tokens, in order, the way HF calls
            them after every token (scores are not materialised per step: None); returns the first count that stops."""
            for n in range(n_from + 1, n_to + 1):
                view = sess.seqs[:, :P + n]
                for c in criteria:
                    r = c(view, None)
                    if bool(r.all()) if isinstance(r, torch.Tensor) else bool(r):
                        return n
            return None

        remaining = max_new_tokens - 1
        n_tokens = 1
        stop = first_hit(0, 1) if criteria else None
        chunk = self.CRITERIA_CHUNK if (eos_token_id is not None or criteria) else max(remaining, 1)
        while remaining > 0 and stop is None:
            if eos_token_id is not None and bool(sess.finished.all()):
                break
            n = min(chunk, remaining)
            stack.graph_launches += sess.steps(n, use_graph=self.use_cuda_graph, ctx=step_ctx)
            if criteria:
                stop = first_hit(n_tokens, n_tokens + n)
            remaining -= n
            n_tokens += n
        return n_tokens if stop is None else stop

    def _finish_generate(self, sess, P, n_tokens, eos_token_id, output_hidden_states, return_dict_in_generate):
        seqs = sess.seqs[:, :P + n_tokens]
        if eos_token_id is not None:
            # HF stops as soon as every sequence has emitted eos: trim the pad-only columns of the last chunk
            gen = seqs[:, P:]
            is_eos = gen == eos_token_id
            if bool(is_eos.any(1).all()):
                first = torch.where(is_eos.any(1), is_eos.int().argmax(1), torch.full_like(gen[:, 0], gen.shape[1]))
                seqs = seqs[:, :P + int(first.max()) + 1]
        sess.cache.length = seqs.shape[1] - 1   # rows written by discarded steps of the last chunk are ignored
        seqs = seqs.clone()
        if not return_dict_in_generate:
            return seqs
        hs = None
        if output_hidden_states:
            hs = ((sess.hid_buf[:, : seqs.shape[1] - 1].clone(),),)
        return GenerateOutput(sequences=seqs, hidden_states=hs, past_key_values=sess.cache)

    def prepare_inputs_for_generation(self, input_ids=None, inputs_embeds=None, attention_mask=None, images=None,
                                      videos=None, labels=None, past_key_values=None, **kwargs):
        """Kept for API parity (reference :357-395); generate() above does not need it."""
        if past_key_values:
            input_ids = input_ids[:, -1:]
        model_inputs = {"inputs_embeds": inputs_embeds} if (inputs_embeds is not None and past_key_values is None) \
            else {"input_ids": input_ids}
        model_inputs.update({"position_ids": kwargs.get("position_ids"), "past_key_values": past_key_values,
                             "use_cache": kwargs.get("use_cache"), "attention_mask": attention_mask,
                             "images": images, "videos": videos})
        return model_inputs


AutoConfig.register("ullava_core", UllavaCoreConfig)
AutoModelForCausalLM.register(UllavaCoreConfig, UllavaCoreForCausalLM)

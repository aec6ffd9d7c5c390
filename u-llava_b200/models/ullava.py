"""B200-native u-LLaVA with the segmentation / detection heads, behind the reference's module API
(/root/reference/models/ullava.py).  Same constructor, attribute tree (llm, seg_projector, visual_model,
det_projector, det_decoder), state_dict keys, forward(inference=True) dict and evaluate() tuple.

What changes underneath (all arithmetic in libullava_sm100.so):
  * seg/det projectors are applied to the gathered [SEG]/[LOC] rows only (the reference projects all
    B*L rows and then masks, models/ullava.py:190-199) -- same values, n rows instead of B*L;
  * all prompts of all images go through ONE batched mask-decoder call (reference: Python loop over
    images, :231-256) and the fused double-bilinear post-process;
  * SAM image embeddings are computed batched (reference: one image at a time with empty_cache(), :139-150);
  * no hard-coded .cuda(): tensors are created on the model's device.
Training (inference=False) is outside the hot path and raises."""
from __future__ import annotations

import copy
import os
import warnings
from typing import List

import torch
import torch.nn as nn
from transformers import AutoConfig, AutoModelForCausalLM, PretrainedConfig, PreTrainedModel

import native
from models.segment_anything import build_sam_vit_h
from models.ullava_core import UllavaCoreConfig, UllavaCoreForCausalLM, registry


class UllavaConfig(PretrainedConfig):
    model_type = "ullava"
    is_composition = True

    def __init__(self, llm_config=None, ce_weight=1.0, bce_weight=2.0, dice_weight=0.5, l1_weight=1.0,
                 iou_weight=1.0, out_dim=256, seg_token_idx=32007, loc_token_idx=32008, train_mask_decoder=True,
                 **kwargs):
        super().__init__(**kwargs)
        if isinstance(llm_config, UllavaCoreConfig):
            llm_config = llm_config.to_dict()
        self.llm_config = UllavaCoreConfig(**llm_config) if llm_config else {}
        self.ce_weight, self.bce_weight, self.dice_weight = ce_weight, bce_weight, dice_weight
        self.l1_weight, self.iou_weight = l1_weight, iou_weight
        self.out_dim = out_dim
        self.seg_token_idx, self.loc_token_idx = seg_token_idx, loc_token_idx
        self.train_mask_decoder = train_mask_decoder

    def to_dict(self):
        output = copy.deepcopy(self.__dict__)
        output["llm_config"] = self.llm_config.to_dict() if self.llm_config else {}
        output["model_type"] = self.__class__.model_type
        return output


def _head(in_dim, out_dim):
    return nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(inplace=True), nn.Linear(in_dim, out_dim), nn.Dropout(0.0))


@registry.register_model('ullava')
class UllavaForCausalLM(PreTrainedModel):
    config_class = UllavaConfig
    sam_builder = staticmethod(build_sam_vit_h)  # tests may swap in a smaller image encoder

    def __init__(self, config):
        super().__init__(config)
        self.config = config
        llm_config = config.llm_config
        self.llm = UllavaCoreForCausalLM(llm_config)
        self.seg_projector, self.visual_model = self.init_seg_modules(llm_config.hidden_size)
        self.det_projector, self.det_decoder = self.init_det_modules(llm_config.hidden_size)
        # optional extras for the sharded eval path (dist_eval.py): when pack_mask_bits is set, evaluate()/forward()
        # also keep the thresholded masks bit-packed ([n, ceil(H*W/32)] int32 per image) in last_mask_bits
        self.pack_mask_bits = False
        self.last_mask_bits = None
        self.timeline = None  # list of (stage name, torch.cuda.Event) when stage timing is requested
        # evaluate(): run the SAM ViT-H image encoder (tensor-bound, depends on images_sam only) CONCURRENTLY with the
        # decode steps (HBM-bound) on a spatial split of the SMs (native.Partition, CUDA green contexts): the decode
        # lane gets `overlap_sms_decode` SMs and the higher stream priority, the encoder the rest.
        # The decode steps barely slow down on ~100 of the 148 SMs while the encoder, on the remaining ones, gets
        # through most of its blocks in that time; its last blocks + neck then run on the whole machine
        # (`overlap_sam_blocks` = blocks done on the lane).  Measured curves: tools/bench_overlap.py, DESIGN.md section 5.
        self.overlap_sam = os.environ.get("ULLAVA_OVERLAP", "1") != "0"
        self.overlap_sms_decode = int(os.environ.get("ULLAVA_OVERLAP_SMS", "88"))
        # blocks of the encoder run on the lane; 0 = as many as fit under the decode steps (_blocks_under_decode)
        self.overlap_sam_blocks = int(os.environ.get("ULLAVA_OVERLAP_BLOCKS", "0"))
        self.overlap_min_batch = 12  # below this the split does not pay (B = 8: 348.6 ms with it, 343.8 ms without;
                                     # B = 16: 452 vs 481 ms; B = 32: 741 vs 776 ms)

    def _mark(self, name: str):
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((name, ev))

    def init_det_modules(self, hidden_size):
        out_dim = self.config.out_dim
        projector = _head(hidden_size, out_dim)
        decoder = nn.Sequential(nn.Linear(out_dim, out_dim), nn.ReLU(inplace=True), nn.Linear(out_dim, out_dim // 2),
                                nn.ReLU(inplace=True), nn.Linear(out_dim // 2, 4))
        return projector, decoder

    def init_seg_modules(self, hidden_size):
        projector = _head(hidden_size, self.config.out_dim)
        visual_model = type(self).sam_builder(checkpoint=None)
        for p in visual_model.parameters():
            p.requires_grad = False
        return projector, visual_model

    def load_visual_checkpoint(self, checkpoint):
        with open(checkpoint, "rb") as f:
            state_dict = torch.load(f)
        self.visual_model.load_state_dict(state_dict, strict=False)

    def _init_weights(self, module):
        return  # sub-modules initialise themselves; nothing extra (keeps PreTrainedModel.post_init cheap)

    # ---- stages ----------------------------------------------------------------------------------
    def get_visual_embs(self, pixel_values: torch.FloatTensor):
        """SAM image embeddings [B,256,64,64] (reference :139-150), batched."""
        with torch.no_grad():
            enc = self.visual_model.image_encoder
            dt = next(enc.parameters()).dtype
            return enc(pixel_values.to(dt))

    def _mlp_head(self, ctx, seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
        """Linear(+ReLU) chains of seg_projector / det_projector / det_decoder on gathered rows."""
        lins = [m for m in seq if isinstance(m, nn.Linear)]
        for i, lin in enumerate(lins):
            x = ctx.gemm(x, lin.weight.detach(), bias=lin.bias.detach(),
                         epilogue=native.EPI_RELU if i < len(lins) - 1 else native.EPI_NONE)
        return x

    def _seg_loc_counts(self, token_ids):
        """[SEG] / [LOC] tokens per sample, on the host: the ONE synchronisation the mask / box heads need for their
        control flow.  evaluate() takes it before it enqueues the SAM image encoder, so that the head kernels are
        queued behind the encoder instead of after a round trip at its end."""
        seg_mask = token_ids[:, 1:] == self.config.seg_token_idx
        loc_mask = token_ids[:, 1:] == self.config.loc_token_idx
        return torch.stack([seg_mask.sum(1), loc_mask.sum(1)], 0).cpu()

    def _decode_heads(self, token_ids, hidden, image_embeddings, raw_size_list, resize_list, pack_bits=False,
                      counts=None):
        """Shared tail of forward(inference=True) and evaluate (reference :168-256, :364-432).
        token_ids [B,T]; hidden [B,>=T-1,H] post-final-norm; position j is used when token j+1 is [SEG]/[LOC]."""
        ctx = native.Context.get(hidden.device)
        B, T = token_ids.shape
        H = hidden.shape[-1]
        hid = hidden[:, : T - 1]
        seg_mask = token_ids[:, 1:] == self.config.seg_token_idx
        loc_mask = token_ids[:, 1:] == self.config.loc_token_idx
        if counts is None:
            counts = self._seg_loc_counts(token_ids)  # one host sync for the control flow
        seg_counts, loc_counts = counts[0].tolist(), counts[1].tolist()
        sam = self.visual_model
        dt = hidden.dtype
        pred_masks: List[torch.Tensor] = [None] * B
        bits_list = [None] * B
        n_seg = sum(seg_counts)
        if n_seg > 0:
            rows = hid[seg_mask].contiguous()  # [n_seg, H] row-major over (b, position) like the reference
            emb = self._mlp_head(ctx, self.seg_projector, rows)  # [n_seg, 256]
            sparse, dense = sam.prompt_encoder(points=None, boxes=None, masks=None, text_embeds=emb.unsqueeze(1))
            sparse = sparse.to(dt)
            prompt_image = torch.repeat_interleave(torch.arange(B, device=hid.device, dtype=torch.int32),
                                                   torch.tensor(seg_counts, device=hid.device))
            masks, _ = sam.mask_decoder.predict_masks_batched(
                image_embeddings, prompt_image, sam.prompt_encoder.get_dense_pe().reshape(256, 64, 64),
                sparse[:, 0], sam.prompt_encoder.no_mask_embed.weight.detach().reshape(-1).to(dt))
            off = 0
            for i in range(B):
                n = seg_counts[i]
                low = masks[off: off + n, 0:1]  # multimask_output=False
                off += n
                if n == 0:
                    pred_masks[i] = torch.zeros((0,) + tuple(int(v) for v in raw_size_list[i]), device=hid.device)
                    continue
                res = sam.postprocess_masks(low, input_size=resize_list[i], original_size=raw_size_list[i],
                                            pack_bits=pack_bits)
                if pack_bits:
                    res, bits_list[i] = res
                pred_masks[i] = res[:, 0]
        else:
            for i in range(B):
                pred_masks[i] = torch.zeros((0,) + tuple(int(v) for v in raw_size_list[i]), device=hid.device)
        pred_boxes: List[torch.Tensor] = []
        n_loc = sum(loc_counts)
        if n_loc > 0:
            le = self._mlp_head(ctx, self.det_projector, hid[loc_mask].contiguous())
            boxes = self._mlp_head(ctx, self.det_decoder, le)
            off = 0
            for i in range(B):
                pred_boxes.append(boxes[off: off + loc_counts[i]])
                off += loc_counts[i]
        else:
            pred_boxes = [torch.zeros((0, 4), dtype=dt, device=hid.device) for _ in range(B)]
        return pred_masks, pred_boxes, bits_list

    def forward(self, images_sam: torch.FloatTensor, images: torch.FloatTensor, input_ids: torch.LongTensor,
                labels: torch.LongTensor, attention_mask: torch.LongTensor, mask_list: List[torch.FloatTensor],
                size_list: List[torch.Tensor], resize_list: List[tuple], bbox_list: List[torch.FloatTensor],
                inference: bool = False):
        if not inference:
            raise NotImplementedError("training losses are outside the B200 inference hot path "
                                      "(SURVEY.md section 8: forward(inference=True) and evaluate() are in scope)")
        with torch.no_grad():
            self._mark("start")
            image_embeddings = self.get_visual_embs(images_sam)
            self._mark("sam_encoder")
            output = self.llm.forward(images=images, attention_mask=attention_mask, input_ids=input_ids, labels=labels,
                                      output_hidden_states=False, use_cache=False, _return_last_hidden=True)
            self._mark("vit_prefill_lmhead")
            last_hidden = output.hidden_states[-1]
            pred_masks, pred_boxes, bits = self._decode_heads(input_ids, last_hidden, image_embeddings, size_list,
                                                              resize_list, pack_bits=self.pack_mask_bits)
            self.last_mask_bits = bits if self.pack_mask_bits else None
            self._mark("mask_heads")
        return {"pred_masks": pred_masks, "pred_boxes": pred_boxes, "gt_masks": mask_list, "gt_boxes": bbox_list,
                "logits": output.logits}

    def _partition(self, device, batch: int):
        """The SM partition evaluate() overlaps on, or None (switched off, batch too small, driver without green
        contexts -- the stages then simply run back to back on the caller's stream)."""
        if not self.overlap_sam or batch < self.overlap_min_batch or max(1, int(self.overlap_sms_decode)) < 8:
            return None
        try:
            return native.Partition.get(device, int(self.overlap_sms_decode))
        except native.NativeError as e:
            warnings.warn(f"SM partition unavailable ({e}); evaluate() runs its stages back to back")
            self.overlap_sam = False
            return None

    def _blocks_under_decode(self, enc, batch: int, max_new_tokens: int, sms_lane: int) -> int:
        """How many encoder blocks fit under the decode steps.  Decode step: the 16-bit LLaMA weights + the KV rows
        stream once per token at ~4.5 TB/s (+ ~10 % on the smaller lane); encoder block on the lane: its GEMM /
        attention FLOPs at ~9 TFLOP/s per SM (tools/bench_overlap.py: ViT-H, B = 32, 60 SMs -> 10.5 ms per block with the
        round-2 attention kernels; bench.py --overlap-blocks 28 / 30 / 32 on one box: 695 / 687 / 688 ms per step).
        A block too many costs more (it runs on a fraction of the machine while the rest idles) than a block too few
        (it runs at full speed afterwards), hence the 0.9."""
        c = self.llm.config
        w_bytes = 2.0 * (c.num_hidden_layers * (4 * c.hidden_size ** 2 + 3 * c.hidden_size * c.intermediate_size) +
                         c.vocab_size * c.hidden_size)
        kv_bytes = batch * c.num_hidden_layers * 2 * 640 * c.hidden_size * 2.0
        t_decode = 1.1 * max(max_new_tokens - 1, 0) * (w_bytes + kv_bytes) / 4.5e12
        n_tok = (enc.img_size // enc.patch_size) ** 2
        d = enc.embed_dim
        flop_block = batch * n_tok * (24.0 * d * d + 4.0 * d * min(n_tok, max(enc.window_size, 1) ** 2 * 4))
        t_block = flop_block / (9.0e12 * max(sms_lane, 1))
        return max(1, min(enc.depth, int(0.9 * t_decode / max(t_block, 1e-9))))

    def evaluate(self, images_sam, images, input_ids, raw_size_list, resize_list, max_new_tokens=32, temperature=0.2,
                 top_p=None, num_beams=1, no_repeat_ngram_size=None, stopping_criteria=None, attention_mask=None):
        """Reference signature (models/ullava.py:335-345) plus `attention_mask` (optional): a right-padded batch of
        prompts of different lengths, as the reference collator builds it (dataset/collators/base_collator.py:28-44).

        Same results as the reference's order (generate, then get_visual_embs, :349-399); the image encoder does not
        depend on the generated tokens, so with overlap_sam it is enqueued on the second lane of an SM partition as soon
        as the prefill is queued and runs under the decode steps."""
        with torch.inference_mode():
            self._mark("start")
            self.llm.timeline = self.timeline
            part = self._partition(input_ids.device, input_ids.shape[0])
            side = {}
            kw = {}
            enc = self.visual_model.image_encoder
            if part is not None and images_sam.shape[0] > enc.max_images_per_pass:
                part = None                                  # several encoder passes: keep the simple order
            if part is not None:
                main = torch.cuda.current_stream()
                px = images_sam.to(next(enc.parameters()).dtype)
                k = int(self.overlap_sam_blocks)
                k = max(1, min(k, enc.depth)) if k > 0 else self._blocks_under_decode(
                    enc, input_ids.shape[0], max_new_tokens, part.sms[1])

                def after_prefill():
                    lane_ctx, lane_stream = part.ctx[1], part.streams[1]
                    lane_stream.wait_stream(main)           # inputs (and the prefill before them) are ordered first
                    with torch.cuda.stream(lane_stream):
                        enc.native_ctx = lane_ctx
                        try:
                            side["emb"] = enc.forward_blocks(px, 0, k)       # None unless k == depth
                        finally:
                            enc.native_ctx = None
                        side["done"] = torch.cuda.Event()
                        side["done"].record(lane_stream)
                    side["k"] = k

                kw = dict(_after_prefill=after_prefill, _decode_lane=(part.ctx[0], part.streams[0]))
            outputs = self.llm.generate(input_ids=input_ids, images=images, max_new_tokens=max_new_tokens,
                                        num_beams=num_beams, top_p=top_p, do_sample=True if temperature > 0 else False,
                                        temperature=temperature, output_hidden_states=True,
                                        return_dict_in_generate=True, no_repeat_ngram_size=no_repeat_ngram_size,
                                        stopping_criteria=stopping_criteria, attention_mask=attention_mask, **kw)
            output_ids = outputs.sequences
            last_hidden = outputs.hidden_states[-1][-1]
            self.llm.timeline = None
            self._mark("decode")
            counts = self._seg_loc_counts(output_ids)   # host sync here, while nothing else is queued
            if "done" in side:
                torch.cuda.current_stream().wait_event(side["done"])
                image_embeddings = side["emb"]
                if image_embeddings is None:                 # remaining blocks + neck on the whole machine
                    image_embeddings = enc.forward_blocks(px, side["k"], enc.depth)
                else:
                    image_embeddings.record_stream(torch.cuda.current_stream())
            else:
                image_embeddings = self.get_visual_embs(images_sam)
            self._mark("sam_encoder")
            pred_masks, pred_boxes, bits = self._decode_heads(output_ids, last_hidden, image_embeddings, raw_size_list,
                                                              resize_list, pack_bits=self.pack_mask_bits, counts=counts)
            self.last_mask_bits = bits if self.pack_mask_bits else None
            self._mark("mask_heads")
        return output_ids, pred_masks, pred_boxes


AutoConfig.register("ullava", UllavaConfig)
AutoModelForCausalLM.register(UllavaConfig, UllavaForCausalLM)

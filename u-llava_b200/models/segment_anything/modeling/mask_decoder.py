"""SAM mask decoder behind the reference's MaskDecoder API
(segment_anything/modeling/mask_decoder.py:75-164); arithmetic in csrc/sam_decoder.cu."""
from typing import List, Tuple

import torch
import torch.nn as nn

import native

from .common import LayerNorm2d


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, sigmoid_output=False):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


def _attn_params(a) -> List[torch.Tensor]:
    return [a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias,
            a.out_proj.weight, a.out_proj.bias]


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim, transformer, num_multimask_outputs=3, activation=nn.GELU,
                 iou_head_depth=3, iou_head_hidden_dim=256):
        super().__init__()
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4), activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2), activation())
        self.output_hypernetworks_mlps = nn.ModuleList(
            MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens))
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self._packed = None

    # ---- weight table in the order of enum SamW (csrc/sam_decoder.cu) ---------------------------
    def _pack(self, no_mask_embed: torch.Tensor):
        tr = self.transformer
        if (self.transformer_dim != 256 or tr.depth != 2 or tr.num_heads != 8 or tr.mlp_dim != 2048
                or self.num_mask_tokens != 4 or len(self.iou_prediction_head.layers) != 3
                or not isinstance(tr.layers[0].mlp.act, nn.ReLU)
                or not isinstance(self.output_upscaling[2], nn.GELU)):
            raise NotImplementedError("the native decoder is compiled for the build_sam_vit_* decoder geometry")
        t: List[torch.Tensor] = [self.iou_token.weight, self.mask_tokens.weight, no_mask_embed.reshape(-1)]
        for blk in tr.layers:
            t += _attn_params(blk.self_attn) + [blk.norm1.weight, blk.norm1.bias]
            t += _attn_params(blk.cross_attn_token_to_image) + [blk.norm2.weight, blk.norm2.bias]
            t += [blk.mlp.lin1.weight, blk.mlp.lin1.bias, blk.mlp.lin2.weight, blk.mlp.lin2.bias]
            t += [blk.norm3.weight, blk.norm3.bias, blk.norm4.weight, blk.norm4.bias]
            t += _attn_params(blk.cross_attn_image_to_token)
        t += _attn_params(tr.final_attn_token_to_image) + [tr.norm_final_attn.weight, tr.norm_final_attn.bias]
        ct1, ln, ct2 = self.output_upscaling[0], self.output_upscaling[1], self.output_upscaling[3]
        # ConvTranspose2d weight [ci, co, dy, dx] -> GEMM B operand [(dy,dx,co), ci]; bias repeated per (dy,dx)
        t += [ct1.weight.permute(2, 3, 1, 0).reshape(-1, ct1.weight.shape[0]), ct1.bias.repeat(4), ln.weight, ln.bias,
              ct2.weight.permute(2, 3, 1, 0).reshape(-1, ct2.weight.shape[0]), ct2.bias.repeat(4)]
        for m in self.output_hypernetworks_mlps:
            for lin in m.layers:
                t += [lin.weight, lin.bias]
        for lin in self.iou_prediction_head.layers:
            t += [lin.weight, lin.bias]
        t = [x.detach().contiguous() for x in t]
        assert len(t) == native.SAM_N_WEIGHTS
        return t, native.Context.pointer_table(t)

    def _table(self, no_mask_embed):
        params = list(self.parameters()) + [no_mask_embed]
        sig = tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in params)
        if self._packed is None or self._packed[0] != sig:
            tensors, table = self._pack(no_mask_embed)
            self._packed = (sig, tensors, table)
        return self._packed[1], self._packed[2]

    def predict_masks_batched(self, image_embeddings, prompt_image, image_pe, text_embeds, no_mask_embed):
        """All prompts of all images in one call.  image_embeddings [n_img,256,64,64]; prompt_image int32 [n]
        (image index per prompt); text_embeds [n,256].  Returns (masks [n,4,256,256], iou [n,4])."""
        dev = text_embeds.device
        if dev.type != "cuda":
            raise RuntimeError("MaskDecoder (B200 build) runs on a CUDA sm_100 device only")
        ctx = native.Context.get(dev)
        tensors, table = self._table(no_mask_embed)
        dt = tensors[0].dtype
        return ctx.sam_mask_decoder(table, len(tensors), image_embeddings.to(dt).contiguous(),
                                    prompt_image.to(torch.int32).contiguous(), image_pe.to(dt).contiguous(),
                                    text_embeds.to(dt).contiguous())

    def predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings):
        n = sparse_prompt_embeddings.shape[0]
        if sparse_prompt_embeddings.shape[1] != 1:
            raise NotImplementedError("the u-LLaVA path has exactly one sparse (text) prompt token per mask")
        d = dense_prompt_embeddings
        if d.dim() == 4 and d.stride(0) == 0 and d.stride(2) == 0 and d.stride(3) == 0:
            no_mask = d[0, :, 0, 0]  # broadcast view of prompt_encoder.no_mask_embed
            emb = image_embeddings
        else:  # genuinely dense prompt: fold it into the image embedding (off the u-LLaVA path)
            if image_embeddings.shape[0] != 1 and image_embeddings.shape[0] != n:
                raise ValueError("image_embeddings batch must be 1 or n_prompts")
            emb = (image_embeddings + d).contiguous()
            no_mask = torch.zeros(self.transformer_dim, dtype=emb.dtype, device=emb.device)
        if emb.shape[0] == 1:
            idx = torch.zeros(n, dtype=torch.int32, device=emb.device)
        else:
            idx = torch.arange(n, dtype=torch.int32, device=emb.device)
        return self.predict_masks_batched(emb, idx, image_pe.reshape(256, 64, 64), sparse_prompt_embeddings[:, 0],
                                          no_mask.to(sparse_prompt_embeddings.dtype))

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        masks, iou_pred = self.predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings,
                                             dense_prompt_embeddings)
        sl = slice(1, None) if multimask_output else slice(0, 1)
        return masks[:, sl, :, :], iou_pred[:, sl]

"""B200-native UniPose deformable layers: drop-ins for
`MSDeformAttn` (visionllmv2/model/unipose/ops/modules/ms_deform_attn.py:33-158),
`DeformableTransformerEncoderLayer` and `DeformableTransformerDecoderLayer`
(visionllmv2/model/unipose/modeling_unipose.py:3132-3186, 3188-3323) -- the consumers of the MSDA operator in the
pose-estimation decoder.  Same constructor arguments, parameter names (`self_attn.{sampling_offsets,attention_weights,
value_proj,output_proj}`, `norm1/2/3`, `linear1/2`, `ca_text`, `catext_norm`, nn.MultiheadAttention's
`in_proj_weight/in_proj_bias/out_proj`) and call signatures (sequence-first `[nq, bs, d]` tensors in the decoder layer,
`True = padding` masks), on the kernels of the Grounding-DINO layers (visionllm_b200/gdino.py): packed
offsets|weights GEMM, MSDA kernel, fused attention, GEMM epilogues (+bias, ReLU/GELU, +residual), LayerNorm kernel.
Forward only; dropouts are identities at eval like in the reference.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .gdino import GroundingDinoMultiscaleDeformableAttention, _LN, _MHA


class MSDeformAttn(GroundingDinoMultiscaleDeformableAttention):
    """ms_deform_attn.py:33-158.  `input_padding_mask`: True = padding (the GDINO module takes True = valid)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, use_4D_normalizer=False):
        super().__init__(SimpleNamespace(d_model=d_model, num_feature_levels=n_levels), num_heads=n_heads,
                         n_points=n_points)
        if use_4D_normalizer:
            raise NotImplementedError("use_4D_normalizer (never set by the UniPose builders, modeling_unipose.py:3144,3200)")
        self.use_4D_normalizer = use_4D_normalizer

    @torch.no_grad()
    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        valid = None if input_padding_mask is None else ~input_padding_mask
        out, _ = super().forward(hidden_states=query, attention_mask=valid, encoder_hidden_states=input_flatten,
                                 encoder_attention_mask=valid, position_embeddings=None,
                                 reference_points=reference_points, spatial_shapes=input_spatial_shapes,
                                 level_start_index=input_level_start_index)
        return out


def _activation(name):
    if name not in ("relu", "gelu"):
        raise NotImplementedError(f"activation={name} (the UniPose configs use relu)")
    return name


class DeformableTransformerEncoderLayer(nn.Module):
    """modeling_unipose.py:3132-3186."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 add_channel_attention=False, use_deformable_box_attn=False, box_attn_type="roi_align"):
        super().__init__()
        if add_channel_attention:
            raise NotImplementedError("add_channel_attention (DyReLU branch, off in the UniPose configs)")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = _LN(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.act = _activation(activation)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = _LN(d_model)
        self.add_channel_attention = False

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    @torch.no_grad()
    def forward_ffn(self, src):
        h = ops.linear(src, self.linear1.weight, bias=self.linear1.bias, act=self.act)
        return self.norm2(ops.linear(h, self.linear2.weight, bias=self.linear2.bias, residual=src))

    @torch.no_grad()
    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        src2 = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes, level_start_index,
                              key_padding_mask)
        return self.forward_ffn(self.norm1(src + src2))


class DeformableTransformerDecoderLayer(nn.Module):
    """modeling_unipose.py:3188-3323: self-attention, optional text cross-attention, deformable cross-attention, FFN;
    tensors are sequence-first ([nq, bs, d]) like the reference's nn.MultiheadAttention calls."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 use_text_feat_guide=False, use_text_cross_attention=False, ffn_extra_layernorm=False):
        super().__init__()
        assert not use_text_feat_guide
        if ffn_extra_layernorm:
            raise NotImplementedError("ffn_extra_layernorm not implemented")         # the reference raises too (:3222)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = _LN(d_model)
        if use_text_cross_attention:
            self.ca_text = _MHA(d_model, n_heads, dropout=dropout)
            self.catext_norm = _LN(d_model)
        self.self_attn = _MHA(d_model, n_heads, dropout=dropout)
        self.norm2 = _LN(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.act = _activation(activation)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm3 = _LN(d_model)
        self.norm_ext = None
        self.key_aware_proj = None
        self.use_text_feat_guide = use_text_feat_guide
        self.use_text_cross_attention = use_text_cross_attention
        self.n_heads = n_heads

    def rm_self_attn_modules(self):
        self.self_attn = None
        self.norm2 = None

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    @torch.no_grad()
    def forward_ffn(self, tgt):
        h = ops.linear(tgt, self.linear1.weight, bias=self.linear1.bias, act=self.act)
        return self.norm3(ops.linear(h, self.linear2.weight, bias=self.linear2.bias, residual=tgt))

    def _attend_mask(self, self_attn_mask, bs, nq):
        """nn.MultiheadAttention attn_mask ([nq, nq] or [bs*heads, nq, nq], bool True = blocked) -> the kernel's
        [bs*heads, nq, nq] 'may attend' mask."""
        if self_attn_mask is None:
            return None
        if self_attn_mask.dtype != torch.bool:
            raise NotImplementedError("float attn_mask")
        m = ~self_attn_mask
        if m.dim() == 2:
            m = m[None].expand(bs * self.n_heads, nq, nq)
        return m.contiguous()

    @torch.no_grad()
    def forward(self, tgt, tgt_query_pos=None, tgt_query_sine_embed=None, tgt_key_padding_mask=None,
                tgt_reference_points=None, memory_text=None, text_attention_mask=None, memory=None,
                memory_key_padding_mask=None, memory_level_start_index=None, memory_spatial_shapes=None,
                memory_pos=None, self_attn_mask=None, cross_attn_mask=None):
        assert cross_attn_mask is None
        x = tgt.transpose(0, 1).contiguous()                                   # [bs, nq, d]
        pos = None if tgt_query_pos is None else tgt_query_pos.transpose(0, 1)
        bs, nq, _ = x.shape
        if self.self_attn is not None:
            qk = self.with_pos_embed(x, pos)
            x = self.norm2(self.self_attn.run(qk, qk, x, attn_mask=self._attend_mask(self_attn_mask, bs, nq), residual=x))
        if self.use_text_cross_attention:
            x = self.catext_norm(self.ca_text.run(self.with_pos_embed(x, pos), memory_text, memory_text,
                                                  key_mask=None if text_attention_mask is None else ~text_attention_mask,
                                                  residual=x))
        attn = self.cross_attn(self.with_pos_embed(x, pos), tgt_reference_points.transpose(0, 1).contiguous(),
                               memory.transpose(0, 1), memory_spatial_shapes, memory_level_start_index,
                               memory_key_padding_mask)
        x = self.forward_ffn(self.norm1(x + attn))
        return x.transpose(0, 1)


# ---------------------------------------------------------------------------------------------------------------------
# The two-stage keypoint decoder (modeling_unipose.py:2869-3130) and its heads
# ---------------------------------------------------------------------------------------------------------------------
import copy   # noqa: E402
import math   # noqa: E402

from .gdino_heads import GroundingDinoMLPPredictionHead as MLP   # noqa: E402  same `layers.N` parameters as utils/model_utils.py:147-160


def inverse_sigmoid(x, eps=1e-3):
    """utils/misc.py:667-671 (the variant modeling_unipose.py imports, eps = 1e-3)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def gen_sineembed_for_position(pos_tensor):
    """utils/model_utils.py:178-204: 128 sine features per coordinate, order (y, x[, w, h]), fp32 arithmetic."""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)

    def emb(c):
        p = (pos_tensor[:, :, c] * scale)[:, :, None] / dim_t
        return torch.stack((p[:, :, 0::2].sin(), p[:, :, 1::2].cos()), dim=3).flatten(2)

    if pos_tensor.size(-1) == 2:
        return torch.cat((emb(1), emb(0)), dim=2)
    if pos_tensor.size(-1) == 4:
        return torch.cat((emb(1), emb(0), emb(2), emb(3)), dim=2)
    raise ValueError("Unknown pos_tensor shape(-1):{}".format(pos_tensor.size(-1)))


class ContrastiveAssign(nn.Module):
    """modeling_unipose.py:947-992: logits = x @ encoded_text^T, -inf on padded text tokens (already max_text_len wide)."""

    def __init__(self, project=False, cal_bias=None, max_text_len=256):
        super().__init__()
        if cal_bias is not None:
            raise NotImplementedError("cal_bias (the reference raises too, :979)")
        self.project, self.cal_bias, self.max_text_len = project, cal_bias, max_text_len

    @torch.no_grad()
    def forward(self, x, text_dict):
        y, mask = text_dict["encoded_text"], text_dict["text_token_mask"]
        B, Q, _ = x.shape
        T = y.shape[1]
        buf = torch.empty((B, Q, (T + 7) // 8 * 8), dtype=x.dtype, device=x.device)     # 16-byte output row pitch
        for b in range(B):
            ops.linear(x[b].contiguous(), y[b].to(x.dtype).contiguous(), out=buf[b, :, :T])
        res = buf[..., :T].masked_fill(~mask[:, None, :], float("-inf"))
        return res.float()                                     # `new_res` is a default-dtype (fp32) buffer (:988-990)


def _get_clones(module, N, layer_share=False):
    return nn.ModuleList([module for _ in range(N)] if layer_share else [copy.deepcopy(module) for _ in range(N)])


class TransformerDecoder(nn.Module):
    """modeling_unipose.py:2869-3130, inference path: `num_box_decoder_layers` box layers over the proposal queries, then
    the 50 best boxes (top-k of the max class logit, indices exact) are expanded into (1 box + num_body_points keypoint)
    queries each and refined by the remaining layers.  Same constructor arguments, parameter names (`layers.N`, `norm`,
    `ref_point_head.layers.N`, `hw`, `hw_append`) and late-bound heads (`bbox_embed`, `class_embed`, `pose_embed`,
    `pose_hw_embed` are assigned by the owner model, :233-239); sequence-first tensors; returns
    `[intermediate (bs-first), reference points (bs-first)]`.  Denoising queries exist only in training (:2985)."""

    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False, d_model=256, query_dim=4,
                 modulate_hw_attn=False, num_feature_levels=1, deformable_decoder=False, decoder_query_perturber=None,
                 dec_layer_number=None, rm_dec_query_scale=False, dec_layer_share=False, dec_layer_dropout_prob=None,
                 use_detached_boxes_dec_out=False, num_box_decoder_layers=2, num_body_points=68):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers, layer_share=dec_layer_share) if num_layers > 0 else []
        self.num_layers, self.norm = num_layers, norm
        assert return_intermediate, "support return_intermediate only"
        assert query_dim in (2, 4)
        self.return_intermediate, self.query_dim = return_intermediate, query_dim
        self.num_feature_levels, self.use_detached_boxes_dec_out = num_feature_levels, use_detached_boxes_dec_out
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.query_pos_sine_scale = None if deformable_decoder else MLP(d_model, d_model, d_model, 2)
        if not rm_dec_query_scale:
            raise NotImplementedError("query_scale (the reference raises too, :2908)")
        self.query_scale = None
        self.bbox_embed = self.class_embed = self.pose_embed = self.pose_hw_embed = None
        self.d_model, self.modulate_hw_attn, self.deformable_decoder = d_model, modulate_hw_attn, deformable_decoder
        self.ref_anchor_head = MLP(d_model, d_model, 2, 2) if (not deformable_decoder and modulate_hw_attn) else None
        if decoder_query_perturber is not None or dec_layer_number is not None or dec_layer_dropout_prob is not None:
            raise NotImplementedError("training-time query perturbation / per-layer query counts / layer dropout")
        self.decoder_query_perturber = self.dec_layer_number = self.dec_layer_dropout_prob = None
        self.rm_detach, self.box_pred_damping = None, None
        self.num_body_points = num_body_points
        self.hw = nn.Embedding(17, 2)
        self.num_box_decoder_layers = num_box_decoder_layers
        self.kpt_index = [x for x in range(50 * (num_body_points + 1)) if x % (num_body_points + 1) != 0]
        self.hw_append = nn.Embedding(num_body_points - 17, 2)

    @torch.no_grad()
    def forward(self, tgt, memory, tgt_mask=None, tgt_mask2=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, refpoints_unsigmoid=None, level_start_index=None,
                spatial_shapes=None, valid_ratios=None, memory_text=None, text_attention_mask=None, text_dict=None,
                dn_meta=None, targets=None, kpt_embed=None):
        if self.training:
            raise NotImplementedError("training (denoising queries) is outside the forward hot path")
        nbp, d = self.num_body_points, self.d_model
        output = tgt
        reference_points = refpoints_unsigmoid.sigmoid()
        intermediate, ref_points = [], [reference_points]
        kpt_index = torch.tensor(self.kpt_index, device=tgt.device)
        new_reference_points = None
        for layer_id, layer in enumerate(self.layers):
            if reference_points.shape[-1] == 4:
                rp_in = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[None, :]
            else:
                rp_in = reference_points[:, :, None] * valid_ratios[None, :]
            query_sine_embed = gen_sineembed_for_position(rp_in[:, :, 0, :]).to(tgt.dtype)
            query_pos = self.ref_point_head(query_sine_embed)
            output = layer(tgt=output, tgt_query_pos=query_pos, tgt_query_sine_embed=query_sine_embed,
                           tgt_key_padding_mask=tgt_key_padding_mask, tgt_reference_points=rp_in.to(tgt.dtype),
                           memory_text=memory_text, text_attention_mask=text_attention_mask, memory=memory,
                           memory_key_padding_mask=memory_key_padding_mask, memory_level_start_index=level_start_index,
                           memory_spatial_shapes=spatial_shapes, memory_pos=pos, self_attn_mask=tgt_mask,
                           cross_attn_mask=memory_mask).contiguous()
            intermediate.append(self.norm(output))
            if layer_id < self.num_box_decoder_layers:                          # box refinement (:3047-3051)
                new_reference_points = (self.bbox_embed[layer_id](output) + inverse_sigmoid(reference_points)).sigmoid()
            if layer_id == self.num_box_decoder_layers - 1:                     # pick 50 boxes, spawn keypoint queries
                cls = self.class_embed[layer_id](output.transpose(0, 1), text_dict).transpose(0, 1)   # [nq, bs, k]
                topk = torch.topk(cls.max(-1)[0], 50, dim=0)[1]                                       # [50, bs]
                if getattr(self, "forced_topk", None) is not None:   # compare runs of different precision on ONE selection
                    topk = self.forced_topk.to(topk.device)
                self.topk_proposals = topk                      # kept for inspection (the reference does not return it)
                box_ref = torch.gather(new_reference_points, 0, topk.unsqueeze(-1).repeat(1, 1, 4))
                box_out = torch.gather(output, 0, topk.unsqueeze(-1).repeat(1, 1, d))
                kpt_out = kpt_embed.transpose(0, 1)[None].repeat(box_out.shape[0], 1, 1, 1).to(output.dtype)   # [50, nkpt, bs, d]
                delta_xy = self.pose_embed[-1](kpt_out)[..., :2]
                kpt_xy = (inverse_sigmoid(box_ref[..., :2][:, None]) + delta_xy).sigmoid()
                nq2, _, bs, _ = kpt_xy.shape
                hw = torch.cat((self.hw.weight, self.hw_append.weight), dim=0)
                kpt_wh = hw.unsqueeze(0).unsqueeze(-2).repeat(nq2, 1, bs, 1).sigmoid() * box_ref[..., 2:][:, None]
                kpt_ref = torch.cat((kpt_xy, kpt_wh.to(kpt_xy.dtype)), dim=-1)
                new_reference_points = torch.cat((box_ref.unsqueeze(1), kpt_ref), dim=1).flatten(0, 1)
                output = torch.cat((box_out.unsqueeze(1), kpt_out), dim=1).flatten(0, 1)
                tgt_mask = tgt_mask2
            if layer_id >= self.num_box_decoder_layers:                         # box + keypoint refinement (:3103-3131)
                before = inverse_sigmoid(reference_points)
                j = layer_id - self.num_box_decoder_layers
                box_new = (self.bbox_embed[layer_id](output[0::(nbp + 1)].contiguous()) + before[0::(nbp + 1)]).sigmoid()
                out_kpt = output.index_select(0, kpt_index)
                unsig = before.index_select(0, kpt_index).clone()
                unsig[..., :2] += self.pose_embed[j](out_kpt)[..., :2]
                unsig[..., 2:] += self.pose_hw_embed[j](out_kpt)
                bs = box_new.shape[1]
                new_reference_points = torch.cat((box_new.unsqueeze(1), unsig.sigmoid().view(-1, nbp, bs, 4)),
                                                 dim=1).flatten(0, 1)
            reference_points = new_reference_points.detach()
            ref_points.append(reference_points if self.use_detached_boxes_dec_out else new_reference_points)
        return [[o.transpose(0, 1) for o in intermediate], [r.transpose(0, 1) for r in ref_points]]


def prepare_for_mask(kpt_mask, nheads, num_body_points, num_group=50):
    """UniPose.prepare_for_mask (modeling_unipose.py:887-917), the inference-time self-attention mask of the expanded
    (box + keypoint) queries, vectorised and on kpt_mask's device (the reference hard-codes 'cuda' and loops over
    50 * num_body_points rows).  kpt_mask [bs, 1 + num_body_points], 1 = the box slot / a real keypoint class.  Returns
    attn_mask2 [bs * nheads, T, T] bool, True = blocked, T = num_group * (1 + num_body_points).  Bug-compatible: the
    first pass groups rows by num_body_points (not 1 + num_body_points), so outside the diagonal blocks the pattern is
    the reference's, stripe for stripe."""
    bs, length = kpt_mask.shape
    dev = kpt_mask.device
    T = num_group * (1 + num_body_points)
    gb = num_body_points
    rows = torch.arange(T, device=dev)
    cols = torch.arange(T, device=dev)
    sj = (rows // gb) * gb
    ej = sj + gb
    touched = rows < num_group * gb
    left = (cols[None, :] < sj[:, None]) & (sj[:, None] > 0)
    right = (cols[None, :] >= ej[:, None]) & (ej[:, None] < num_group * gb)
    base = (left | right) & touched[:, None]                                    # [T, T]
    m = base[None].repeat(bs, 1, 1)
    equal = kpt_mask[:, :, None] == kpt_mask[:, None, :]                        # [bs, length, length]
    for idx in range(num_group):                                                # 50 diagonal blocks
        s0 = idx * length
        m[:, s0:s0 + length, s0:s0 + length] = ~equal
    return m[:, None].repeat(1, nheads, 1, 1).flatten(0, 1)

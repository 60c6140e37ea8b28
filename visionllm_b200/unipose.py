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

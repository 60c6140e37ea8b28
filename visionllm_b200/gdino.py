"""B200-native Grounding-DINO deformable blocks (the region/mask decoder's hot layers).

Same class names, constructor arguments, parameter names and call signatures as
visionllmv2/model/grounding_dino/modeling_ov_grounding_dino_mask_dn.py:
  GroundingDinoMultiscaleDeformableAttention  :646-784   (sampling_offsets / attention_weights / value_proj /
                                                           output_proj -> MSDA gather)
  GroundingDinoDeformableLayer                :1104-1182 (encoder: MSDA + LN + FFN + LN)
  GroundingDinoDecoderLayer                   :1292-1407 (self-MHA, text cross-MHA, MSDA cross-attn, FFN)
so reference state dicts load unchanged and the modules can be assigned over the reference classes
(INTEGRATION.md).  Projections run on the tcgen05 GEMM (offset and weight projections share one launch),
LayerNorm/residuals on the row kernels, attention on the fused attention kernel, the gather on
vllm_msda_forward_f32 -- never the grid_sample fallback the reference silently drops to (:777-779).
Forward / inference only.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import msda as msda_ext
from . import ops


class GroundingDinoMultiscaleDeformableAttention(nn.Module):
    def __init__(self, config, num_heads, n_points):
        super().__init__()
        if config.d_model % num_heads != 0:
            raise ValueError(f"embed_dim (d_model) must be divisible by num_heads, but got {config.d_model} "
                             f"and {num_heads}")
        self.im2col_step = 64
        self.d_model = config.d_model
        self.n_levels = config.num_feature_levels
        self.n_heads = num_heads
        self.n_points = n_points
        self.sampling_offsets = nn.Linear(config.d_model, num_heads * self.n_levels * n_points * 2)
        self.attention_weights = nn.Linear(config.d_model, num_heads * self.n_levels * n_points)
        self.value_proj = nn.Linear(config.d_model, config.d_model)
        self.output_proj = nn.Linear(config.d_model, config.d_model)
        self._packed = None
        self._reset_parameters()

    def _reset_parameters(self):
        # same init as the reference (:688-706)
        nn.init.constant_(self.sampling_offsets.weight.data, 0.0)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(
            1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.0)
        nn.init.constant_(self.attention_weights.bias.data, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.0)

    def _packed_query_proj(self, dtype):
        ws = (self.sampling_offsets.weight, self.attention_weights.weight, self.sampling_offsets.bias,
              self.attention_weights.bias)
        key = tuple((w.data_ptr(), w._version) for w in ws) + (dtype,)
        if self._packed is None or self._packed[0] != key:
            w = torch.cat([ws[0].detach(), ws[1].detach()], 0).to(dtype).contiguous()
            b = torch.cat([ws[2].detach(), ws[3].detach()], 0).to(dtype).contiguous()
            self._packed = (key, w, b)
        return self._packed[1], self._packed[2]

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                position_embeddings=None, reference_points=None, spatial_shapes=None, level_start_index=None,
                output_attentions=False):
        if position_embeddings is not None:
            hidden_states = hidden_states + position_embeddings
        B, Lq, _ = hidden_states.shape
        _, S, _ = encoder_hidden_states.shape
        if (spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum() != S:
            raise ValueError("Make sure to align the spatial shapes with the sequence length of the encoder "
                             "hidden states")
        M, L, P, D = self.n_heads, self.n_levels, self.n_points, self.d_model // self.n_heads
        value = ops.linear(encoder_hidden_states, self.value_proj.weight, bias=self.value_proj.bias)
        if attention_mask is not None:
            value = value.masked_fill(~attention_mask[..., None], float(0))
        w, b = self._packed_query_proj(hidden_states.dtype)
        qp = ops.linear(hidden_states, w, bias=b)                    # offsets | weights in one GEMM
        n_off = M * L * P * 2
        sampling_offsets = qp[..., :n_off].reshape(B, Lq, M, L, P, 2)
        attention_weights = F.softmax(qp[..., n_off:].reshape(B, Lq, M, L * P), -1).view(B, Lq, M, L, P)
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + sampling_offsets / normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = (reference_points[:, :, None, :, None, :2]
                   + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but got {reference_points.shape[-1]}")
        # the reference upcasts value / weights to fp32 for the kernel (:764-766); locations follow type promotion
        out = msda_ext.ms_deform_attn_forward(value.view(B, S, M, D).float().contiguous(), spatial_shapes,
                                              level_start_index, loc.float().contiguous(),
                                              attention_weights.float().contiguous(), self.im2col_step)
        out = ops.linear(out.to(self.output_proj.weight.dtype), self.output_proj.weight, bias=self.output_proj.bias)
        return out, attention_weights


def _act(config):
    a = getattr(config, "activation_function", "relu")
    if a not in ("relu", "gelu", "silu"):
        raise NotImplementedError(f"activation_function={a}")
    return a


class _LN(nn.LayerNorm):
    def forward(self, x):
        return ops.layernorm(x, self.weight, self.bias, self.eps)


class GroundingDinoDeformableLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = GroundingDinoMultiscaleDeformableAttention(
            config, num_heads=config.encoder_attention_heads, n_points=config.encoder_n_points)
        self.self_attn_layer_norm = _LN(self.embed_dim)
        self.act = _act(config)
        self.fc1 = nn.Linear(self.embed_dim, config.encoder_ffn_dim)
        self.fc2 = nn.Linear(config.encoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = _LN(self.embed_dim)

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask, position_embeddings=None, reference_points=None,
                spatial_shapes=None, level_start_index=None, output_attentions=False):
        attn, w = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask,
                                 encoder_hidden_states=hidden_states, encoder_attention_mask=attention_mask,
                                 position_embeddings=position_embeddings, reference_points=reference_points,
                                 spatial_shapes=spatial_shapes, level_start_index=level_start_index)
        x = self.self_attn_layer_norm(hidden_states + attn)
        h = ops.linear(x, self.fc1.weight, bias=self.fc1.bias, act=self.act)
        x = self.final_layer_norm(ops.linear(h, self.fc2.weight, bias=self.fc2.bias, residual=x))
        return x, w


class _MHA(nn.MultiheadAttention):
    """nn.MultiheadAttention parameters (in_proj_weight/in_proj_bias/out_proj) with a kernel forward."""

    def run(self, query, key, value, key_lengths=None, residual=None):
        E, H = self.embed_dim, self.num_heads
        B, Tq, _ = query.shape
        Tk = key.shape[1]
        w, b = self.in_proj_weight, self.in_proj_bias
        if query is key:
            qk = ops.linear(query, w[:2 * E], bias=b[:2 * E])
            q, k = qk[..., :E], qk[..., E:]
        else:
            q = ops.linear(query, w[:E], bias=b[:E])
            k = ops.linear(key, w[E:2 * E], bias=b[E:2 * E])
        v = ops.linear(value, w[2 * E:], bias=b[2 * E:])
        ctx = ops.attention(q.unflatten(-1, (H, E // H)), k.unflatten(-1, (H, E // H)),
                            v.unflatten(-1, (H, E // H)), causal=False, seqlens=key_lengths)
        return ops.linear(ctx, self.out_proj.weight, bias=self.out_proj.bias, residual=residual)


def _prefix_lengths(pad_mask):
    """key_padding_mask (True = ignore) -> int32 lengths; only right padding is expressible."""
    if pad_mask is None:
        return None
    keep = ~pad_mask
    lens = keep.sum(-1).to(torch.int32)
    T = pad_mask.shape[1]
    if not bool((keep == (torch.arange(T, device=pad_mask.device)[None] < lens[:, None])).all()):
        raise NotImplementedError("text_encoder_attention_mask must mask a suffix (it does on the reference path: "
                                  "text_query_masks marks the first num_patches entries, mv2.py:779-786)")
    return lens


class GroundingDinoDecoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embed_dim = config.d_model
        self.self_attn = _MHA(self.embed_dim, config.decoder_attention_heads, dropout=config.attention_dropout,
                              batch_first=True)
        self.self_attn_layer_norm = _LN(self.embed_dim)
        self.encoder_attn_text = _MHA(self.embed_dim, config.decoder_attention_heads,
                                      dropout=config.attention_dropout, batch_first=True)
        self.encoder_attn_text_layer_norm = _LN(self.embed_dim)
        self.encoder_attn = GroundingDinoMultiscaleDeformableAttention(
            config, num_heads=config.decoder_attention_heads, n_points=config.decoder_n_points)
        self.encoder_attn_layer_norm = _LN(self.embed_dim)
        self.act = _act(config)
        self.fc1 = nn.Linear(self.embed_dim, config.decoder_ffn_dim)
        self.fc2 = nn.Linear(config.decoder_ffn_dim, self.embed_dim)
        self.final_layer_norm = _LN(self.embed_dim)

    @torch.no_grad()
    def forward(self, hidden_states, position_embeddings=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, vision_encoder_hidden_states=None, vision_encoder_attention_mask=None,
                text_encoder_hidden_states=None, text_encoder_attention_mask=None, self_attn_mask=None,
                output_attentions=False):
        if self_attn_mask is not None:
            raise NotImplementedError("self_attn_mask is only used by contrastive-DN training (:2659-2829)")
        x = hidden_states
        pos = position_embeddings
        qk = x if pos is None else x + pos
        x = self.self_attn_layer_norm(self.self_attn.run(qk, qk, x, residual=x))
        q = x if pos is None else x + pos
        x = self.encoder_attn_text_layer_norm(
            self.encoder_attn_text.run(q, text_encoder_hidden_states, text_encoder_hidden_states,
                                       key_lengths=_prefix_lengths(text_encoder_attention_mask), residual=x))
        attn, _ = self.encoder_attn(hidden_states=x, attention_mask=vision_encoder_attention_mask,
                                    encoder_hidden_states=vision_encoder_hidden_states,
                                    encoder_attention_mask=vision_encoder_attention_mask, position_embeddings=pos,
                                    reference_points=reference_points, spatial_shapes=spatial_shapes,
                                    level_start_index=level_start_index)
        x = self.encoder_attn_layer_norm(x + attn)
        h = ops.linear(x, self.fc1.weight, bias=self.fc1.bias, act=self.act)
        x = self.final_layer_norm(ops.linear(h, self.fc2.weight, bias=self.fc2.bias, residual=x))
        return (x,)

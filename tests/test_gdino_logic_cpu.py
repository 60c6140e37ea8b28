"""CPU: HOST LOGIC of the GDINO modules (masks, packing, residual/LayerScale wiring, bug-compatible mask
expansion) against the reference-generated goldens, with the CUDA kernels replaced -- in this test only -- by
fp32 torch stand-ins of the same op contracts.  This exercises no product compute path (that is what the -m gpu
tests do); it pins the Python around it so a logic error shows up without a GPU."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from weights_util import key_shapes, seeded_state_dict  # noqa: E402


@pytest.fixture()
def torch_kernels(monkeypatch):
    import visionllm_b200.msda as msda
    import visionllm_b200.ops as ops
    from oracle import msda_oracle as O

    def linear(x, w, bias=None, act=None, colscale=None, residual=None, out_dtype=None, out=None, row_keep=None):
        y = F.linear(x.float(), w.float(), None if bias is None else bias.float())
        y = {"relu": torch.relu, "gelu": F.gelu, "silu": F.silu, None: lambda z: z}[act](y)
        if colscale is not None:
            y = y * colscale.float()
        if residual is not None:
            y = y + residual.float().reshape(y.shape)
        if row_keep is not None:
            y = y.masked_fill(~row_keep.bool().reshape(y.shape[:-1])[..., None], 0.0)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def attention(q, k, v, causal=False, scale=None, seqlens=None, key_mask=None, attn_mask=None, attn_bias=None, out=None):
        B, Tq, H, D = q.shape
        Tk = k.shape[1]
        s = (q.float().permute(0, 2, 1, 3) @ k.float().permute(0, 2, 3, 1)) * (scale or D ** -0.5)
        if attn_bias is not None:
            s = s + attn_bias[torch.arange(B) % attn_bias.shape[0]]
        if attn_mask is not None:
            s = s.masked_fill(~attn_mask.bool().view(B, H, Tq, Tk), float("-inf"))
        if key_mask is not None:
            s = s.masked_fill(~key_mask.bool()[:, None, None, :], float("-inf"))
        if seqlens is not None:
            s = s.masked_fill(torch.arange(Tk)[None, None, None, :] >= seqlens[:, None, None, None], float("-inf"))
        return (torch.softmax(s, -1) @ v.float().permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Tq, H * D)

    monkeypatch.setattr(ops, "linear", linear)
    monkeypatch.setattr(ops, "layernorm", lambda x, w, b, eps, out=None: F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps))
    monkeypatch.setattr(ops, "attention", attention)

    def layernorm_gather(x, index, w, b, eps):            # LN + zero pad + row gather (Swin window partition)
        h = F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)
        h = torch.cat((h, h.new_zeros(h.shape[0], 1, h.shape[2])), 1)
        return h.index_select(1, index.clamp(max=x.shape[1]))

    monkeypatch.setattr(ops, "layernorm_gather", layernorm_gather)
    monkeypatch.setattr(ops, "gather_rows", lambda src, idx: src[idx])      # Swin window reverse (one row gather)
    monkeypatch.setattr(msda, "ms_deform_attn_forward",
                        lambda value, shapes, lsi, loc, w, step, **kw: O.forward_grid_sample(value, shapes, loc, w))


def cfg():
    return SimpleNamespace(d_model=256, encoder_attention_heads=8, decoder_attention_heads=8, encoder_ffn_dim=512,
                           decoder_ffn_dim=512, num_feature_levels=4, encoder_n_points=4, decoder_n_points=4,
                           dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, activation_function="relu")


def test_encoder_layer_logic(golden_dir, torch_kernels):
    from visionllm_b200.gdino import GroundingDinoEncoderLayer
    g = np.load(os.path.join(golden_dir, "mod_gdino_encoder_layer.npz"))
    m = GroundingDinoEncoderLayer(cfg())
    assert json.loads(str(g["keys"])) == [list(k) for k in key_shapes(m)]
    sd = seeded_state_dict(m, 505)
    for k in sd:
        if k.endswith("vision_param") or k.endswith("text_param"):
            sd[k] = sd[k] * 0 + 0.5
    m.load_state_dict(sd)
    m.eval()
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(dt)  # noqa: E731
    tq = t("tq_mask", torch.bool)
    (v, txt), _ = m(vision_features=t("src"), vision_position_embedding=t("pos"), spatial_shapes=t("shapes", torch.int64),
                    level_start_index=t("lsi", torch.int64), key_padding_mask=t("kpm", torch.bool),
                    reference_points=t("ref"), text_features=t("text"), text_attention_mask=~tq,
                    text_position_embedding=None, text_self_attention_masks=t("tsa", torch.bool),
                    text_position_ids=t("pids", torch.int64))
    ref = t("out_f32")
    B, S, C = v.shape
    assert (v - ref[:, :S * C].view(B, S, C)).abs().max() < 1e-4
    # every text row, padded ones included: the batch-mixing of the reference's mask expansion is reproduced
    assert (txt - ref[:, S * C:].view(B, -1, C)).abs().max() < 1e-4


def test_decoder_and_deformable_layer_logic(golden_dir, torch_kernels):
    from visionllm_b200.gdino import GroundingDinoDecoderLayer, GroundingDinoDeformableLayer
    g = np.load(os.path.join(golden_dir, "mod_gdino_deformable_layer.npz"))
    m = GroundingDinoDeformableLayer(cfg()); m.load_state_dict(seeded_state_dict(m, 202)); m.eval()
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(dt)  # noqa: E731
    out, _ = m(t("src"), t("mask", torch.bool), position_embeddings=t("pos"), reference_points=t("ref"),
               spatial_shapes=t("shapes", torch.int64), level_start_index=t("lsi", torch.int64))
    assert (out - t("out_f32")).abs().max() < 1e-4
    g = np.load(os.path.join(golden_dir, "mod_gdino_decoder_layer.npz"))
    d = GroundingDinoDecoderLayer(cfg()); d.load_state_dict(seeded_state_dict(d, 303)); d.eval()
    (out,) = d(t("hs"), position_embeddings=t("qpos"), reference_points=t("ref"), spatial_shapes=t("shapes", torch.int64),
               level_start_index=t("lsi", torch.int64), vision_encoder_hidden_states=t("src"),
               vision_encoder_attention_mask=t("mask", torch.bool), text_encoder_hidden_states=t("text"),
               text_encoder_attention_mask=t("tpad", torch.bool))
    assert (out - t("out_f32")).abs().max() < 1e-4

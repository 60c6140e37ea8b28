"""CPU: host logic of the B200 UniPose deformable layers (visionllm_b200/unipose.py) against the reference's own
classes (tests/golden/mod_unipose_layers.npz from gen_golden_unipose.py): parameter names, `True = padding` mask
conventions, sequence-first decoder tensors, nn.MultiheadAttention attn_mask / key_padding_mask semantics, 2-d and 4-d
reference points.  Kernels replaced IN THIS TEST ONLY by the fp32 stand-ins of test_gdino_logic_cpu."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from test_gdino_logic_cpu import torch_kernels  # noqa: E402,F401
from unipose_inputs import inputs  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402


def build():
    from visionllm_b200.unipose import DeformableTransformerDecoderLayer, DeformableTransformerEncoderLayer
    enc = DeformableTransformerEncoderLayer(256, 512, 0.0, "relu", 4, 8, 4).eval()
    enc.load_state_dict(seeded_state_dict(enc, 41))
    dec = DeformableTransformerDecoderLayer(256, 512, 0.0, "relu", 4, 8, 4, use_text_cross_attention=True).eval()
    dec.load_state_dict(seeded_state_dict(dec, 42))
    return enc, dec


def run(enc, dec, x, c=lambda t: t):
    e = enc(c(x["src"]), c(x["pos"]), c(x["ref2"]), x["shapes"], x["lsi"], x["pad"])
    d = dec(tgt=c(x["tgt"]), tgt_query_pos=c(x["qpos"]), tgt_reference_points=c(x["ref4"]),
            memory_text=c(x["memory_text"]), text_attention_mask=x["text_mask"], memory=c(x["memory"]),
            memory_key_padding_mask=x["pad"], memory_level_start_index=x["lsi"], memory_spatial_shapes=x["shapes"],
            self_attn_mask=x["attn_mask"])
    return e, d


def test_unipose_layers_logic_matches_reference(golden_dir, torch_kernels):  # noqa: F811
    g = np.load(os.path.join(golden_dir, "mod_unipose_layers.npz"))
    enc, dec = build()
    assert json.loads(str(g["enc_keys"])) == [list(k) for k in key_shapes(enc)], "encoder-layer keys differ"
    assert json.loads(str(g["dec_keys"])) == [list(k) for k in key_shapes(dec)], "decoder-layer keys differ"
    e, d = run(enc, dec, inputs())
    sub = int(g["sub"])
    for got, ref in ((e[:, ::sub], torch.from_numpy(g["enc_f32"])), (d, torch.from_numpy(g["dec_f32"]))):
        assert got.shape == ref.shape
        assert (got.float() - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())

"""Full-WIDTH layer parity: the q/k-norm-over-3200, LayerScale-epilogue and long-T RoPE
compositions at the real InternViT-6B / Vicuna-7B widths, not only at the toy sizes of test_modules_gpu.py.

* InternViT-6B layer (3200 / 25 heads / MLP 12800 / 1025 tokens): golden from the reference's OWN
  `InternVisionEncoderLayer` (tests/golden/gen_golden_fullwidth.py; weights + input regenerated from seeds).
* Vicuna-7B layer (4096 / 32 heads / MLP 11008 / T = 1536, causal): the installed HF `LlamaForCausalLM` (the reference's
  LLM is third-party transformers) run here in fp32 and in bf16 as the oracle.

Tolerance = the module rule of test_modules_gpu.py: rel_l2(ours, ref_fp32) <= 1.5 * rel_l2(ref_bf16, ref_fp32) + 1e-3.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from weights_util import key_shapes, seeded_state_dict  # noqa: E402


def rel_l2(a, b):
    return float(torch.linalg.norm(a.float() - b.float()) / torch.linalg.norm(b.float()))


def test_internvit6b_layer_full_width_matches_reference(golden_dir):
    from visionllm_b200.internvit import InternVisionConfig, InternVisionEncoderLayer
    g = np.load(os.path.join(golden_dir, "mod_internvit6b_layer.npz"))
    cfg = InternVisionConfig(hidden_size=3200, num_attention_heads=25, num_hidden_layers=1, intermediate_size=12800,
                             image_size=448, patch_size=14, qk_normalization=True, qkv_bias=False)
    layer = InternVisionEncoderLayer(cfg)
    assert json.loads(str(g["keys"])) == [list(k) for k in key_shapes(layer)], "state-dict keys differ from reference"
    seed_w, seed_x = (int(v) for v in g["seeds"])
    layer.load_state_dict(seeded_state_dict(layer, seed_w))
    layer = layer.to("cuda", torch.bfloat16).eval()
    x = torch.randn(1, 1025, 3200, generator=torch.Generator().manual_seed(seed_x)).to(torch.bfloat16)
    with torch.no_grad():
        out = layer(x.cuda())
    out = out.reshape(1025, 3200)[torch.from_numpy(g["rows"]).cuda()]
    ref32 = torch.from_numpy(g["out_f32"]).cuda()
    ref16 = torch.from_numpy(g["out_refbf16"]).cuda()
    budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
    got = rel_l2(out, ref32)
    assert got <= budget, f"rel_l2 {got:.3e} > budget {budget:.3e}"
    assert (out.float() - ref32).abs().max() <= 3 * (ref16 - ref32).abs().max() + 1e-2


def test_vicuna7b_layer_full_width_matches_hf():
    from transformers import LlamaConfig, LlamaForCausalLM
    from visionllm_b200.llama import B200LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=1, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=1024, rms_norm_eps=1e-5, max_position_embeddings=4096,
                      attn_implementation="eager")
    torch.manual_seed(0)
    hf = LlamaForCausalLM(cfg).eval()
    sd = {k: v.to(torch.bfloat16).float() for k, v in hf.state_dict().items()}
    hf.load_state_dict(sd)
    mine = B200LlamaForCausalLM(cfg)
    mine.load_state_dict(sd, strict=True)
    mine = mine.to("cuda", torch.bfloat16).eval()
    B, T = 1, 1536
    emb = (torch.randn(B, T, 4096, generator=torch.Generator().manual_seed(1)) * 0.5).bfloat16().cuda()
    am = torch.ones(B, T, dtype=torch.long, device="cuda")
    with torch.no_grad():
        ref32 = hf.float().cuda()(inputs_embeds=emb.float(), attention_mask=am, output_hidden_states=True)
        r32 = (ref32.hidden_states[1].clone(), ref32.logits.float().clone())
        del ref32
        ref16 = hf.bfloat16()(inputs_embeds=emb, attention_mask=am, output_hidden_states=True)
        r16 = (ref16.hidden_states[1].float(), ref16.logits.float())
    out = mine(inputs_embeds=emb, attention_mask=am, output_hidden_states=True)
    for a, b32, b16, what in ((out.hidden_states[1], r32[0], r16[0], "layer output"),
                              (out.logits, r32[1], r16[1], "logits")):
        assert rel_l2(a, b32) <= 1.5 * rel_l2(b16, b32) + 1e-3, (what, rel_l2(a, b32), rel_l2(b16, b32))
    # long-T RoPE: the tail of the sequence is as accurate as its head
    tail = slice(T - 128, T)
    assert rel_l2(out.hidden_states[1][:, tail], r32[0][:, tail]) <= 1.5 * rel_l2(r16[0][:, tail], r32[0][:, tail]) + 1e-3

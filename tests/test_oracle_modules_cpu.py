"""CPU: pin the fp32 torch oracle of the ViT / LLM layers (oracle/vit_llm_oracle.py)."""
import os
import sys

import numpy as np
import torch

from oracle import vit_llm_oracle as VO

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from weights_util import seeded_state_dict  # noqa: E402


def test_internvit_oracle_matches_reference_golden(golden_dir):
    from visionllm_b200.internvit import B200InternVisionModel, InternVisionConfig
    g = np.load(os.path.join(golden_dir, "mod_internvit_small.npz"))
    cfg = InternVisionConfig(hidden_size=256, num_attention_heads=2, num_hidden_layers=2, intermediate_size=512,
                             image_size=56, patch_size=14)
    sd = seeded_state_dict(B200InternVisionModel(cfg), 101)       # same keys/shapes as the reference (recorded in the golden)
    states = VO.internvit_forward(torch.from_numpy(g["pixel_values"]), sd, layers=2, heads=2, patch=14)
    out = torch.stack([states[-1], states[-2], states[0]]).numpy()
    assert np.abs(out - g["out_f32"]).max() < 2e-5


def test_llama_layer_oracle_matches_hf():
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=1, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=64, rms_norm_eps=1e-5, attn_implementation="eager")
    torch.manual_seed(0)
    hf = LlamaForCausalLM(cfg).eval()
    x = torch.randn(2, 19, 128)
    with torch.no_grad():
        ref = hf.model(inputs_embeds=x, output_hidden_states=True).hidden_states[1]
    sd = {k[len("model."):]: v for k, v in hf.state_dict().items() if k.startswith("model.")}
    out = VO.llama_layer(x, sd, "layers.0.", heads=4, eps=1e-5)
    out = VO.rmsnorm(out, sd["norm.weight"], 1e-5)          # HF's last hidden state is post final-norm
    assert (out - ref).abs().max() < 2e-5

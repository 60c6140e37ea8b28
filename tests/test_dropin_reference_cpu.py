"""CPU, build container only (needs /root/reference): the drop-in claim of INTEGRATION.md section 4 -- assign the
B200 layer classes over the reference's names, let the REFERENCE'S OWN `GroundingDinoEncoder` stack (its python
loop, reference points, hidden-state bookkeeping) drive them, and compare with the untouched reference stack.
Kernels are replaced by fp32 torch stand-ins here (no GPU in this container); the -m gpu tests cover the kernels."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/VisionLLMv2"), reason="reference tree not mounted")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from test_gdino_logic_cpu import torch_kernels  # noqa: E402,F401  (fixture)


def test_reference_encoder_stack_runs_on_b200_layers(torch_kernels):  # noqa: F811
    import ref_shim
    from weights_util import seeded_state_dict
    import visionllm_b200.gdino as b200
    cfgm, gd = ref_shim.load_gdino()
    cfg = cfgm.GroundingDinoConfig(d_model=256, encoder_layers=2, encoder_attention_heads=8, encoder_ffn_dim=512,
                                   num_feature_levels=4, encoder_n_points=4, dropout=0.0, attention_dropout=0.0,
                                   activation_dropout=0.0, fusion_dropout=0.0, fusion_droppath=0.0,
                                   text_enhancer_dropout=0.0, disable_custom_kernels=True)
    ref_enc = gd.GroundingDinoEncoder(cfg).eval()
    sd = seeded_state_dict(ref_enc, 77)
    ref_enc.load_state_dict(sd)
    saved = gd.GroundingDinoEncoderLayer
    try:
        gd.GroundingDinoEncoderLayer = b200.GroundingDinoEncoderLayer          # the one-line swap
        cfg.activation_function = "relu"
        new_enc = gd.GroundingDinoEncoder(cfg).eval()
    finally:
        gd.GroundingDinoEncoderLayer = saved
    assert type(new_enc.layers[0]).__module__ == "visionllm_b200.gdino"
    new_enc.load_state_dict(sd, strict=True)                                    # identical parameter names
    shapes_l = [(8, 10), (4, 5), (2, 3), (1, 2)]
    shapes = torch.tensor(shapes_l)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S, B, T = int(shapes.prod(1).sum()), 1, 6
    g = torch.Generator().manual_seed(1)
    src, pos, text = (torch.randn(B, S, 256, generator=g), torch.randn(B, S, 256, generator=g) * 0.5,
                      torch.randn(B, T, 256, generator=g))
    vmask = torch.zeros(B, S, dtype=torch.bool)
    tq = torch.ones(B, T, dtype=torch.bool); tq[0, 4:] = False
    tsa, pids = gd.generate_masks_with_text_query_masks(tq)
    kw = dict(vision_features=src, vision_attention_mask=vmask, vision_position_embedding=pos, spatial_shapes=shapes,
              level_start_index=lsi, valid_ratios=torch.ones(B, 4, 2), text_features=text, text_attention_mask=~tq,
              text_position_embedding=None, text_self_attention_masks=tsa, text_position_ids=pids,
              output_attentions=False, output_hidden_states=False, return_dict=True)
    with torch.no_grad():
        a = ref_enc(**kw)
        b = new_enc(**kw)
    assert (a.last_hidden_state_vision - b.last_hidden_state_vision).abs().max() < 1e-4
    assert (a.last_hidden_state_text - b.last_hidden_state_text).abs().max() < 1e-4

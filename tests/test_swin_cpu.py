"""CPU: host logic of visionllm_b200.swin.B200SwinBackbone (window-row indices, shift masks, bias slabs, patch merging,
stage bookkeeping, HF state-dict compatibility) against HF `SwinBackbone` -- the third-party code the reference
instantiates through AutoBackbone (grounding_dino/modeling_ov_grounding_dino_mask_dn.py:483) -- in fp32, with the
CUDA kernels replaced in this test only by torch stand-ins."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from test_gdino_logic_cpu import torch_kernels  # noqa: E402,F401  (fixture)


@pytest.mark.parametrize("hw,ws,depths", [((64, 64), 4, [2, 2, 2, 2]), ((56, 72), 7, [2, 2, 2, 2]), ((50, 38), 4, [1, 2, 1, 1])])
def test_swin_backbone_matches_hf(torch_kernels, hw, ws, depths):  # noqa: F811
    from transformers import SwinConfig
    from transformers.models.swin.modeling_swin import SwinBackbone
    from weights_util import seeded_state_dict
    from visionllm_b200.swin import B200SwinBackbone
    cfg = SwinConfig(image_size=64, embed_dim=32, depths=depths, num_heads=[1, 2, 4, 8], window_size=ws,
                     out_features=["stage1", "stage2", "stage3", "stage4"])
    hf = SwinBackbone(cfg).eval()
    sd = seeded_state_dict(hf, 7)
    hf.load_state_dict(sd)
    ours = B200SwinBackbone(cfg).eval()
    assert sorted(ours.state_dict().keys()) == sorted(hf.state_dict().keys())
    ours.load_state_dict(sd, strict=True)
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a = hf(x).feature_maps
        b = ours(x, nchw=True).feature_maps
        c = ours(x).feature_maps
    assert len(a) == len(b) == 4
    for fa, fb, fc in zip(a, b, c):
        assert fa.shape == fb.shape
        assert (fa - fb).abs().max() < 2e-4 * max(1.0, fa.abs().max().item())
        assert getattr(fc, "_b200_nhwc", False) and torch.equal(fc.permute(0, 3, 1, 2), fb)


def test_window_rows_are_a_permutation_with_padding():
    from visionllm_b200.swin import _window_rows
    for (H, W, ws, shift) in [(8, 8, 4, 0), (8, 8, 4, 2), (10, 7, 4, 2), (37, 50, 7, 3)]:
        fwd, inv, Hp, Wp = _window_rows(H, W, ws, shift, "cpu")
        assert fwd.dtype == torch.int64 and fwd.numel() == Hp * Wp and inv.numel() == H * W
        real = fwd[fwd < H * W]
        assert torch.equal(torch.sort(real)[0], torch.arange(H * W))          # every token lands in exactly one slot
        assert torch.equal(fwd[inv], torch.arange(H * W))                     # and inv finds it again

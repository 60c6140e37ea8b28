"""GPU: the region encoder on our kernels against the reference's own `RegionEncoder` run
(tests/golden/mod_region_encoder.npz), module tolerance rule of test_modules_gpu.py; the 'grid_sample' mode is fed the
reference's recorded point draw and pools through the MSDA kernel."""
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


@pytest.mark.parametrize("mode", ["mean", "cross_attn", "grid_sample"])
def test_region_encoder_matches_reference(golden_dir, mode):
    from visionllm_b200.region_encoder import B200RegionEncoder
    g = np.load(os.path.join(golden_dir, "mod_region_encoder.npz"))
    cfg = json.loads(str(g["cfg"]))
    m = B200RegionEncoder(mask_pool_type=mode, **cfg)
    assert json.loads(str(g[f"keys_{mode}"])) == [list(k) for k in key_shapes(m)]
    m.load_state_dict(seeded_state_dict(m, 77))
    m = m.to("cuda", torch.bfloat16).eval()
    feats = [torch.from_numpy(g[f"feat_{i}"]).cuda().bfloat16() for i in range(3)]
    if mode == "cross_attn":
        feats = feats[:1]
    B = g["images"].shape[0]
    pts = None
    if mode == "grid_sample":
        pts = [[torch.from_numpy(g[f"points_{lv}_{i}"]).cuda() for i in range(B)] for lv in range(3)]
    out = m(torch.from_numpy(g["images"]).cuda().bfloat16(), torch.from_numpy(g["masks"]).cuda().bfloat16(), feats,
            sample_points=pts)
    ref32 = torch.from_numpy(g[f"out_f32_{mode}"]).cuda()
    ref16 = torch.from_numpy(g[f"out_refbf16_{mode}"]).cuda()
    assert out.shape == ref32.shape and out.dtype == torch.bfloat16
    budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
    assert rel_l2(out, ref32) <= budget, (rel_l2(out, ref32), budget)


def test_region_encoder_draws_its_own_points():
    """Production path: no points given -> the restated sampler runs on the device; output finite, right shape."""
    from visionllm_b200.region_encoder import B200RegionEncoder
    torch.manual_seed(0)
    m = B200RegionEncoder(hidden_dim=64, embed_dim=256, out_dim=96, mask_pool_type="grid_sample").to("cuda", torch.bfloat16).eval()
    images = torch.randn(2, 3, 112, 112, device="cuda").bfloat16()
    masks = torch.zeros(2, 1, 112, 112, device="cuda", dtype=torch.bfloat16)
    masks[0, 0, 20:90, 10:100] = 1                       # 6300 pixels > 2304 points: the cap applies
    feats = [torch.randn(2, 64, 256, device="cuda").bfloat16() for _ in range(2)]
    out = m(images, masks, feats)                        # region 1 is empty: nan_to_num path
    assert out.shape == (2, 96) and torch.isfinite(out.float()).all()

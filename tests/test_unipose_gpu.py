"""GPU: the UniPose deformable encoder / decoder layers on our kernels against the reference's own
classes (tests/golden/mod_unipose_layers.npz), module tolerance rule of test_modules_gpu.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from unipose_inputs import inputs  # noqa: E402
from weights_util import key_shapes  # noqa: E402
from test_unipose_cpu import build, run  # noqa: E402


def rel_l2(a, b):
    return float(torch.linalg.norm(a.float() - b.float()) / torch.linalg.norm(b.float()))


def test_unipose_layers_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "mod_unipose_layers.npz"))
    enc, dec = build()
    assert json.loads(str(g["enc_keys"])) == [list(k) for k in key_shapes(enc)]
    assert json.loads(str(g["dec_keys"])) == [list(k) for k in key_shapes(dec)]
    enc, dec = enc.to("cuda", torch.bfloat16), dec.to("cuda", torch.bfloat16)
    x = {k: v.cuda() for k, v in inputs().items()}
    e, d = run(enc, dec, x, c=lambda t: t.bfloat16() if t.is_floating_point() else t)
    sub = int(g["sub"])
    for name, got in (("enc", e[:, ::sub]), ("dec", d)):
        ref32 = torch.from_numpy(g[f"{name}_f32"]).cuda()
        ref16 = torch.from_numpy(g[f"{name}_refbf16"]).cuda()
        assert got.shape == ref32.shape and got.dtype == torch.bfloat16
        budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
        assert rel_l2(got, ref32) <= budget, (name, rel_l2(got, ref32), budget)

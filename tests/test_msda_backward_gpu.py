"""GPU parity of the MSDA backward operator against autograd through the reference's own
pytorch function (tests/golden/msda_bwd_*.npz, fp64) and against the C oracle (fp32); mirrors
mmcv/tests/test_ops/test_ms_deformable_attn.py:137-181 (channel counts 4, 30, 32, 71)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import msda_oracle as O  # noqa: E402

NAMES = ["msda_bwd_d32.npz", "msda_bwd_d4.npz", "msda_bwd_d30.npz", "msda_bwd_d71.npz"]


def dev(g, keys, dt):
    out = []
    for k in keys:
        t = torch.from_numpy(g[k]).cuda()
        out.append(t.to(dt) if t.is_floating_point() else t)
    return out


@pytest.mark.parametrize("name", NAMES)
def test_backward_fp64_vs_reference_autograd(golden_dir, name):
    import visionllm_b200.msda as ext
    g = np.load(os.path.join(golden_dir, name))
    v, sh, lsi, loc, w, go = dev(g, ["value", "shapes", "lsi", "loc", "attw", "grad_out"], torch.float64)
    gv, gl, gw = ext.ms_deform_attn_backward(v, sh, lsi, loc, w, go, 64)              # unipose / HF flavour
    assert (gv.cpu().numpy() - g["grad_value"]).__abs__().max() < 1e-9
    assert np.abs(gl.cpu().numpy() - g["grad_loc"]).max() < 1e-8 * max(1.0, np.abs(g["grad_loc"]).max())
    assert np.abs(gw.cpu().numpy() - g["grad_attw"]).max() < 1e-9
    # mmcv flavour: pre-zeroed buffers, keyword im2col_step, returns None
    b = [torch.zeros_like(v), torch.zeros_like(loc), torch.zeros_like(w)]
    assert ext.ms_deform_attn_backward(v, sh, lsi, loc, w, go, *b, im2col_step=64) is None
    assert torch.allclose(b[0], gv, rtol=0, atol=1e-12) and torch.equal(b[1], gl) and torch.equal(b[2], gw)


@pytest.mark.parametrize("name", NAMES)
def test_backward_fp32_vs_oracle(golden_dir, name):
    import visionllm_b200.msda as ext
    g = np.load(os.path.join(golden_dir, name))
    f32 = {k: g[k].astype(np.float32) if g[k].dtype == np.float64 else g[k] for k in g.files}
    rv, rl, rw = O.backward_kernel_semantics(f32["value"], f32["shapes"], f32["lsi"], f32["loc"], f32["attw"],
                                             f32["grad_out"])
    v, sh, lsi, loc, w, go = dev(f32, ["value", "shapes", "lsi", "loc", "attw", "grad_out"], torch.float32)
    gv, gl, gw = ext.ms_deform_attn_backward(v, sh, lsi, loc, w, go, 64)
    for got, ref in ((gv, rv), (gl, rl), (gw, rw)):                 # fp32 summation order differs (atomics, shuffles)
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_autograd_function_end_to_end():
    import visionllm_b200.msda as ext
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = torch.tensor([(5, 4), (3, 2)], device="cuda")
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    v = torch.randn(1, 26, 2, 8, device="cuda", dtype=torch.float64, generator=g, requires_grad=True)
    loc = torch.rand(1, 3, 2, 2, 2, 2, device="cuda", dtype=torch.float64, generator=g, requires_grad=True)
    w = torch.rand(1, 3, 2, 2, 2, device="cuda", dtype=torch.float64, generator=g, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b, c: ext.MultiScaleDeformableAttentionFunction.apply(a, shapes, lsi, b, c, 64),
                                    (v, loc, w), eps=1e-6, atol=1e-5, rtol=1e-4)


def test_backward_full_size_runs_and_matches_forward_linearity():
    """BASELINE shape (decoder, 900 queries): <grad_value, dv> == <grad_out, forward(dv)> (adjoint identity)."""
    import visionllm_b200.msda as ext
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = torch.tensor([(128, 128), (64, 64), (32, 32), (16, 16)], device="cuda")
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    v = torch.randn(2, S, 8, 32, device="cuda", generator=g)
    loc = torch.rand(2, 900, 8, 4, 4, 2, device="cuda", generator=g)
    w = torch.softmax(torch.randn(2, 900, 8, 16, device="cuda", generator=g), -1).view(2, 900, 8, 4, 4)
    go = torch.randn(2, 900, 256, device="cuda", generator=g)
    gv, gl, gw = ext.ms_deform_attn_backward(v, shapes, lsi, loc, w, go, 64)
    dv = torch.randn_like(v)
    lhs = (gv.double() * dv.double()).sum()
    rhs = (go.double() * ext.ms_deform_attn_forward(dv, shapes, lsi, loc, w, 64).double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-4 * abs(float(rhs)) + 1e-3

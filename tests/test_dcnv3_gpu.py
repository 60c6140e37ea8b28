"""GPU parity of the DCNv3 forward operator through the C-ABI (visionllm_b200.dcnv3) against the C oracle and the
reference-generated golden vectors; mirrors visionllmv2/model/ops_dcnv3/test.py:33-90."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dcnv3_oracle as DO  # noqa: E402


def run(g_in, g_off, g_mask, p, scale, flags):
    import visionllm_b200.dcnv3 as ext
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (g_in, g_off, g_mask)]
    return ext.dcnv3_forward(*t, *p[:8], p[8], p[9], scale, 2 if t[0].shape[0] % 2 == 0 else 1, flags=flags).cpu().numpy()


@pytest.mark.parametrize("name", ["dcnv3_ref_testpy.npz", "dcnv3_ref_c32_s2.npz", "dcnv3_ref_c32_k5.npz"])
@pytest.mark.parametrize("flags", [0, 1])
def test_golden_reference_outputs(golden_dir, name, flags):
    g = np.load(os.path.join(golden_dir, name))
    p = [int(x) for x in g["params"]]
    out = run(g["input"], g["offset"], g["mask"], p, float(g["offset_scale"]), flags)
    ref = g["out_f64"]
    assert np.allclose(out, ref, rtol=1e-2, atol=1e-3)        # ops_dcnv3/test.py:79
    assert np.abs(out - ref).max() < 1e-7


def make(N, H, W, G, C, K, stride, pad, dil, seed, amp):
    rng = np.random.default_rng(seed)
    H_out = (H + 2 * pad - (dil * (K - 1) + 1)) // stride + 1
    W_out = (W + 2 * pad - (dil * (K - 1) + 1)) // stride + 1
    inp = rng.standard_normal((N, H, W, G * C), dtype=np.float32)
    off = ((rng.random((N, H_out, W_out, G * K * K * 2), dtype=np.float32) - 0.5) * amp).astype(np.float32)
    m = rng.random((N, H_out, W_out, G, K * K), dtype=np.float32) + 1e-3
    m = (m / m.sum(-1, keepdims=True)).reshape(N, H_out, W_out, G * K * K).astype(np.float32)
    return inp, off, m, [K, K, stride, stride, pad, pad, dil, dil, G, C]


CASES = [(2, 16, 16, 4, 32, 3, 1, 1, 1, 4.0), (1, 17, 23, 10, 32, 3, 1, 1, 1, 6.0), (2, 9, 7, 3, 32, 3, 2, 1, 2, 3.0),
         (1, 12, 12, 2, 32, 5, 1, 2, 1, 2.0), (2, 8, 8, 4, 16, 3, 1, 1, 1, 10.0), (1, 6, 5, 2, 7, 3, 1, 1, 1, 2.0)]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_strict_bit_exact_and_fast_close(case):
    inp, off, m, p = make(*case[:9], seed=11, amp=case[9])
    ref = DO.forward(inp, off, m, *p[:8], p[8], p[9], 1.5)
    strict = run(inp, off, m, p, 1.5, 1)
    assert np.array_equal(strict.view(np.uint32), ref.view(np.uint32))
    fast = run(inp, off, m, p, 1.5, 0)
    assert np.abs(fast - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_errors_and_interior_partition_of_unity():
    import visionllm_b200.dcnv3 as ext
    inp, off, m, p = make(2, 64, 64, 4, 32, 3, 1, 1, 1, seed=3, amp=0.5)
    t = [torch.from_numpy(a).cuda() for a in (np.ones_like(inp), off, m)]
    out = ext.dcnv3_forward(*t, *p[:8], p[8], p[9], 1.0, 2)
    assert (out[:, 2:-2, 2:-2] - 1).abs().max().item() < 1e-5      # interior taps: bilinear stencils sum to 1
    with pytest.raises(RuntimeError):
        ext.dcnv3_forward(t[0].half(), t[1], t[2], *p[:8], p[8], p[9], 1.0, 2)
    with pytest.raises(RuntimeError):
        ext.dcnv3_forward(t[0].transpose(1, 2), t[1], t[2], *p[:8], p[8], p[9], 1.0, 2)
    with pytest.raises(RuntimeError):
        ext.dcnv3_forward(t[0].cpu(), t[1], t[2], *p[:8], p[8], p[9], 1.0, 2)


# ---- backward: dcnv3_backward vs the C oracle and the reference-autograd golden ----
def run_bwd(inp, off, m, gout, p, scale):
    import visionllm_b200.dcnv3 as ext
    t = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in (inp, off, m)]
    go = torch.from_numpy(np.ascontiguousarray(gout, dtype=np.float32)).cuda()
    return [x.cpu().numpy() for x in ext.dcnv3_backward(*t, *p[:8], p[8], p[9], scale, go, 2 if t[0].shape[0] % 2 == 0 else 1)]


@pytest.mark.parametrize("name", ["dcnv3_bwd_testpy.npz", "dcnv3_bwd_c32_s2.npz"])
def test_backward_golden_reference_autograd(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    p = [int(x) for x in g["params"]]
    gi, go, gm = run_bwd(g["input"], g["offset"], g["mask"], g["grad_out"], p, float(g["offset_scale"]))
    for ours, key in ((gi, "grad_input"), (go, "grad_offset"), (gm, "grad_mask")):
        ref = g[key]
        assert np.abs(ours - ref).max() <= 2e-5 * np.abs(ref).max(), key       # fp32 op vs fp64 autograd


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_backward_vs_oracle(case):
    inp, off, m, p = make(*case[:9], seed=12, amp=case[9])
    rng = np.random.default_rng(5)
    H_out, W_out = off.shape[1], off.shape[2]
    gout = rng.standard_normal((inp.shape[0], H_out, W_out, p[8] * p[9]), dtype=np.float32)
    ref = DO.backward(inp, off, m, gout, *p[:8], p[8], p[9], 1.5)
    ours = run_bwd(inp, off, m, gout, p, 1.5)
    for a, b, k in zip(ours, ref, ("grad_input", "grad_offset", "grad_mask")):
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max()), k      # atomics / shuffle order only


def test_backward_autograd_function_and_adjoint_identity():
    """<grad_out, J_input . d> == <J_input^T . grad_out, d>: forward is linear in `input`, so the backward's
    grad_input must be its exact adjoint (size-independent property, run at an InternImage-like shape)."""
    import visionllm_b200.dcnv3 as ext
    inp, off, m, p = make(2, 40, 56, 10, 32, 3, 1, 1, 1, seed=21, amp=3.0)
    ti, to, tm = (torch.from_numpy(a).cuda() for a in (inp, off, m))
    ti.requires_grad_(True); to.requires_grad_(True); tm.requires_grad_(True)
    out = ext.DCNv3Function.apply(ti, to, tm, *p[:8], p[8], p[9], 1.0, 2)
    gout = torch.randn_like(out)
    gi, go, gm = torch.autograd.grad(out, (ti, to, tm), gout)
    d = torch.randn_like(ti)
    lhs = (ext.dcnv3_forward(d.contiguous(), to.detach(), tm.detach(), *p[:8], p[8], p[9], 1.0, 2).double() * gout.double()).sum()
    rhs = (gi.double() * d.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-4 * abs(lhs.item())
    # forward is linear in mask too: <grad_mask, mask> == <grad_out, out>
    assert abs((gm.double() * tm.detach().double()).sum().item() - (gout.double() * out.detach().double()).sum().item()) \
        <= 1e-4 * abs((gout.double() * out.detach().double()).sum().item())
    with pytest.raises(RuntimeError):
        ext.dcnv3_backward(ti.detach(), to.detach(), tm.detach(), *p[:8], p[8], p[9], 1.0, gout[:, :, :-1].contiguous(), 2)

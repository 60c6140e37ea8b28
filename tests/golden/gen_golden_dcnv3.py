"""DCNv3 golden vectors from the reference's own pure-PyTorch implementation (build container only).
`dcnv3_core_pytorch` and its two helpers are lifted verbatim (ast, no edits) out of
/root/reference/VisionLLMv2/visionllmv2/model/ops_dcnv3/functions/dcnv3_func.py:64-161 and run in fp64 on
(a) the vector of ops_dcnv3/test.py:20-60 (N,M,D = 2,4,16; 8x8; 3x3; offset_scale 2.0; seed 3) and
(b) a group_channels = 32 case with stride 2 / dilation 2 / non-square input.  Only inputs + outputs are stored."""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/VisionLLMv2/visionllmv2/model/ops_dcnv3/functions/dcnv3_func.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load():
    tree = ast.parse(open(REF).read())
    names = {"_get_reference_points", "_generate_dilation_grids", "dcnv3_core_pytorch"}
    mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names], type_ignores=[])
    ns = {"torch": torch, "F": F}
    exec(compile(mod, REF, "exec"), ns)
    return ns["dcnv3_core_pytorch"]


def case(fn, name, N, H_in, W_in, M, D, Kh, Kw, stride, pad, dil, offset_scale, seed, off_amp):
    torch.manual_seed(seed)
    P = Kh * Kw
    H_out = (H_in + 2 * pad - (dil * (Kh - 1) + 1)) // stride + 1
    W_out = (W_in + 2 * pad - (dil * (Kw - 1) + 1)) // stride + 1
    inp = torch.rand(N, H_in, W_in, M * D) * 0.01
    offset = torch.rand(N, H_out, W_out, M * P * 2) * off_amp
    mask = torch.rand(N, H_out, W_out, M, P) + 1e-5
    mask /= mask.sum(-1, keepdim=True)
    mask = mask.reshape(N, H_out, W_out, M * P)
    out = fn(inp.double(), offset.double(), mask.double(), Kh, Kw, stride, stride, pad, pad, dil, dil, M, D, offset_scale)
    np.savez(os.path.join(OUT, name), input=inp.numpy(), offset=offset.numpy(), mask=mask.numpy(), out_f64=out.numpy(),
             params=np.array([Kh, Kw, stride, stride, pad, pad, dil, dil, M, D], dtype=np.int64),
             offset_scale=np.float64(offset_scale))
    print(name, tuple(out.shape), float(out.abs().mean()))


def bwd_case(fn, name, N, H_in, W_in, M, D, Kh, Kw, stride, pad, dil, offset_scale, seed, off_amp):
    """fp64 autograd through the reference's dcnv3_core_pytorch: the gradients `dcnv3_backward`
    (ops_dcnv3/src/dcnv3.h:40-59; col2im kernels dcnv3_im2col_cuda.cuh:82-147,278-370) must reproduce."""
    torch.manual_seed(seed)
    P = Kh * Kw
    H_out = (H_in + 2 * pad - (dil * (Kh - 1) + 1)) // stride + 1
    W_out = (W_in + 2 * pad - (dil * (Kw - 1) + 1)) // stride + 1
    inp = (torch.rand(N, H_in, W_in, M * D) - 0.3).double().requires_grad_(True)
    offset = (torch.rand(N, H_out, W_out, M * P * 2) * off_amp - off_amp / 3).double().requires_grad_(True)
    mask = torch.rand(N, H_out, W_out, M, P) + 1e-5
    mask = (mask / mask.sum(-1, keepdim=True)).reshape(N, H_out, W_out, M * P).double().requires_grad_(True)
    out = fn(inp, offset, mask, Kh, Kw, stride, stride, pad, pad, dil, dil, M, D, offset_scale)
    gout = torch.randn(out.shape, dtype=torch.float64)
    gi, go, gm = torch.autograd.grad(out, (inp, offset, mask), gout)
    np.savez(os.path.join(OUT, name), input=inp.detach().numpy(), offset=offset.detach().numpy(), mask=mask.detach().numpy(),
             grad_out=gout.numpy(), grad_input=gi.numpy(), grad_offset=go.numpy(), grad_mask=gm.numpy(),
             params=np.array([Kh, Kw, stride, stride, pad, pad, dil, dil, M, D], dtype=np.int64),
             offset_scale=np.float64(offset_scale))
    print(name, tuple(out.shape), float(gi.abs().mean()), float(go.abs().mean()), float(gm.abs().mean()))


if __name__ == "__main__":
    fn = load()
    bwd_case(fn, "dcnv3_bwd_testpy.npz", 2, 8, 8, 4, 16, 3, 3, 1, 1, 1, 2.0, 13, 4.0)
    bwd_case(fn, "dcnv3_bwd_c32_s2.npz", 2, 13, 10, 3, 32, 3, 3, 2, 1, 2, 1.0, 14, 3.0)
    case(fn, "dcnv3_ref_testpy.npz", 2, 8, 8, 4, 16, 3, 3, 1, 1, 1, 2.0, 3, 10.0)
    case(fn, "dcnv3_ref_c32_s2.npz", 2, 13, 10, 3, 32, 3, 3, 2, 1, 2, 1.0, 4, 3.0)
    case(fn, "dcnv3_ref_c32_k5.npz", 1, 12, 12, 2, 32, 5, 5, 1, 2, 1, 1.5, 5, 2.0)

"""Generate golden vectors by RUNNING THE REFERENCE's own Python code.

Run in the build container only (needs /root/reference; the GPU box does not
have it):  python tests/golden/gen_golden.py

MSDA: the reference function ``multi_scale_deformable_attn_pytorch`` is lifted
verbatim (ast, no edits) out of
/root/reference/VisionLLMv2/mmcv/mmcv/ops/multi_scale_deform_attn.py:100-159
and executed on
  * the mmcv unit-test vector (mmcv/tests/test_ops/test_ms_deformable_attn.py:72-134:
    seed 3, N,M,D=1,2,2, Lq,L,P=2,2,2, shapes [(6,4),(3,2)]) in fp64 and fp32;
  * larger seeded cases (non-power-of-two levels, out-of-range locations,
    pixel-centre reference points) in fp64.
Only inputs + the reference's outputs are stored; no reference source.
"""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/VisionLLMv2/mmcv/mmcv/ops/multi_scale_deform_attn.py"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_fn():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "multi_scale_deformable_attn_pytorch"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "F": F}
    exec(compile(mod, REF, "exec"), ns)
    return ns["multi_scale_deformable_attn_pytorch"]


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def main():
    ref = load_reference_fn()
    # --- 1. mmcv unit-test vector -------------------------------------------------
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attw = torch.rand(N, Lq, M, L, P) + 1e-5
    attw /= attw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out64 = ref(value.double(), shapes, loc.double(), attw.double())
    out32 = ref(value, shapes, loc, attw)
    np.savez(os.path.join(OUT, "msda_mmcv_seed3.npz"), value=value.numpy(), shapes=shapes.numpy(),
             lsi=lsi_of(shapes).numpy(), loc=loc.numpy(), attw=attw.numpy(),
             out_f64=out64.numpy(), out_f32=out32.numpy())
    print("mmcv seed3 fp64:", out64.flatten().tolist())

    # --- 2. larger seeded cases (fp64 reference output) -------------------------------
    def case(name, shapes_l, N, M, D, Lq, P, seed, mode):
        g = torch.Generator().manual_seed(seed)
        shapes = torch.as_tensor(shapes_l, dtype=torch.long)
        L = len(shapes_l)
        S = int(shapes.prod(1).sum())
        value = torch.randn(N, S, M, D, generator=g)
        if mode == "uniform":          # includes out-of-range samples on every side
            loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.3 - 0.15
        elif mode == "pixel":          # encoder-style: queries are the pixels, refs at pixel centres
            assert Lq == S
            refs = []
            for (H, W) in shapes_l:
                ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
                refs.append(torch.stack(((xs + 0.5) / W, (ys + 0.5) / H), -1).reshape(-1, 2))
            ref_pts = torch.cat(refs, 0)[None, :, None, None, None, :]
            off = torch.randn(N, Lq, M, L, P, 2, generator=g) * 0.05
            off[:, ::3] = 0.0           # every third query samples exactly at pixel centres
            loc = (ref_pts + off).contiguous()
        attw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
        out = ref(value.double(), shapes, loc.double(), attw.double())
        np.savez(os.path.join(OUT, name), value=value.numpy(), shapes=shapes.numpy(), lsi=lsi_of(shapes).numpy(),
                 loc=loc.numpy(), attw=attw.numpy(), out_f64=out.numpy())
        print(name, tuple(out.shape), float(out.abs().mean()))

    # --- 3. backward: autograd through the reference's own pytorch function, fp64 (the channel counts of
    #        mmcv/tests/test_ops/test_ms_deformable_attn.py:137-181 that are small enough to ship) ------------
    def bwd_case(name, shapes_l, N, M, D, Lq, P, seed):
        g = torch.Generator().manual_seed(seed)
        shapes = torch.as_tensor(shapes_l, dtype=torch.long)
        L = len(shapes_l)
        S = int(shapes.prod(1).sum())
        value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64).requires_grad_()
        loc = (torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * 1.2 - 0.1).requires_grad_()
        attw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, dtype=torch.float64), -1).view(
            N, Lq, M, L, P).detach().requires_grad_()
        gout = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)
        out = ref(value, shapes, loc, attw)
        gv, gl, gw = torch.autograd.grad(out, (value, loc, attw), gout)
        np.savez(os.path.join(OUT, name), value=value.detach().numpy(), shapes=shapes.numpy(),
                 lsi=lsi_of(shapes).numpy(), loc=loc.detach().numpy(), attw=attw.detach().numpy(),
                 grad_out=gout.numpy(), grad_value=gv.numpy(), grad_loc=gl.numpy(), grad_attw=gw.numpy())
        print(name, float(gv.abs().mean()), float(gl.abs().mean()), float(gw.abs().mean()))

    bwd_case("msda_bwd_d32.npz", [(9, 11), (5, 6), (3, 3), (2, 2)], 2, 4, 32, 23, 4, 31)
    bwd_case("msda_bwd_d4.npz", [(6, 4), (3, 2)], 1, 2, 4, 7, 2, 32)
    bwd_case("msda_bwd_d30.npz", [(7, 5)], 1, 3, 30, 5, 3, 33)
    bwd_case("msda_bwd_d71.npz", [(6, 6), (3, 3)], 1, 2, 71, 4, 2, 34)

    case("msda_ref_d32_npot.npz", [(13, 17), (7, 9), (4, 5), (2, 3)], 2, 8, 32, 37, 4, 11, "uniform")
    case("msda_ref_d32_pixel.npz", [(12, 10), (6, 5), (3, 3)], 1, 4, 32, 159, 4, 12, "pixel")
    case("msda_ref_d16_l2p2.npz", [(9, 11), (5, 6)], 3, 3, 16, 21, 2, 13, "uniform")
    case("msda_ref_d71_l1p5.npz", [(8, 7)], 1, 2, 71, 9, 5, 14, "uniform")


if __name__ == "__main__":
    main()

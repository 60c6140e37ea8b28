"""Golden vectors for the InternImage backbone from the REFERENCE's own `InternImage` class
(grounding_dino/modeling_ov_grounding_dino_mask_dn.py:4978-5152) run on CPU in this build container with the reference's
pure-PyTorch core op (`core_op='DCNv3_pytorch'`, ops_dcnv3/modules/dcnv3.py:86-208 -- same parameters and keys as the
CUDA-backed `DCNv3`, whose extension cannot run here).  A scaled-down InternImage-H: every H/G-only switch on
(dw_kernel_size=5, res_post_norm, level2_post_norm, center_feature_scale).  Stored: bf16-representable input, fp32
outputs of all four levels, the reference's own bf16 run, and the state-dict key/shape list."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

CFG = dict(core_op="DCNv3_pytorch", channels=32, depths=[1, 1, 6, 1], groups=[2, 4, 8, 16], mlp_ratio=4.,
           drop_path_rate=0., norm_layer="LN", layer_scale=None, offset_scale=1.0, post_norm=False, dw_kernel_size=5,
           res_post_norm=True, level2_post_norm=True, level2_post_norm_block_ids=[2, 5], center_feature_scale=True,
           with_cp=False, out_indices=(0, 1, 2, 3))


def load_reference():
    return ref_shim.load_gdino_with_dcnv3()[1]


def main():
    gd = load_reference()
    torch.manual_seed(0)
    ref = gd.InternImage(**CFG).eval()
    ref.load_state_dict(seeded_state_dict(ref, 303))
    # the deformable branch is zero-initialised in the reference (offset/mask = 0): give it life, bf16-representable
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 96, generator=g).to(torch.bfloat16).float()
    with torch.no_grad():
        out32 = ref(x)
        # bf16 leg = the reference's deployed precision: its CUDA-backed DCNv3 module upcasts x / offset / mask to fp32
        # around the core op and casts the result back (modules/dcnv3.py:331-341); DCNv3_pytorch's core is wrapped the
        # same way here (grid_sample itself has no mixed-dtype path).
        mm = sys.modules["refpkg_dcnv3.modules.dcnv3"]
        core = mm.dcnv3_core_pytorch

        def core_like_cuda_module(x_, offset, mask, *a):
            return core(x_.float(), offset.float(), mask.float(), *a).to(x_.dtype)
        mm.dcnv3_core_pytorch = core_like_cuda_module
        ref16 = ref.to(torch.bfloat16)
        out16 = [o.float() for o in ref16(x.to(torch.bfloat16))]
    np.savez(os.path.join(HERE, "mod_internimage_small.npz"), pixel_values=x.numpy(),
             keys=json.dumps(key_shapes(ref)), cfg=json.dumps({k: v for k, v in CFG.items() if k != "core_op"}),
             **{f"out_f32_{i}": o.numpy() for i, o in enumerate(out32)},
             **{f"out_refbf16_{i}": o.numpy() for i, o in enumerate(out16)})
    for i, (a, b) in enumerate(zip(out32, out16)):
        print(i, tuple(a.shape), float(a.abs().mean()), "bf16 rel_l2", float((a - b).norm() / a.norm()))


if __name__ == "__main__":
    main()

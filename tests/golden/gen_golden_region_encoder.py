"""Golden vectors for the region encoder from the REFERENCE's own `RegionEncoder`
(visionllmv2/model/region_encoder.py:66-145) run on CPU in this build container, all three pooling modes.  The
'grid_sample' mode draws its points with torch.multinomial (random by design): the draw of the fp32 run is recorded and
replayed for the bf16 run, and stored so the B200 module can be fed the same points."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

REF = "/root/reference/VisionLLMv2/visionllmv2/model/region_encoder.py"
CFG = dict(hidden_dim=64, embed_dim=256, out_dim=96, patch_size=14)


def main():
    spec = importlib.util.spec_from_file_location("ref_region_encoder", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(3)
    B, Hh = 3, 112
    images = torch.randn(B, 3, Hh, Hh, generator=g).to(torch.bfloat16).float()
    masks = torch.zeros(B, 1, Hh, Hh)
    masks[0, 0, 10:70, 20:90] = 1
    masks[1, 0, 50:, :40] = 1
    masks[2, 0, 5:9, 5:9] = 1                                   # a tiny region: fewer points than one query group
    feats = [(torch.randn(B, 64, 256, generator=g) * 0.5).to(torch.bfloat16).float() for _ in range(3)]
    out = {"images": images.numpy(), "masks": masks.numpy(), "cfg": json.dumps(CFG)}
    for i, f in enumerate(feats):
        out[f"feat_{i}"] = f.numpy()
    for mode in ("mean", "cross_attn", "grid_sample"):
        torch.manual_seed(0)
        ref = mod.RegionEncoder(mask_pool_type=mode, **CFG).eval()
        ref.load_state_dict(seeded_state_dict(ref, 77))
        out[f"keys_{mode}"] = json.dumps(key_shapes(ref))
        drawn = []
        real = mod.rand_sample

        def record(x, divisor, max_len):
            p = real(x, divisor, max_len)
            drawn.append(p.clone())
            return p
        mod.rand_sample = record
        # the reference's 'cross_attn' branch flattens masks_out in place, so it only runs with ONE feature level
        lv = feats[:1] if mode == "cross_attn" else feats
        with torch.no_grad():
            o32 = ref(images, masks, lv)
        if mode == "grid_sample":
            pts = list(drawn)                                   # one draw per (level, region): 3 x B
            assert len(pts) == len(feats) * B
            replay = iter(pts)
            mod.rand_sample = lambda x, d, m: next(replay).to(x.dtype)
            for i, p in enumerate(pts):
                out[f"points_{i // B}_{i % B}"] = p.numpy()
        with torch.no_grad():
            r16 = ref.to(torch.bfloat16)
            o16 = r16(images.to(torch.bfloat16), masks.to(torch.bfloat16), [f.to(torch.bfloat16) for f in lv]).float()
        mod.rand_sample = real
        out[f"out_f32_{mode}"], out[f"out_refbf16_{mode}"] = o32.numpy(), o16.numpy()
        print(mode, tuple(o32.shape), float(o32.abs().mean()), "bf16 rel_l2", float((o32 - o16).norm() / o32.norm()))
    np.savez_compressed(os.path.join(HERE, "mod_region_encoder.npz"), **out)


if __name__ == "__main__":
    main()

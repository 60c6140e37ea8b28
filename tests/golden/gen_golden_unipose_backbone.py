"""Golden vectors for UniPose's image backbone (the remaining piece of the UniPose path): the
REFERENCE's own `Joiner(SwinTransformer, PositionEmbeddingSineHW)` (visionllmv2/model/unipose/modeling_unipose.py:
1212-1226, 1638-1858, 1037-1078) executed on CPU in this build container on a padded two-image NestedTensor, fp32 and
bf16 ("the reference's deployed precision").  Small widths (head_dim 32 like every Swin preset), sizes chosen so that the
patch-embed pad, the window pad (25 x 35 tokens -> 28 x 35) and the odd patch merges (25 -> 13 -> 7 -> 4) all happen."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

from unipose_inputs import BACKBONE, backbone_inputs  # noqa: E402


def main():
    mu = ref_shim.load_unipose()
    c = BACKBONE
    x, mask = backbone_inputs()
    out = {}
    keys = None
    for name, dtype in (("f32", torch.float32), ("refbf16", torch.bfloat16)):
        swin = mu.SwinTransformer(pretrain_img_size=224, embed_dim=c["embed_dim"], depths=c["depths"], num_heads=c["num_heads"],
                                  window_size=c["window_size"], out_indices=tuple(c["out_indices"]), dilation=False)
        pe = mu.PositionEmbeddingSineHW(c["hidden_dim"] // 2, temperatureH=20, temperatureW=20, normalize=True)
        joiner = mu.Joiner(swin, pe).eval()
        joiner.load_state_dict(seeded_state_dict(joiner, 53))
        keys = key_shapes(joiner)
        joiner = joiner.to(dtype)
        with torch.no_grad():
            feats, poss = joiner(mu.NestedTensor(x.to(dtype), mask))
        out[name] = ([f.tensors.float() for f in feats], [f.mask for f in feats], [p.float() for p in poss])
    save = dict(keys=json.dumps(keys))
    for i in range(len(out["f32"][0])):
        save[f"map{i}_f32"], save[f"map{i}_refbf16"] = out["f32"][0][i].numpy(), out["refbf16"][0][i].numpy()
        save[f"mask{i}"], save[f"pos{i}_f32"] = out["f32"][1][i].numpy(), out["f32"][2][i].numpy()
        a, b = out["f32"][0][i], out["refbf16"][0][i]
        print(i, tuple(a.shape), "bf16 rel_l2", float((a - b).norm() / a.norm()), "mask true", int(out["f32"][1][i].sum()))
    np.savez_compressed(os.path.join(HERE, "mod_unipose_backbone.npz"), **save)


if __name__ == "__main__":
    main()

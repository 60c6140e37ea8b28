"""Golden vectors for the UniPose two-stage keypoint decoder (second half) from the REFERENCE's own
`TransformerDecoder` (visionllmv2/model/unipose/modeling_unipose.py:2869-3130) with its own decoder layers, `MLP` heads,
`ContrastiveAssign` and `UniPose.prepare_for_mask` (:887-917), run on CPU in this build container through
ref_shim.load_unipose (MSDA served by the reference's pure-PyTorch core).  fp32 outputs + the reference's bf16 run ON THE fp32 RUN'S top-50 selection (the selection is a discrete function of
bf16-noisy logits: the reference's free bf16 run picks another set / order, recorded as topk_refbf16)."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

from unipose_inputs import DEC, decoder_inputs  # noqa: E402

GROUP_STEP = 6               # every 6th (box + keypoints) query group of the expanded layers is stored


def build(mu):
    """Shared with the tests through `build_decoder(ns)`: ns provides the class names of either implementation."""
    c = DEC
    layer = mu.DeformableTransformerDecoderLayer(c["d_model"], c["d_ffn"], 0.0, "relu", 4, c["n_heads"], 4,
                                                 use_text_cross_attention=True)
    dec = mu.TransformerDecoder(layer, c["num_layers"], nn.LayerNorm(c["d_model"]) if not hasattr(mu, "_LN") else mu._LN(c["d_model"]),
                                return_intermediate=True, d_model=c["d_model"], query_dim=4, num_feature_levels=4,
                                deformable_decoder=True, rm_dec_query_scale=True, dec_layer_share=False,
                                num_box_decoder_layers=c["num_box_decoder_layers"], num_body_points=c["num_body_points"]).eval()
    dec.load_state_dict(seeded_state_dict(dec, 51))
    keys = key_shapes(dec)
    bbox, pose, pose_hw = mu.MLP(256, 256, 4, 3), mu.MLP(256, 256, 2, 3), mu.MLP(256, 256, 2, 3)
    for m, seed in ((bbox, 52), (pose, 53), (pose_hw, 54)):
        m.load_state_dict(seeded_state_dict(m, seed))
    nl, nb = c["num_layers"], c["num_box_decoder_layers"]
    cls = mu.ContrastiveAssign()
    dec.bbox_embed = nn.ModuleList([bbox for _ in range(nl)])
    dec.class_embed = nn.ModuleList([cls for _ in range(nl)])
    dec.pose_embed = nn.ModuleList([pose for _ in range(nl - nb + 1)])
    dec.pose_hw_embed = nn.ModuleList([pose_hw for _ in range(nl - nb)])
    return dec, keys


def reference_mask(mu, kpt_vis):
    """The reference's own prepare_for_mask (it hard-codes .to('cuda'): Tensor.to is patched for this one call)."""
    kpt_mask = torch.cat((torch.ones_like(kpt_vis)[..., 0].unsqueeze(-1), kpt_vis), dim=-1)
    real_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else real_to(self, *a, **k)
    try:
        out = mu.UniPose.prepare_for_mask(SimpleNamespace(num_body_points=DEC["num_body_points"], nheads=DEC["n_heads"]), kpt_mask)
    finally:
        torch.Tensor.to = real_to
    return out[3]


def run(dec, x, mask2, dtype):
    c = lambda t: t.to(dtype) if t.is_floating_point() else t  # noqa: E731
    dec = dec.to(dtype)
    text_dict = {"encoded_text": c(x["encoded_text"]), "text_token_mask": ~x["text_mask"]}
    with torch.no_grad():
        hs, refs = dec(tgt=c(x["tgt"]).clone(), memory=c(x["memory"]), tgt_mask=None, tgt_mask2=mask2,
                       memory_key_padding_mask=x["pad"], pos=None, refpoints_unsigmoid=c(x["ref_unsig"]),
                       level_start_index=x["lsi"], spatial_shapes=x["shapes"], valid_ratios=c(x["valid_ratios"]),
                       memory_text=c(x["memory_text"]), text_attention_mask=x["text_mask"], text_dict=text_dict,
                       kpt_embed=c(x["kpt_embed"]))
    return [h.float() for h in hs], [r.float() for r in refs]


def main():
    mu = ref_shim.load_unipose()
    x = decoder_inputs()
    mask2 = reference_mask(mu, x["kpt_vis"])
    # the top-k the reference takes inside the loop is not returned: record it through torch.topk
    picked = []
    real_topk = torch.topk

    def spy(*a, **k):
        out = real_topk(*a, **k)
        picked.append(out[1].clone())
        return out

    out = {}
    forced = [None]

    def spy(*a, **k):                                      # noqa: F811
        o = real_topk(*a, **k)
        idx = o[1] if forced[0] is None else forced[0]
        picked.append(idx.clone())
        return o[0], idx

    for name, dtype in (("f32", torch.float32), ("refbf16", torch.bfloat16), ("refbf16_forced", torch.bfloat16)):
        dec, keys = build(mu)
        forced[0] = out["f32"][2] if name.endswith("forced") else None
        torch.topk = spy
        try:
            hs, refs = run(dec, x, mask2, dtype)
        finally:
            torch.topk = real_topk
        out[name] = (hs, refs, picked[-1])
    hs32, refs32, top32 = out["f32"]
    hs16, refs16, top16 = out["refbf16"]
    B = mask2.shape[0] // DEC["n_heads"]
    m = mask2.view(B, DEC["n_heads"], *mask2.shape[1:])
    assert bool((m == m[:, :1]).all()), "mask differs across heads"
    nb, group = DEC["num_box_decoder_layers"], DEC["num_body_points"] + 1
    rows = torch.cat([torch.arange(gi * group, (gi + 1) * group) for gi in range(0, 50, GROUP_STEP)])   # stored expanded queries
    sel = lambda i, t: t if i < nb else t[:, rows]                                                       # noqa: E731
    hsF, refsF = out["refbf16_forced"][0], out["refbf16_forced"][1]
    np.savez_compressed(
        os.path.join(HERE, "mod_unipose_decoder.npz"), keys=json.dumps(keys), group_step=np.int64(GROUP_STEP),
        mask2_bits=np.packbits(m[:, 0].numpy(), axis=-1), mask2_shape=np.array(m[:, 0].shape),
        topk_f32=top32.numpy(), topk_refbf16=top16.numpy(),
        **{f"hs{i}_f32": sel(i, h).numpy() for i, h in enumerate(hs32)},
        **{f"hs{i}_refbf16": sel(i, h).numpy() for i, h in enumerate(hsF)},
        **{f"ref{i}_f32": sel(i - 1, r).numpy() for i, r in enumerate(refs32)},
        **{f"ref{i}_refbf16": sel(i - 1, r).numpy() for i, r in enumerate(refsF)})
    for i, (a, b) in enumerate(zip(hs32, out["refbf16_forced"][0])):
        print("forced-selection hs", i, "bf16 rel_l2", float((a - b).norm() / a.norm()))
    for i, (a, b) in enumerate(zip(refs32, out["refbf16_forced"][1])):
        print("forced-selection ref", i, "bf16 max abs", float((a - b).abs().max()))
    for i, (a, b) in enumerate(zip(hs32, hs16)):
        print("hs", i, tuple(a.shape), float(a.abs().mean()), "bf16 rel_l2", float((a - b).norm() / a.norm()))
    for i, (a, b) in enumerate(zip(refs32, refs16)):
        print("ref", i, tuple(a.shape), "bf16 max abs", float((a - b).abs().max()))
    print("topk equal across dtypes:", bool((top32 == top16).all()), "set-equal:",
          [set(top32[:, b].tolist()) == set(top16[:, b].tolist()) for b in range(top32.shape[1])])


if __name__ == "__main__":
    main()

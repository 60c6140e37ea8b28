"""Golden vectors for the UniPose transformer: the REFERENCE's own `DeformableTransformer`
(visionllmv2/model/unipose/modeling_unipose.py:2206-2700: text-fused deformable encoder, two-stage 'standard' query
selection, two-stage keypoint decoder) with its own `MLP` / `ContrastiveAssign` heads bound like `UniPose.__init__` does
(:233-257), `generate_masks_with_text_query_masks` (:928) and `prepare_for_mask` (:887), run on CPU in this build container
through ref_shim.load_unipose.  fp32 outputs + the reference's bf16 run on the fp32 run's two top-k selections."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

from gen_golden_unipose_decoder import reference_mask  # noqa: E402
from unipose_inputs import TR, transformer_inputs, transformer_kwargs  # noqa: E402

GROUP_STEP = 8


def build(mu):
    c = TR
    kw = transformer_kwargs()
    try:
        tr = mu.DeformableTransformer(num_box_decoder_layers=c["num_box_decoder_layers"], num_body_points=c["num_body_points"], **kw)
    except TypeError:                                     # the reference takes the two decoder sizes through attributes
        tr = mu.DeformableTransformer(**kw)
        tr.decoder.num_box_decoder_layers = c["num_box_decoder_layers"]
        tr.decoder.num_body_points = c["num_body_points"]
        tr.decoder.kpt_index = [x for x in range(50 * (c["num_body_points"] + 1)) if x % (c["num_body_points"] + 1) != 0]
        tr.decoder.hw_append = nn.Embedding(c["num_body_points"] - 17, 2)
    tr = tr.eval()
    tr.load_state_dict(seeded_state_dict(tr, 61))
    keys = key_shapes(tr)
    bbox, pose, pose_hw = mu.MLP(256, 256, 4, 3), mu.MLP(256, 256, 2, 3), mu.MLP(256, 256, 2, 3)
    for m, seed in ((bbox, 62), (pose, 63), (pose_hw, 64)):
        m.load_state_dict(seeded_state_dict(m, seed))
    nl, nb = c["num_decoder_layers"], c["num_box_decoder_layers"]
    cls = mu.ContrastiveAssign()
    tr.decoder.bbox_embed = nn.ModuleList([bbox for _ in range(nl)])
    tr.decoder.class_embed = nn.ModuleList([cls for _ in range(nl)])
    tr.decoder.pose_embed = nn.ModuleList([pose for _ in range(nl - nb + 1)])
    tr.decoder.pose_hw_embed = nn.ModuleList([pose_hw for _ in range(nl - nb)])
    tr.enc_out_bbox_embed, tr.enc_out_class_embed = bbox, cls
    return tr, keys


def text_dict_of(mu, x, dtype):
    sa, pid = mu.generate_masks_with_text_query_masks(x["obj_mask"])
    return {"encoded_text": x["encoded_text"].to(dtype), "text_token_mask": x["obj_mask"].bool(), "position_ids": pid,
            "text_self_attention_masks": sa}


def run(tr, mu, x, mask2, dtype):
    c = lambda t: t.to(dtype)  # noqa: E731
    tr = tr.to(dtype)
    with torch.no_grad():
        hs, refs, hs_enc, ref_enc, init_box = tr([c(s) for s in x["srcs"]], x["masks"], None, [c(p) for p in x["poss"]], None,
                                                 None, mask2, text_dict_of(mu, x, dtype), None, None, c(x["kpt_embed"]))
    return [h.float() for h in hs], [r.float() for r in refs], hs_enc.float(), ref_enc.float(), init_box.float()


def main():
    mu = ref_shim.load_unipose()
    x = transformer_inputs()
    mask2 = reference_mask(mu, x["kpt_vis"]) if TR["num_body_points"] == 19 else None
    real_topk = torch.topk
    picked, forced = [], [None]

    def spy(*a, **k):
        o = real_topk(*a, **k)
        i = len(picked)
        idx = o[1] if forced[0] is None else forced[0][i]
        picked.append(idx.clone())
        return o[0], idx

    out = {}
    for name, dtype in (("f32", torch.float32), ("refbf16", torch.bfloat16)):
        tr, keys = build(mu)
        picked.clear()
        forced[0] = out["f32"][5] if name != "f32" else None
        torch.topk = spy
        try:
            res = run(tr, mu, x, mask2, dtype)
        finally:
            torch.topk = real_topk
        out[name] = (*res, list(picked))
    sa, pid = mu.generate_masks_with_text_query_masks(x["obj_mask"])
    nb, group = TR["num_box_decoder_layers"], TR["num_body_points"] + 1
    rows = torch.cat([torch.arange(gi * group, (gi + 1) * group) for gi in range(0, 50, GROUP_STEP)])
    sel = lambda i, t: t if i < nb else t[:, rows]  # noqa: E731
    save = dict(keys=json.dumps(keys), group_step=np.int64(GROUP_STEP), text_sa=sa.numpy(), text_pid=pid.numpy(),
                topk_enc=out["f32"][5][0].numpy(), topk_dec=out["f32"][5][1].numpy())
    for name in ("f32", "refbf16"):
        hs, refs, hs_enc, ref_enc, init_box, _ = out[name]
        for i, h in enumerate(hs):
            save[f"hs{i}_{name}"] = sel(i, h).numpy()
        for i, r in enumerate(refs):
            save[f"ref{i}_{name}"] = sel(i - 1, r).numpy()
        save[f"hs_enc_{name}"], save[f"ref_enc_{name}"], save[f"init_box_{name}"] = hs_enc.numpy(), ref_enc.numpy(), init_box.numpy()
    np.savez_compressed(os.path.join(HERE, "mod_unipose_transformer.npz"), **save)
    a, b = out["f32"], out["refbf16"]
    for i, (p, q) in enumerate(zip(a[0], b[0])):
        print("hs", i, tuple(p.shape), float(p.abs().mean()), "bf16 rel_l2", float((p - q).norm() / p.norm()))
    for i, (p, q) in enumerate(zip(a[1], b[1])):
        print("ref", i, tuple(p.shape), "bf16 max abs", float((p - q).abs().max()))
    print("hs_enc bf16 rel_l2", float((a[2] - b[2]).norm() / a[2].norm()), "topk shapes", [tuple(t.shape) for t in a[5]])


if __name__ == "__main__":
    main()

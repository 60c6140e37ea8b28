"""Golden vectors for the GDINO heads / query selection / post-processing from the REFERENCE's own code
(build container only).  Reference classes come through ref_shim (gd.py); post_process_* and box_cxcywh_to_xyxy are
lifted verbatim (ast) from visionllmv2/eval/eval_det.py:18-104 and visionllmv2/util/box_ops.py:13-22."""
import ast
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

V = "/root/reference/VisionLLMv2/visionllmv2"


def lift(path, names, ns):
    tree = ast.parse(open(path).read())
    mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names], type_ignores=[])
    exec(compile(mod, path, "exec"), ns)


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def main():
    torch.manual_seed(0)
    cfgm, gd = ref_shim.load_gdino()
    out = {}
    seed = 21
    while True:      # pick a vector whose top-k boundary is not inside a group of tied logits (padded pixels tie)
        g = torch.Generator().manual_seed(seed)
        ok = _selection_case(gd, g, out)
        if ok:
            break
        seed += 1
    print("selection seed", seed)
    _rest(gd, g, out)


def _selection_case(gd, g, out):
    # ---- proposals + enc_output/LN + contrastive class + bbox head + top-k (gd.py:2228-2276, 2503-2534) ----
    shapes_l = [(10, 14), (5, 7), (3, 4), (2, 2)]
    shapes = torch.tensor(shapes_l)
    S = sum(h * w for h, w in shapes_l)
    B, C, T, NQ = 2, 256, 9, 20
    enc = bf16r(torch.randn(B, S, C, generator=g))
    pad = torch.zeros(B, S, dtype=torch.bool)
    m0 = torch.zeros(10, 14, dtype=torch.bool); m0[:, 11:] = True; m0[8:, :] = True     # image 1 padded right/bottom
    pos = 0
    for (h, w) in shapes_l:
        mm = F.interpolate(m0[None, None].float(), size=(h, w)).bool()[0, 0]
        pad[1, pos:pos + h * w] = mm.flatten(); pos += h * w
    fake = SimpleNamespace(enc_output=torch.nn.Linear(C, C), enc_output_norm=torch.nn.LayerNorm(C))
    holder = torch.nn.Module(); holder.enc_output = fake.enc_output; holder.enc_output_norm = fake.enc_output_norm
    holder.load_state_dict(seeded_state_dict(holder, 404))
    bbox = gd.GroundingDinoMLPPredictionHead(C, C, 4, 3); bbox.load_state_dict(seeded_state_dict(bbox, 405))
    contr = gd.GroundingDinoContrastiveEmbedding(SimpleNamespace(max_text_len=16))
    text = bf16r(torch.randn(B, T, C, generator=g))
    tmask = torch.ones(B, T, dtype=torch.bool); tmask[1, 6:] = False
    with torch.no_grad():
        oq, prop = gd.OVGroundingDinoModel.gen_encoder_output_proposals(fake, enc, pad, shapes)
        cls = contr(oq, text, tmask)
        coord = bbox(oq) + prop
        topk_logits = cls.max(-1)[0]
        topk = torch.topk(topk_logits, NQ, dim=1)[1]
        ref_pts = torch.gather(coord, 1, topk.unsqueeze(-1).repeat(1, 1, 4)).sigmoid()
    srt = topk_logits.sort(1, descending=True)[0]
    gaps = (srt[:, :NQ] - srt[:, 1:NQ + 1])
    if float(gaps.min()) < 1e-3:
        return False
    out.update(enc=enc, pad=pad.numpy(), shapes=shapes.numpy(), text=text, tmask=tmask.numpy(), oq=oq, prop=prop, cls=cls,
               coord=coord, topk=topk.numpy(), ref_pts=ref_pts, keys_holder=np.array(json.dumps(key_shapes(holder))),
               keys_bbox=np.array(json.dumps(key_shapes(bbox))), topk_margin=np.float64(gaps.min()))
    return True


def _rest(gd, g, out):
    B, C = 2, 256
    # ---- mask head (gd.py:2278-2281) ----
    Q, Hm, Wm = 7, 12, 10
    me = gd.GroundingDinoMLPPredictionHead(C, C, C, 3); me.load_state_dict(seeded_state_dict(me, 406))
    hs = bf16r(torch.randn(B, Q, C, generator=g)); mf = bf16r(torch.randn(B, C, Hm, Wm, generator=g))
    with torch.no_grad():
        masks = gd.OVGroundingDinoModel.forward_seg_heads(SimpleNamespace(mask_embed=me), hs, mf)
    out.update(mask_hs=hs, mask_feat=mf, masks=masks, keys_me=np.array(json.dumps(key_shapes(me))))
    # ---- post-processing (eval_det.py:18-104) ----
    ns = {"torch": torch, "F": F, "List": list}
    lift(f"{V}/util/box_ops.py", {"box_cxcywh_to_xyxy"}, ns)
    lift(f"{V}/eval/eval_det.py", {"post_process_det_gdino", "post_process_instseg_gdino"}, ns)
    Qp, K = 30, 16
    logits = torch.full((B, Qp, 256), float("-inf")); logits[:, :, :K] = torch.randn(B, Qp, K, generator=g) * 2
    boxes = torch.rand(B, Qp, 4, generator=g) * 0.5 + 0.2
    pmasks = torch.randn(B, Qp, 16, 20, generator=g) * 3
    outs = SimpleNamespace(gdino_outputs=SimpleNamespace(logits=logits, pred_boxes=boxes, pred_masks=pmasks))
    tsz = [(60, 75), (48, 64)]; isz = [(61, 77), (50, 70)]
    det = ns["post_process_det_gdino"](outs, torch.tensor(tsz), K, threshold=0.3, topk=25)
    seg = ns["post_process_instseg_gdino"](outs, tsz, isz, num_classes=K, topk=10, mask_stride=4)
    out.update(pp_logits=logits, pp_boxes=boxes, pp_masks=pmasks, pp_tsz=np.array(tsz), pp_isz=np.array(isz))
    for i, d in enumerate(det):
        out[f"det{i}_scores"], out[f"det{i}_labels"], out[f"det{i}_boxes"] = d["scores"], d["labels"].numpy(), d["boxes"]
    for i, d in enumerate(seg):
        out[f"seg{i}_scores"], out[f"seg{i}_labels"], out[f"seg{i}_boxes"] = d["scores"], d["labels"].numpy(), d["boxes"]
        out[f"seg{i}_masks"] = np.packbits(d["masks"].numpy())
        out[f"seg{i}_masks_shape"] = np.array(d["masks"].shape)
    np.savez_compressed(os.path.join(HERE, "mod_gdino_heads.npz"),
                        **{k: (v.detach().float().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote mod_gdino_heads.npz; min gap inside top-k+1", float(out["topk_margin"]))


if __name__ == "__main__":
    main()

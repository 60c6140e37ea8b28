"""GPU parity of the GDINO query selection, heads and post-processing (visionllm_b200.gdino_heads) against goldens
produced by the reference's own code (tests/golden/gen_golden_gdino_heads.py).

Integer results (top-k proposal indices, topk//K, topk%K, binary masks) are compared EXACTLY, with the same score
tensors as input -- that is the "bit-exact box/mask indices" contract.  Dense heads run in bf16 and are compared to
the reference's fp32 output: |out - ref| <= 2^-7*|ref| + 2e-2*max|ref| (three chained bf16 GEMMs + LN)."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from weights_util import key_shapes, seeded_state_dict  # noqa: E402


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "mod_gdino_heads.npz"))


def t(g, k, dt=None):
    x = torch.from_numpy(g[k]).cuda()
    return x.to(dt) if dt is not None else x


def close(out, ref, extra=2e-2):
    out, ref = out.float(), ref.float()
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(out), fin)
    assert torch.equal(out[~fin], ref[~fin])                      # +-inf placed identically
    tol = ref[fin].abs() * 2.0 ** -7 + extra * ref[fin].abs().max()
    assert ((out[fin] - ref[fin]).abs() <= tol).all(), float((out[fin] - ref[fin]).abs().max())


def load(mod, g, key, seed):
    assert json.loads(str(g[key])) == [list(k) for k in key_shapes(mod)]
    mod.load_state_dict(seeded_state_dict(mod, seed))
    return mod.to("cuda", torch.bfloat16).eval()


def test_proposals_class_bbox_and_topk(g):
    from visionllm_b200 import gdino_heads as H
    prop_mod = load(H.EncoderOutputProposals(256), g, "keys_holder", 404)
    bbox = load(H.GroundingDinoMLPPredictionHead(256, 256, 4, 3), g, "keys_bbox", 405)
    contr = H.GroundingDinoContrastiveEmbedding(SimpleNamespace(max_text_len=16))
    shapes = t(g, "shapes").long()
    pad = t(g, "pad").bool()
    oq, prop = prop_mod(t(g, "enc", torch.bfloat16), pad, shapes)
    # proposals: the +inf sentinels (padded / out-of-range pixels) must sit identically; the finite values are
    # fp32 log/div whose last ulp differs between CPU and GPU libm
    ref_prop = t(g, "prop")
    assert torch.equal(torch.isinf(prop), torch.isinf(ref_prop))
    fin = torch.isfinite(ref_prop)
    assert torch.allclose(prop[fin], ref_prop[fin], rtol=1e-5, atol=1e-6)
    close(oq, t(g, "oq"))
    cls = contr(oq, t(g, "text", torch.bfloat16), t(g, "tmask").bool())
    assert cls.dtype == torch.float32 and cls.shape == (2, oq.shape[1], 16)
    close(cls, t(g, "cls"), extra=3e-2)
    coord = bbox(oq).float() + prop
    close(coord, t(g, "coord"), extra=3e-2)
    # integer contract: same logits in -> identical proposal indices and gathered reference points
    idx, ref_pts, _, _, _ = H.select_topk_proposals(t(g, "cls"), t(g, "coord"), t(g, "oq"), 20)
    assert torch.equal(idx.cpu(), torch.from_numpy(g["topk"]))
    assert torch.allclose(ref_pts, t(g, "ref_pts"), rtol=1e-6, atol=1e-7)          # sigmoid: libm ulp


def test_mask_head_einsum_as_gemm(g):
    from visionllm_b200 import gdino_heads as H
    me = load(H.GroundingDinoMLPPredictionHead(256, 256, 256, 3), g, "keys_me", 406)
    out = H.forward_seg_heads(me, t(g, "mask_hs", torch.bfloat16), t(g, "mask_feat", torch.bfloat16))
    assert out.shape == (2, 7, 12, 10)
    close(out, t(g, "masks"), extra=2e-2)
    # channels_last features take the zero-copy path and give the same result
    cl = t(g, "mask_feat", torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert torch.equal(H.forward_seg_heads(me, t(g, "mask_hs", torch.bfloat16), cl), out)


def test_post_process_indices_exact(g):
    from visionllm_b200 import gdino_heads as H
    logits, boxes, masks = t(g, "pp_logits"), t(g, "pp_boxes"), t(g, "pp_masks")
    tsz = [tuple(int(v) for v in r) for r in g["pp_tsz"]]
    isz = [tuple(int(v) for v in r) for r in g["pp_isz"]]
    det, _, _ = H.post_process_det_gdino(logits, boxes, tsz, 16, threshold=0.3, topk=25)
    seg = H.post_process_instseg_gdino(logits, boxes, masks, tsz, isz, num_classes=16, topk=10, mask_stride=4)
    for i in range(2):
        assert np.array_equal(det[i]["labels"].cpu().numpy(), g[f"det{i}_labels"])
        assert torch.allclose(det[i]["scores"].cpu(), torch.from_numpy(g[f"det{i}_scores"]), rtol=0, atol=1e-6)
        assert torch.allclose(det[i]["boxes"].cpu(), torch.from_numpy(g[f"det{i}_boxes"]), rtol=1e-6, atol=1e-4)
        assert np.array_equal(seg[i]["labels"].cpu().numpy(), g[f"seg{i}_labels"])
        assert torch.allclose(seg[i]["boxes"].cpu(), torch.from_numpy(g[f"seg{i}_boxes"]), rtol=1e-6, atol=1e-4)
        shape = tuple(int(v) for v in g[f"seg{i}_masks_shape"])
        ref = np.unpackbits(g[f"seg{i}_masks"])[:int(np.prod(shape))].reshape(shape).astype(bool)
        got = seg[i]["masks"].cpu().numpy()
        # bilinear resize on GPU vs CPU differs in the last ulp: allow pixels whose pre-threshold logit is ~0
        assert got.shape == ref.shape and (got != ref).mean() < 1e-3


def test_patch2query_mean(g):
    from visionllm_b200 import gdino_heads as H
    torch.manual_seed(0)
    head = H.GroundingDinoMLPPredictionHead(512, 256, 256, 3).to("cuda", torch.bfloat16)
    x = torch.randn(2, 5, 4, 512, device="cuda").bfloat16()
    out = H.patch2query_mean(head, x)
    ref = x.float()
    for i, l in enumerate(head.layers):
        ref = torch.nn.functional.linear(ref, l.weight.float(), l.bias.float())
        if i < 2:
            ref = torch.relu(ref)
    close(out, ref.mean(-2), extra=2e-2)


# ---- fused post-processing kernels (csrc/postproc.cu) vs the reference's torch primitives on the GPU ----
def _ref_det(logits, boxes, tsz, K, topk):
    """eval_det.py:26-46 with the reference's own torch calls (on the GPU, like the eval loop)."""
    from visionllm_b200.gdino_heads import box_cxcywh_to_xyxy
    lg = logits[:, :, :K]
    prob = lg.sigmoid().view(lg.shape[0], -1)
    k = min(topk, prob.size(1))
    tv, ti = torch.topk(prob, k, dim=1)
    tb = torch.div(ti, lg.shape[2], rounding_mode="floor")
    bx = torch.gather(box_cxcywh_to_xyxy(boxes), 1, tb.unsqueeze(-1).repeat(1, 1, 4))
    img_h = torch.Tensor([i[0] for i in tsz]); img_w = torch.Tensor([i[1] for i in tsz])
    bx = bx * torch.stack([img_w, img_h, img_w, img_h], dim=1).to(bx.device)[:, None, :]
    return tv, ti, tb, ti % lg.shape[2], bx


@pytest.mark.parametrize("Q,K,ld,topk", [(100, 80, 256, 100), (900, 80, 256, 300), (20, 16, 16, 25), (7, 3, 8, 100), (300, 256, 256, 1000)])
def test_fused_det_topk_equals_torch_primitives(Q, K, ld, topk):
    from visionllm_b200 import gdino_heads as H
    g = torch.Generator(device="cuda").manual_seed(Q + K)
    B = 3
    # distinct probabilities (torch.topk leaves the order of ties unspecified; ties have their own test below): a permuted
    # grid of logits whose sigmoid values are >= 40 fp32 ulps apart
    logits = torch.full((B, Q, ld), float("-inf"), device="cuda")   # -inf = the class padding of the contrastive head (gd.py:1415-1428)
    for b in range(B):
        perm = torch.randperm(Q * K, device="cuda", generator=g)
        logits[b, :, :K] = torch.linspace(-8.0, 4.0, Q * K, device="cuda")[perm].view(Q, K)
    boxes = torch.rand(B, Q, 4, device="cuda", generator=g)
    tsz = [(480, 640), (1024, 1024), (333, 500)]
    tv, ti, tb, tl, bx = _ref_det(logits, boxes, tsz, K, topk)
    scores, labels, fboxes, idx, box_idx = H.det_topk_fused(logits, boxes, tsz, K, topk)
    assert torch.equal(scores, tv), (scores - tv).abs().max()        # sigmoid bit-identical to torch's, same order
    assert torch.equal(idx, ti) and torch.equal(box_idx, tb) and torch.equal(labels, tl)
    assert idx.dtype == torch.int64 and torch.equal(fboxes, bx)


def test_fused_det_topk_ties_take_the_lowest_index_first():
    from visionllm_b200 import gdino_heads as H
    logits = torch.zeros(1, 10, 4, device="cuda")
    logits[0, 3, 1] = 5.0; logits[0, 7, 2] = 5.0; logits[0, 2, 0] = 9.0
    boxes = torch.rand(1, 10, 4, device="cuda")
    scores, labels, _, idx, box_idx = H.det_topk_fused(logits, boxes, [(10, 10)], 4, 6)
    assert idx[0].tolist() == [8, 13, 30, 0, 1, 2]                   # 9.0; the two 5.0 by index; then the zeros by index
    assert box_idx[0].tolist() == [2, 3, 7, 0, 0, 0] and labels[0].tolist() == [0, 1, 2, 0, 1, 2]
    assert scores[0, 3:].eq(0.5).all()


@pytest.mark.parametrize("case", [(64, 64, 4, (250, 256), (480, 517)), (32, 40, 4, (128, 160), (128, 160)),
                                  (56, 56, 4, (200, 224), (1000, 1333)), (16, 24, 2, (31, 47), (97, 61))])
def test_fused_mask_chain_equals_torch_interpolate_chain(case):
    """masks[box_idx] -> F.interpolate x stride -> crop -> F.interpolate to the original size -> sigmoid() > 0.5 (eval_det.py
    :88-99) vs the single fused kernel.  The boolean masks agree except where the pre-threshold value is within fp32
    rounding of 0 (FMA contraction inside ATen's kernel is the compiler's choice): < 1e-5 of the pixels, and never where
    |value| > 1e-5."""
    import torch.nn.functional as F
    from visionllm_b200 import gdino_heads as H
    Hm, Wm, stride, isz, tsz = case
    g = torch.Generator(device="cuda").manual_seed(Hm * 7 + Wm)
    Q, k = 20, 9
    masks = torch.randn(Q, Hm, Wm, device="cuda", generator=g) * 4
    bi = torch.randint(0, Q, (k,), device="cuda", generator=g)
    m = F.interpolate(masks[bi][:, None], size=(Hm * stride, Wm * stride), mode="bilinear", align_corners=False)
    m = m[:, :, :isz[0], :isz[1]]
    pre = F.interpolate(m, size=tsz, mode="bilinear", align_corners=False)[:, 0]
    ref = pre.sigmoid() > 0.5
    got = H.mask_chain_fused(masks, bi, isz, tsz, stride)
    assert got.dtype == torch.bool and got.shape == ref.shape
    diff = got != ref
    assert diff.float().mean().item() < 1e-5, diff.float().mean().item()
    assert (pre[diff].abs() < 1e-5).all()


def test_fused_instseg_path_matches_torch_path(g):
    from visionllm_b200 import gdino_heads as H
    logits, boxes, masks = t(g, "pp_logits"), t(g, "pp_boxes"), t(g, "pp_masks")
    tsz = [tuple(int(v) for v in r) for r in g["pp_tsz"]]
    isz = [tuple(int(v) for v in r) for r in g["pp_isz"]]
    fused = H.post_process_instseg_gdino(logits, boxes, masks, tsz, isz, num_classes=16, topk=10, mask_stride=4)
    H.FUSED_POSTPROCESS = False
    try:
        eager = H.post_process_instseg_gdino(logits, boxes, masks, tsz, isz, num_classes=16, topk=10, mask_stride=4)
    finally:
        H.FUSED_POSTPROCESS = True
    for a, b in zip(fused, eager):
        assert torch.equal(a["topk_indexes"], b["topk_indexes"]) and torch.equal(a["labels"], b["labels"])
        assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"])
        assert (a["masks"] != b["masks"]).float().mean().item() < 1e-5

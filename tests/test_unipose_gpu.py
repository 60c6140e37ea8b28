"""GPU: the UniPose deformable encoder / decoder layers on our kernels against the reference's own
classes (tests/golden/mod_unipose_layers.npz), module tolerance rule of test_modules_gpu.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from unipose_inputs import inputs  # noqa: E402
from weights_util import key_shapes  # noqa: E402
from test_unipose_cpu import build, run  # noqa: E402


def rel_l2(a, b):
    return float(torch.linalg.norm(a.float() - b.float()) / torch.linalg.norm(b.float()))


def test_unipose_layers_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "mod_unipose_layers.npz"))
    enc, dec = build()
    assert json.loads(str(g["enc_keys"])) == [list(k) for k in key_shapes(enc)]
    assert json.loads(str(g["dec_keys"])) == [list(k) for k in key_shapes(dec)]
    enc, dec = enc.to("cuda", torch.bfloat16), dec.to("cuda", torch.bfloat16)
    x = {k: v.cuda() for k, v in inputs().items()}
    e, d = run(enc, dec, x, c=lambda t: t.bfloat16() if t.is_floating_point() else t)
    sub = int(g["sub"])
    for name, got in (("enc", e[:, ::sub]), ("dec", d)):
        ref32 = torch.from_numpy(g[f"{name}_f32"]).cuda()
        ref16 = torch.from_numpy(g[f"{name}_refbf16"]).cuda()
        assert got.shape == ref32.shape and got.dtype == torch.bfloat16
        budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
        assert rel_l2(got, ref32) <= budget, (name, rel_l2(got, ref32), budget)


def test_unipose_keypoint_decoder_matches_reference(golden_dir):
    """The two-stage keypoint decoder (modeling_unipose.py:2869-3130) on our kernels in bf16 vs the reference's fp32 run,
    module rule against the reference's own bf16 error.  The top-50 box selection in the middle of the loop is a discrete
    function of bf16-noisy logits (the reference's free bf16 run picks another set / order than its fp32 run), so (1) the
    free run must select as stably as the reference's bf16 run does and (2) values are compared on the fp32 run's
    selection (`forced_topk`), which is also how the golden's bf16 run was made."""
    from unipose_inputs import DEC, decoder_inputs
    from test_unipose_cpu import build_decoder, decoder_mask, run_decoder, stored_rows
    g = np.load(os.path.join(golden_dir, "mod_unipose_decoder.npz"))
    dec, keys = build_decoder()
    assert json.loads(str(g["keys"])) == [list(k) for k in keys]
    dec = dec.to("cuda", torch.bfloat16)
    x = {k: v.cuda() for k, v in decoder_inputs().items()}
    mask2 = decoder_mask(x["kpt_vis"], DEC["n_heads"], DEC["num_body_points"])
    cast = lambda t: t.bfloat16() if t.is_floating_point() else t  # noqa: E731
    top32, top16 = torch.from_numpy(g["topk_f32"]), torch.from_numpy(g["topk_refbf16"])
    run_decoder(dec, x, mask2, c=cast)                                           # free run: selection stability
    ours = dec.topk_proposals.cpu()
    assert ours.shape == top32.shape
    for b in range(ours.shape[1]):
        common = len(set(ours[:, b].tolist()) & set(top32[:, b].tolist()))
        common_ref = len(set(top16[:, b].tolist()) & set(top32[:, b].tolist()))
        assert common >= common_ref - 2, (b, common, common_ref)
    dec.forced_topk = top32.cuda()
    hs, refs = run_decoder(dec, x, mask2, c=cast)
    rows, nb = stored_rows(g).cuda(), DEC["num_box_decoder_layers"]
    for i, h in enumerate(hs):
        ref32, ref16 = torch.from_numpy(g[f"hs{i}_f32"]).cuda(), torch.from_numpy(g[f"hs{i}_refbf16"]).cuda()
        h = h if i < nb else h[:, rows]
        assert h.shape == ref32.shape and h.dtype == torch.bfloat16
        budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
        assert rel_l2(h, ref32) <= budget, (i, rel_l2(h, ref32), budget)
    for i, r in enumerate(refs):
        r32, r16 = torch.from_numpy(g[f"ref{i}_f32"]).cuda(), torch.from_numpy(g[f"ref{i}_refbf16"]).cuda()
        r = r if i - 1 < nb else r[:, rows]
        assert r.shape == r32.shape
        assert (r.float() - r32).abs().max().item() <= 1.5 * (r16 - r32).abs().max().item() + 4e-3, i


def test_unipose_transformer_matches_reference(golden_dir):
    """The whole UniPose transformer (text-fused deformable encoder -> two-stage query selection -> keypoint decoder,
    modeling_unipose.py:2206-2700) on our kernels in bf16 vs the reference's fp32 run; both top-k selections are pinned to
    the fp32 run's (as for the golden's bf16 run), module rule against the reference's own bf16 error."""
    from unipose_inputs import TR, transformer_inputs
    from test_unipose_cpu import build_transformer, decoder_mask, run_transformer, stored_rows
    g = np.load(os.path.join(golden_dir, "mod_unipose_transformer.npz"))
    tr, keys = build_transformer()
    assert json.loads(str(g["keys"])) == [list(k) for k in keys]
    tr = tr.to("cuda", torch.bfloat16)
    x = transformer_inputs()
    x = {k: ([t.cuda() for t in v] if isinstance(v, list) else v.cuda()) for k, v in x.items()}
    mask2 = decoder_mask(x["kpt_vis"], TR["nhead"], TR["num_body_points"])
    cast = lambda t: t.bfloat16() if t.is_floating_point() else t  # noqa: E731
    tr.forced_topk = torch.from_numpy(g["topk_enc"]).cuda()
    tr.decoder.forced_topk = torch.from_numpy(g["topk_dec"]).cuda()
    hs, refs, hs_enc, ref_enc, init_box = run_transformer(tr, x, mask2, c=cast)
    rows, nb = stored_rows(g).cuda(), TR["num_box_decoder_layers"]
    for i, h in enumerate(hs):
        ref32, ref16 = torch.from_numpy(g[f"hs{i}_f32"]).cuda(), torch.from_numpy(g[f"hs{i}_refbf16"]).cuda()
        h = h if i < nb else h[:, rows]
        assert h.shape == ref32.shape and h.dtype == torch.bfloat16
        budget = 1.5 * rel_l2(ref16, ref32) + 1e-3
        assert rel_l2(h, ref32) <= budget, (i, rel_l2(h, ref32), budget)
    for i, r in enumerate(refs):
        r32, r16 = torch.from_numpy(g[f"ref{i}_f32"]).cuda(), torch.from_numpy(g[f"ref{i}_refbf16"]).cuda()
        r = r if i - 1 < nb else r[:, rows]
        assert r.shape == r32.shape
        assert (r.float() - r32).abs().max().item() <= 1.5 * (r16 - r32).abs().max().item() + 4e-3, i
    e32, e16 = torch.from_numpy(g["hs_enc_f32"]).cuda(), torch.from_numpy(g["hs_enc_refbf16"]).cuda()
    assert rel_l2(hs_enc, e32) <= 1.5 * rel_l2(e16, e32) + 1e-3


def test_unipose_model_matches_reference_forward(golden_dir):
    """B200UniPose on our kernels (bf16) vs the reference's own `UniPose.forward` run (fp32 golden, bf16 golden on the same
    selections): pred_boxes / pred_keypoints within 1.5x the reference's own bf16 deviation, finite logits likewise, the
    -inf pattern of the padded class slots exact."""
    from unipose_inputs import model_inputs
    from test_unipose_cpu import build_unipose_model
    g = np.load(os.path.join(golden_dir, "mod_unipose_model.npz"))
    m = build_unipose_model().to("cuda", torch.bfloat16)
    x = model_inputs()
    cast = lambda t: (t.bfloat16() if t.is_floating_point() else t).cuda()  # noqa: E731
    feats = [(cast(t), mk.cuda()) for t, mk in x["feats"]]
    poss = [cast(p) for p in x["poss"]]
    tq = {k: cast(v) for k, v in x["text_query"].items()}
    m.transformer.forced_topk = torch.from_numpy(g["topk_enc"]).cuda()
    m.transformer.decoder.forced_topk = torch.from_numpy(g["topk_dec"]).cuda()
    out = m(feats, poss, tq, sample_mask=x["sample_mask"].cuda())
    l32, l16 = torch.from_numpy(g["logits_f32"]).cuda(), torch.from_numpy(g["logits_refbf16"]).cuda()
    fin = torch.isfinite(l32)
    assert torch.equal(torch.isfinite(out.pred_logits), fin)
    assert (out.pred_logits[fin] - l32[fin]).abs().max().item() <= 1.5 * (l16[fin] - l32[fin]).abs().max().item() + 5e-2
    for name, got in (("boxes", out.pred_boxes), ("keypoints", out.pred_keypoints)):
        r32, r16 = torch.from_numpy(g[f"{name}_f32"]).cuda(), torch.from_numpy(g[f"{name}_refbf16"]).cuda()
        assert got.dtype == torch.float32 and got.shape == r32.shape
        assert (got - r32).abs().max().item() <= 1.5 * (r16 - r32).abs().max().item() + 4e-3, name


def test_unipose_backbone_matches_reference(golden_dir):
    """UniPose's image backbone on our kernels (bf16) vs the reference's own `Joiner(SwinTransformer, PositionEmbeddingSineHW)`
    (fp32 golden + its bf16 run): every out-index map at rel_l2 <= 1.5 x the reference's own bf16 error + 1e-3, padding masks
    exact, position embeddings at one bf16 rounding of the fp32 golden."""
    from unipose_inputs import backbone_inputs
    from test_unipose_backbone_cpu import build_joiner
    g = np.load(os.path.join(golden_dir, "mod_unipose_backbone.npz"))
    j = build_joiner().to("cuda", torch.bfloat16)
    x, mask = backbone_inputs()
    feats, poss = j(x.bfloat16().cuda(), mask.cuda())
    assert len(feats) == 3
    for i, ((t, m), p) in enumerate(zip(feats, poss)):
        r32, r16 = torch.from_numpy(g[f"map{i}_f32"]).cuda(), torch.from_numpy(g[f"map{i}_refbf16"]).cuda()
        assert t.shape == r32.shape and t.dtype == torch.bfloat16
        assert rel_l2(t, r32) <= 1.5 * rel_l2(r16, r32) + 1e-3, (i, rel_l2(t, r32), rel_l2(r16, r32))
        assert torch.equal(m.cpu(), torch.from_numpy(g[f"mask{i}"])), i
        assert p.dtype == torch.bfloat16
        assert (p.float().cpu() - torch.from_numpy(g[f"pos{i}_f32"])).abs().max().item() <= 2 ** -8, i


def test_unipose_model_from_pixels_runs_on_gpu():
    """`B200UniPose(backbone=...).forward_samples` (the reference's :430 wiring) end to end on the GPU kernels: same result
    as forward() on the backbone's own maps, finite boxes / keypoints."""
    from unipose_inputs import MODEL, TR, backbone_inputs, model_inputs, transformer_kwargs
    from test_unipose_backbone_cpu import build_joiner
    from weights_util import seeded_state_dict
    from visionllm_b200.unipose import B200UniPose
    j = build_joiner()
    kw = transformer_kwargs()
    for k in ("d_model", "nhead", "num_queries", "num_feature_levels"):
        kw.pop(k)
    m = B200UniPose(hidden_dim=TR["d_model"], l_hidden_size=MODEL["l_hidden"], backbone_channels=tuple(j.num_channels),
                    num_feature_levels=4, num_queries=TR["num_queries"], num_body_points=TR["num_body_points"],
                    num_box_decoder_layers=TR["num_box_decoder_layers"], nheads=TR["nhead"], backbone=j, **kw).eval()
    m.load_state_dict(seeded_state_dict(m, 5))
    m = m.to("cuda", torch.bfloat16)
    x, mask = backbone_inputs()
    cast = lambda t: (t.bfloat16() if t.is_floating_point() else t).cuda()  # noqa: E731
    tq = {k: cast(v) for k, v in model_inputs()["text_query"].items()}
    a = m.forward_samples(x.bfloat16().cuda(), mask.cuda(), tq)
    feats, poss = m.backbone(x.bfloat16().cuda(), mask.cuda())
    b = m(feats, poss, tq, sample_mask=mask.cuda())
    assert torch.equal(a.pred_boxes, b.pred_boxes) and torch.equal(a.pred_keypoints, b.pred_keypoints)
    assert torch.isfinite(a.pred_boxes).all() and torch.isfinite(a.pred_keypoints).all()
    assert a.pred_boxes.dtype == torch.float32


def test_unipose_from_pixels_replays_as_one_cuda_graph():
    """The whole UniPose forward from pixels has no host synchronisation (constant index tensors are cached per device, shapes
    stay on the host): captured into one CUDA graph it replays to exactly the eager result, also after the inputs change."""
    from unipose_inputs import MODEL, TR, backbone_inputs, model_inputs, transformer_kwargs
    from test_unipose_backbone_cpu import build_joiner
    from weights_util import seeded_state_dict
    from visionllm_b200.graphs import GraphedForward
    from visionllm_b200.unipose import B200UniPose
    j = build_joiner()
    kw = transformer_kwargs()
    for k in ("d_model", "nhead", "num_queries", "num_feature_levels"):
        kw.pop(k)
    m = B200UniPose(hidden_dim=TR["d_model"], l_hidden_size=MODEL["l_hidden"], backbone_channels=tuple(j.num_channels),
                    num_feature_levels=4, num_queries=TR["num_queries"], num_body_points=TR["num_body_points"],
                    num_box_decoder_layers=TR["num_box_decoder_layers"], nheads=TR["nhead"], backbone=j, **kw).eval()
    m.load_state_dict(seeded_state_dict(m, 5))
    m = m.to("cuda", torch.bfloat16)
    x, mask = backbone_inputs()
    x, mask = x.bfloat16().cuda(), mask.cuda()
    cast = lambda t: (t.bfloat16() if t.is_floating_point() else t).cuda()  # noqa: E731
    tq = {k: cast(v) for k, v in model_inputs()["text_query"].items()}

    def fwd(images, obj, kpt):
        o = m.forward_samples(images, mask, dict(tq, obj_querys=obj, kpt_querys=kpt))
        return o.pred_boxes, o.pred_keypoints, o.pred_logits

    gf = GraphedForward(fwd)
    for trial in range(3):
        xi = x if trial == 0 else (x * (1.0 + 0.2 * trial)).contiguous()
        oq = tq["obj_querys"] if trial < 2 else (tq["obj_querys"] * 0.5).contiguous()
        eager = [t.clone() for t in fwd(xi, oq, tq["kpt_querys"])]
        got = gf(xi, oq, tq["kpt_querys"])
        for a, b in zip(eager, got):
            assert torch.equal(a, b), trial
    assert len(gf._cache) == 1 and gf.launches_per_replay > 100

"""GPU: the sequence-assembly kernels (csrc/seqglue.cu) against the vectorised torch indexing of modeling.py, which the
CPU tests pin index-for-index to the reference's python loops (modeling_visionllmv2.py:426-527, 582-605, 776-787) --
integer outputs and copied rows must be IDENTICAL; the fused pixel-shuffle LayerNorm within one bf16 ulp."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

IMP, EMB, DET, SEG, POSE, C = 900, 910, 901, 902, 903, 64


def _model():
    import torch.nn as nn
    from visionllm_b200.modeling import B200VisionLLMv2Model

    class FakeViT(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, patch_size=2)

    class FakeLLM(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, vocab_size=1000)
            self.emb = nn.Embedding(1000, C)
            self.lm_head = nn.Linear(C, 1000, bias=False)

        @property
        def dtype(self):
            return self.emb.weight.dtype

        def get_input_embeddings(self):
            return self.emb

    cfg = SimpleNamespace(use_pixelshuffle=False, vl_bridge_type="linear", vis_output_layer=-1, num_embs=4,
                          imp_token_id=IMP, emb_token_id=EMB, det_tool_id=DET, seg_tool_id=SEG, grd_tool_id=-1,
                          pose_tool_id=POSE)
    return B200VisionLLMv2Model(cfg, FakeViT(), FakeLLM()).to("cuda", torch.bfloat16).eval()


def _batch(seed, B=4, L=96, tiles=(2, 0, 1, 3), tpt=8):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 800, (B, L), generator=g)
    for b, n in enumerate(tiles):
        ids[b, :n * tpt] = IMP                                  # sample 1 has no image tokens: its tiles are skipped
    for b in range(B):
        p = 40 + 3 * b
        for tool in ((DET, SEG, POSE, DET)[b], DET, POSE)[: 1 + b % 3]:
            ids[b, p] = tool
            ids[b, p + 1:p + 5] = EMB
            p += 7
    return ids.cuda(), list(tiles), tpt


@pytest.mark.parametrize("seed", [0, 1])
def test_seq_index_and_assembly_equal_the_torch_indexing(seed):
    from visionllm_b200 import ops
    m = _model()
    ids, tiles, tpt = _batch(seed)
    B, L = ids.shape
    split = [t if t > 0 else 1 for t in tiles]                  # every sample carries >= 1 tile; sample 1 owns no <im_patch>
    feats = torch.randn(sum(split), tpt, C, device="cuda").bfloat16()
    emb0 = m.llm.get_input_embeddings()(ids)
    ref_ids, ref_emb = m.inject_emb(ids, emb0)
    ref_emb = m.scatter_image_tokens(ref_ids, ref_emb, feats, split)
    plan = ops.seq_index(ids, (DET, SEG, -1), (POSE,), EMB, 4, IMP, split, tpt)
    assert int(plan.status.item()) == 0
    got = ops.assemble_embeds(plan, m.llm.get_input_embeddings().weight, m.emb_embeddings_det.weight,
                              m.emb_embeddings_pose.weight, feats.reshape(-1, C))
    assert torch.equal(plan.new_ids, ref_ids) and torch.equal(got, ref_emb)
    # caller-provided inputs_embeds instead of the lookup
    base = torch.randn(B, L, C, device="cuda").bfloat16()
    r_ids, r_emb = m.inject_emb(ids, base)
    r_emb = m.scatter_image_tokens(r_ids, r_emb, feats, split)
    got2 = ops.assemble_embeds(plan, None, m.emb_embeddings_det.weight, m.emb_embeddings_pose.weight, feats.reshape(-1, C),
                               base_embeds=base)
    assert torch.equal(got2, r_emb)
    # [EMB] hidden states -> text_query / masks
    hidden = torch.randn(B, L, C, device="cuda").bfloat16()
    tq_ref, tm_ref = m.gather_text_query(ref_ids, hidden)
    mx = int((plan.emb_count // 4).max())
    tq, tm = ops.text_query_gather(plan, hidden, 4, mx)
    assert torch.equal(tq, tq_ref) and torch.equal(tm, tm_ref)


def test_seq_index_flags_the_cases_the_torch_path_refuses():
    from visionllm_b200 import ops
    ids = torch.randint(0, 800, (2, 32)).cuda()
    ids[0, 5] = DET                                            # tool token followed by ordinary tokens (insert form)
    assert int(ops.seq_index(ids, (DET,), (), EMB, 4, IMP, False).status.item()) & 1
    ids[0, 6:10] = EMB
    assert int(ops.seq_index(ids, (DET,), (), EMB, 4, IMP, False).status.item()) == 0
    ids[1, 30] = DET                                           # slots run past the row end
    assert int(ops.seq_index(ids, (DET,), (), EMB, 4, IMP, False).status.item()) & 1
    ids[1, 30] = 7
    ids[0, :12] = IMP                                          # 12 slots vs 2 tiles x 8 tokens
    assert int(ops.seq_index(ids, (DET,), (), EMB, 4, IMP, [2, 1], 8).status.item()) & 2


def test_composite_forward_fused_equals_torch_sequence_path():
    """The whole B200VisionLLMv2Model.forward with FUSED_SEQUENCE on / off: same ids, logits, text_query."""
    import test_modules_gpu as T  # its tiny Llama config
    import visionllm_b200.modeling as M
    from visionllm_b200.internvit import B200InternVisionModel, InternVisionConfig
    from visionllm_b200.llama import B200LlamaForCausalLM
    torch.manual_seed(0)
    vcfg = InternVisionConfig(hidden_size=256, num_attention_heads=2, num_hidden_layers=2, intermediate_size=512,
                              image_size=56, patch_size=14)
    lcfg = T.tiny_llama_cfg()
    I, E, D = 990, 991, 989
    cfg = SimpleNamespace(use_pixelshuffle=True, vl_bridge_type="internvl_mlp", vis_output_layer=-1, num_embs=4,
                          imp_token_id=I, emb_token_id=E, det_tool_id=D, seg_tool_id=-1, grd_tool_id=-1, pose_tool_id=-1)
    seen = {}

    class FakeGdino(torch.nn.Module):
        def forward(self, pixel_values, pixel_mask=None, text_query=None, text_query_masks=None, **kw):
            seen.update(tq=text_query, tm=text_query_masks)
            return SimpleNamespace(logits=None)

    model = M.B200VisionLLMv2Model(cfg, B200InternVisionModel(vcfg), B200LlamaForCausalLM(lcfg), gdino=FakeGdino())
    model = model.to("cuda", torch.bfloat16).eval()
    ids = torch.randint(0, 900, (2, 40)); ids[:, :4] = I
    ids[0, 20] = D; ids[0, 21:25] = E; ids[1, 10] = D; ids[1, 11:15] = E; ids[1, 30] = D; ids[1, 31:35] = E
    ids = ids.cuda()
    images = [torch.randn(1, 3, 56, 56).cuda().bfloat16() for _ in range(2)]
    aug = torch.randn(2, 3, 64, 64).cuda().bfloat16()
    kw = dict(input_ids=ids, attention_mask=torch.ones_like(ids), images=images, images_aug=aug, img_metas=[{"task": "det"}] * 2)
    a = model(**kw); tq_a, tm_a = seen["tq"], seen["tm"]
    M.FUSED_SEQUENCE = False
    try:
        b = model(**kw); tq_b, tm_b = seen["tq"], seen["tm"]
    finally:
        M.FUSED_SEQUENCE = True
    assert torch.equal(a.input_ids, b.input_ids) and torch.equal(tm_a, tm_b)
    # the fused pixel-shuffle LayerNorm sums the 1024-wide row in another order: logits agree to bf16 rounding noise
    assert (a.logits - b.logits).abs().max().item() <= 2e-2 * b.logits.abs().max().item()
    assert (tq_a.float() - tq_b.float()).abs().max().item() <= 2e-2 * tq_b.float().abs().max().item()
    # lm_head on requested rows == the same rows of the all-positions logits
    rows = torch.tensor([3, 39, 40 + 11, -1], device="cuda")
    c = model(logits_rows=rows, **kw)
    assert c.logits.shape == (4, lcfg.vocab_size) and c.logits.dtype == torch.float32
    assert torch.equal(c.logits, a.logits.reshape(-1, lcfg.vocab_size)[rows])


@pytest.mark.parametrize("g,Cc,with_ln", [(4, 64, True), (32, 3200, True), (16, 1024, False), (2, 8, True)])
def test_pixel_shuffle_rows_matches_reference_permutes(g, Cc, with_ln):
    from visionllm_b200 import ops
    from visionllm_b200.modeling import pixel_shuffle
    gen = torch.Generator(device="cuda").manual_seed(g)
    tiles = 3
    hs = torch.randn(tiles, 1 + g * g, Cc, device="cuda", generator=gen).bfloat16()
    ref = pixel_shuffle(hs[:, 1:].reshape(tiles, g, g, Cc), 0.5).reshape(tiles, -1, 4 * Cc)
    if not with_ln:
        assert torch.equal(ops.pixel_shuffle_rows(hs, 1), ref)
        return
    w = (1 + 0.1 * torch.randn(4 * Cc, device="cuda", generator=gen)).bfloat16()
    b = (0.1 * torch.randn(4 * Cc, device="cuda", generator=gen)).bfloat16()
    want = torch.nn.functional.layer_norm(ref.float(), (4 * Cc,), w.float(), b.float(), 1e-5)
    got = ops.pixel_shuffle_rows(hs, 1, w, b, 1e-5)
    assert ((got.float() - want).abs() <= 2.0 ** -8 * want.abs() + 1e-3).all()
    unfused = ops.layernorm(ref.contiguous(), w, b, 1e-5)
    assert ((got.float() - unfused.float()).abs() <= 2.0 ** -7 * unfused.float().abs() + 1e-3).all()


def test_gather_rows():
    from visionllm_b200 import ops
    src = torch.randn(50, 128, device="cuda").bfloat16()
    idx = torch.tensor([0, 49, 7, 7, -1, -50], device="cuda")
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    view = torch.randn(50, 256, device="cuda").bfloat16()[:, :128]           # row pitch != cols
    assert torch.equal(ops.gather_rows(view, idx), view[idx])
    # wide rows (>= 128 vectors) take the CTA-per-row kernel, narrow ones the flat vector-per-thread kernel; both are copies
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, C in ((3000, 96), (777, 8), (40, 1016), (33, 1024), (9, 4096), (100000, 192)):
        src = torch.randn(rows, C, device="cuda", generator=g).bfloat16()
        idx = torch.randint(-rows, rows, (rows * 2 + 5,), device="cuda", generator=g)
        assert torch.equal(ops.gather_rows(src, idx), src[idx]), (rows, C)
    assert ops.gather_rows(src, idx[:0]).shape == (0, 192)


def test_swin_patch_merging_order_and_layernorm():
    """order=1: HF SwinPatchMerging's cat([x(0::2,0::2), x(1::2,0::2), x(0::2,1::2), x(1::2,1::2)]) + LayerNorm(4C), on a
    non-square grid."""
    from visionllm_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(4)
    B, H, W, Cc = 2, 12, 20, 96
    x = torch.randn(B, H * W, Cc, device="cuda", generator=gen).bfloat16()
    x4 = x.view(B, H, W, Cc)
    ref = torch.cat([x4[:, 0::2, 0::2], x4[:, 1::2, 0::2], x4[:, 0::2, 1::2], x4[:, 1::2, 1::2]], -1).reshape(B, -1, 4 * Cc)
    assert torch.equal(ops.pixel_shuffle_rows(x, 0, grid=(H, W), order=1), ref)
    w = (1 + 0.1 * torch.randn(4 * Cc, device="cuda", generator=gen)).bfloat16()
    b = (0.1 * torch.randn(4 * Cc, device="cuda", generator=gen)).bfloat16()
    got = ops.pixel_shuffle_rows(x, 0, w, b, 1e-5, grid=(H, W), order=1)
    want = torch.nn.functional.layer_norm(ref.float(), (4 * Cc,), w.float(), b.float(), 1e-5)
    assert ((got.float() - want).abs() <= 2.0 ** -8 * want.abs() + 1e-3).all()

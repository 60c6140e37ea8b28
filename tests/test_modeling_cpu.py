"""CPU: the vectorised integer index logic of the composite forward equals the reference's python loops
(modeling_visionllmv2.py:447-468 [EMB] overwrite; :776-787 text_query gather; :381-392 pixel shuffle)."""
import torch

from visionllm_b200.modeling import emb_overwrite_indices, pixel_shuffle


def ref_emb_overwrite(ids, tools, emb_token_id, num_embs):
    """Loop restatement of mv2.py:447-468 (gap_len == num_embs)."""
    out = []
    emb_ids = torch.arange(emb_token_id, emb_token_id + num_embs)
    for cur in ids:
        pos = torch.cat([torch.where(cur == t)[0] for t in tools])
        new = cur
        for p in pos:
            new = torch.cat([new[: p + 1], emb_ids, new[p + num_embs + 1:]])
        out.append(new)
    return torch.stack(out)


def test_emb_overwrite_indices_match_reference_loop():
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 50, (4, 64), generator=g)
    DET, SEG, EMB = 100, 101, 200
    for b, p, t in ((0, 3, DET), (0, 20, SEG), (1, 59, DET), (3, 0, DET), (3, 30, SEG), (3, 40, DET)):
        ids[b, p] = t
        ids[b, p + 1:p + 5] = EMB
    ref = ref_emb_overwrite(ids, [DET, SEG], EMB, 4)
    bi, pi, j = emb_overwrite_indices(ids, [DET, SEG], 4)
    got = ids.clone()
    got[bi, pi] = EMB + j
    assert torch.equal(got, ref)


def test_pixel_shuffle_is_space_to_depth():
    x = torch.arange(2 * 4 * 6 * 8, dtype=torch.float32).view(2, 4, 6, 8)
    y = pixel_shuffle(x, 0.5)
    assert y.shape == (2, 2, 3, 32)
    # the reference's op sequence (view / permute / view / permute), element for element
    n, w, h, c = x.shape
    r = x.view(n, w, h // 2, c * 2).permute(0, 2, 1, 3).contiguous().view(n, h // 2, w // 2, c * 4).permute(0, 2, 1, 3)
    assert torch.equal(y, r.contiguous())


def test_vl_bridge_state_dict_keys_match_reference_layouts():
    """modeling_visionllmv2.py:162-184: 'linear' is a bare nn.Linear, the others nn.Sequential with GELU placeholders."""
    from visionllm_b200.modeling import build_vl_bridge
    assert sorted(build_vl_bridge("linear", 8, 16).state_dict()) == ["bias", "weight"]
    assert sorted(build_vl_bridge("mlp2x_gelu", 8, 16).state_dict()) == ["0.bias", "0.weight", "2.bias", "2.weight"]
    assert sorted(build_vl_bridge("internvl_mlp", 8, 16).state_dict()) == ["0.bias", "0.weight", "1.bias", "1.weight",
                                                                           "3.bias", "3.weight"]


def _views(tag, n):
    return torch.arange(n, dtype=torch.float32)[:, None, None, None] + torch.zeros(n, 3, 4, 4) + {"a": 0, "b": 100, "c": 200}[tag]


def test_region_encoder_inputs_pair_regions_with_the_global_view():
    """mv2.py:609-687: anyres -> last tile of the sample; pad -> the sample's image; mmic (num_splits) -> the last tile
    of the r-th image for the r-th region; features = patch tokens (CLS dropped) of that view, last three levels."""
    from visionllm_b200.modeling import region_encoder_inputs
    regions = [torch.ones(2, 4, 4), torch.zeros(0, 4, 4), torch.ones(1, 4, 4) * 2]
    # anyres: 3 samples with 3 / 1 / 2 tiles -> ViT batch rows 0-2 | 3 | 4-5
    images = [_views("a", 3), _views("b", 1), _views("c", 2)]
    hs = [torch.arange(6, dtype=torch.float32)[:, None, None].expand(6, 5, 2) + 10 * lv for lv in range(4)]
    ai, ar, af = region_encoder_inputs(images, regions, hs, [3, 1, 2])
    assert ai.shape == (3, 3, 4, 4) and ar.shape == (3, 1, 4, 4) and len(af) == 3
    assert ai[:, 0, 0, 0].tolist() == [2.0, 2.0, 201.0]                 # last tile of sample 0 twice, of sample 2 once
    assert ar[:, 0, 0, 0].tolist() == [1.0, 1.0, 2.0]
    for lv, f in enumerate(af):                                         # hidden_states[-3:] = levels 1, 2, 3
        assert f.shape == (3, 4, 2) and f[:, 0, 0].tolist() == [2 + 10 * (lv + 1), 2 + 10 * (lv + 1), 5 + 10 * (lv + 1)]
    # pad: one image per sample
    pad = torch.cat([_views("a", 1), _views("b", 1), _views("c", 1)])
    hs3 = [h[:3] for h in hs]
    ai, _, af = region_encoder_inputs(pad, regions, hs3, None)
    assert ai[:, 0, 0, 0].tolist() == [0.0, 0.0, 200.0] and af[0][:, 0, 0].tolist() == [10.0, 10.0, 12.0]
    # mmic: sample 0 holds two images of 2 + 1 tiles, its two regions go to image 0 and image 1
    ai, _, af = region_encoder_inputs([_views("a", 3), _views("b", 1), _views("c", 2)], regions, hs, [3, 1, 2],
                                      num_splits=[[2, 1], [1], [2]])
    assert ai[:, 0, 0, 0].tolist() == [1.0, 2.0, 201.0]
    assert af[2][:, 0, 0].tolist() == [31.0, 32.0, 35.0]


def test_scatter_region_tokens_in_order():
    from visionllm_b200.modeling import scatter_region_tokens
    ids = torch.tensor([[1, 9, 2, 9], [9, 3, 4, 5]])
    emb = torch.zeros(2, 4, 3)
    feats = torch.tensor([[1., 1, 1], [2, 2, 2], [3, 3, 3]])
    out = scatter_region_tokens(ids, emb, feats, 9)
    assert out[0, 1].tolist() == [1, 1, 1] and out[0, 3].tolist() == [2, 2, 2] and out[1, 0].tolist() == [3, 3, 3]
    assert out.sum().item() == 18 and emb.sum().item() == 0            # input untouched
    import pytest
    with pytest.raises(RuntimeError):
        scatter_region_tokens(ids, emb, feats[:2], 9)


def test_composite_forward_region_path_end_to_end(monkeypatch):
    """B200VisionLLMv2Model.forward with `regions`: the region encoder is called with the global view of every sample
    (anyres list input) and its outputs replace the `<region>` tokens in order, after the image tokens were scattered
    (mv2.py:582-698).  Sub-models are torch stand-ins; only the composite's own glue runs."""
    from types import SimpleNamespace
    import torch.nn as nn
    import torch.nn.functional as F
    import visionllm_b200.ops as ops
    from visionllm_b200.modeling import B200VisionLLMv2Model
    monkeypatch.setattr(ops, "linear", lambda x, w, bias=None, act=None, **kw: F.linear(x, w, bias))
    C = 8

    class FakeViT(nn.Module):
        config = SimpleNamespace(hidden_size=C, patch_size=14)

        def forward(self, x, output_hidden_states=True):
            n = x.shape[0]
            tok = x[:, 0, 0, 0].view(n, 1, 1).expand(n, 1 + 4, C)              # every token carries its tile's marker
            return SimpleNamespace(hidden_states=[tok + 0.1 * i for i in range(4)])

    class FakeLLM(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, vocab_size=50)
            self.emb = nn.Embedding(50, C)
            self.dtype = torch.float32

        def get_input_embeddings(self):
            return self.emb

        def forward(self, attention_mask=None, inputs_embeds=None, output_hidden_states=True):
            return SimpleNamespace(hidden_states=(inputs_embeds,), logits=None)

    seen = {}

    class FakeRegionEncoder(nn.Module):
        def forward(self, images, masks, feats, sample_points=None):
            seen.update(images=images, masks=masks, feats=feats)
            return masks.flatten(1).sum(1, keepdim=True).expand(-1, C) * 1000.0   # one marker row per region

    IMP, REG = 40, 41
    cfg = SimpleNamespace(use_pixelshuffle=False, vl_bridge_type="linear", vis_output_layer=-1, num_embs=4,
                          imp_token_id=IMP, emb_token_id=45, det_tool_id=-1, seg_tool_id=-1, grd_tool_id=-1,
                          pose_tool_id=-1, reg_token_id=REG)
    m = B200VisionLLMv2Model(cfg, FakeViT(), FakeLLM(), region_encoder=FakeRegionEncoder()).eval()
    with torch.no_grad():
        m.vl_bridge.weight.copy_(torch.eye(C)); m.vl_bridge.bias.zero_()
    ids = torch.randint(0, 30, (2, 12))
    ids[0, :8] = IMP                                          # sample 0: 2 tiles x 4 tokens
    ids[1, :4] = IMP                                          # sample 1: 1 tile
    ids[0, 9] = REG; ids[0, 11] = REG; ids[1, 6] = REG
    images = [torch.full((2, 3, 4, 4), 5.0), torch.full((1, 3, 4, 4), 7.0)]
    images[0][1] = 6.0                                        # sample 0's LAST tile (the global view) is marked 6
    regions = [torch.ones(2, 4, 4), torch.ones(1, 4, 4)]
    regions[0][1, :2] = 0                                     # second region of sample 0 covers 8 pixels
    out = m(input_ids=ids, images=images, regions=regions)
    assert seen["images"][:, 0, 0, 0].tolist() == [6.0, 6.0, 7.0]
    want = torch.tensor([[6.1, 6.1, 7.1], [6.2, 6.2, 7.2], [6.3, 6.3, 7.3]])
    assert torch.allclose(torch.stack([f[:, 0, 0] for f in seen["feats"]]), want) and seen["feats"][0].shape == (3, 4, C)
    h = out.last_hidden_state
    assert h[0, 9, 0].item() == 16000.0 and h[0, 11, 0].item() == 8000.0 and h[1, 6, 0].item() == 16000.0
    assert torch.allclose(h[0, :4], torch.full((4, C), 5.3)) and torch.allclose(h[0, 4:8], torch.full((4, C), 6.3))
    assert torch.allclose(h[1, :4], torch.full((4, C), 7.3))


def _ref_nested_tensor(tensor_list, size_divisibility):
    """util/misc.py:288-316 restated as the reference's loop (split into 3-channel images, zero-pad to the max size
    rounded up to the divisibility)."""
    tensor_list = [piece for t in tensor_list for piece in t.split(3, dim=0)]
    max_size = [max(s) for s in zip(*[list(img.shape) for img in tensor_list])]
    max_size[-2] = (max_size[-2] + size_divisibility - 1) // size_divisibility * size_divisibility
    max_size[-1] = (max_size[-1] + size_divisibility - 1) // size_divisibility * size_divisibility
    tensor = torch.zeros([len(tensor_list)] + max_size, dtype=tensor_list[0].dtype)
    for img, pad_img in zip(tensor_list, tensor):
        pad_img[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
    return tensor


def test_pad_images_aug_matches_reference_nested_tensor():
    """ragged, non-/32 images_aug must reach the GDINO stage zero-padded like mv2.py:771."""
    from visionllm_b200.modeling import pad_images_aug
    g = torch.Generator().manual_seed(0)
    ragged = [torch.randn(3, 50, 67, generator=g), torch.randn(3, 64, 40, generator=g), torch.randn(3, 33, 96, generator=g)]
    got = pad_images_aug(ragged, 32)
    ref = _ref_nested_tensor(ragged, 32)
    assert got.shape == (3, 3, 64, 96) and torch.equal(got, ref)
    same = torch.randn(2, 3, 64, 96, generator=g)              # already aligned: a plain stack, values untouched
    assert torch.equal(pad_images_aug(same, 32), same) and torch.equal(pad_images_aug(list(same), 32), same)
    odd = torch.randn(2, 3, 50, 70, generator=g)                # a tensor whose H/W is not a multiple of 32
    assert torch.equal(pad_images_aug(odd, 32), _ref_nested_tensor(list(odd), 32))
    mask = pad_images_aug(ragged, 32)[:, 0] != 0                # mv2.py:773: the padding is what the mask sees
    assert not mask[0, 50:, :].any() and not mask[0, :, 67:].any() and mask[0, :50, :67].all()


def _tiny_composite(gdino=None, **ids):
    from types import SimpleNamespace
    import torch.nn as nn
    from visionllm_b200.modeling import B200VisionLLMv2Model
    C = 8

    class FakeViT(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, patch_size=2)

        def forward(self, x, output_hidden_states=True):
            t = torch.zeros(x.shape[0], 5, C)
            return SimpleNamespace(hidden_states=(t, t, t))

    class FakeLLM(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, vocab_size=64)
            self.emb = nn.Embedding(64, C)
            self.dtype = torch.float32

        def get_input_embeddings(self):
            return self.emb

        def forward(self, attention_mask=None, inputs_embeds=None, output_hidden_states=True):
            return SimpleNamespace(hidden_states=(inputs_embeds,), logits=None)

    cfg = SimpleNamespace(use_pixelshuffle=False, vl_bridge_type="linear", vis_output_layer=-1, num_embs=4,
                          imp_token_id=40, emb_token_id=45, det_tool_id=50, seg_tool_id=-1, grd_tool_id=-1,
                          pose_tool_id=51)
    return B200VisionLLMv2Model(cfg, FakeViT(), FakeLLM(), gdino=gdino).eval()


def test_inject_emb_refuses_to_clobber_real_tokens():
    """a tool token whose following slots are NOT pre-placed [EMB] ids is the reference's insert form
    (mv2.py:428-431, gap_len = 0); the overwrite form must not silently eat the next num_embs tokens."""
    import pytest
    m = _tiny_composite()
    ids = torch.randint(0, 30, (1, 16))
    ids[0, 3] = 50                                             # [DET] followed by ordinary tokens, no [EMB] id in row 0:
    out = m(input_ids=ids)                                     # the reference's INSERT form (gap_len = 0), see the test below
    assert out.input_ids.shape == (1, 20) and out.input_ids[0, 4:8].tolist() == [45, 46, 47, 48]
    ids[0, 4:8] = 45                                           # collator form: [EMB] x 4 placeholders after the tool
    out = m(input_ids=ids)
    assert out.input_ids.shape == (1, 16) and out.input_ids[0, 4:8].tolist() == [45, 46, 47, 48]
    ids2 = ids.clone(); ids2[0, 15] = 50                       # [EMB] ids present (overwrite form) but a tool token at the very
    with pytest.raises(NotImplementedError):                   # end has no slots: the reference would clobber / misalign
        m(input_ids=ids2)
    ids3 = ids.clone(); ids3[0, 10] = 50                       # ... or is followed by ordinary tokens
    with pytest.raises(NotImplementedError):
        m(input_ids=ids3)


def test_gdino_task_gate_follows_reference():
    """mv2.py:755-770: gdino runs only for det/det_cap/grd/seg/count_*/interactive/ic_mask; 'pose' is not wired."""
    import pytest
    from types import SimpleNamespace
    import torch.nn as nn
    calls = []

    class FakeGdino(nn.Module):
        def forward(self, pixel_values, pixel_mask=None, text_query=None, text_query_masks=None, **kw):
            calls.append(tuple(pixel_values.shape))
            return SimpleNamespace(logits=None)

    m = _tiny_composite(gdino=FakeGdino())
    ids = torch.randint(0, 30, (1, 16)); ids[0, 3] = 50; ids[0, 4:8] = 45
    aug = [torch.ones(3, 40, 50)]
    assert m(input_ids=ids, images_aug=aug).gdino_outputs is None and not calls                # no img_metas -> task None
    assert m(input_ids=ids, images_aug=aug, img_metas=[{"task": "vqa"}]).gdino_outputs is None and not calls
    assert m(input_ids=ids, images_aug=aug, img_metas=[{"task": "seg"}]).gdino_outputs is not None
    assert calls == [(1, 3, 64, 64)]                                                            # padded to /32
    with pytest.raises(NotImplementedError):
        m(input_ids=ids, images_aug=aug, img_metas=[{"task": "pose"}])


def _tiny_composite_with_head(unipose=None):
    """_tiny_composite whose LLM stand-in also has an lm_head (fp32 logits), for the loss / pose paths."""
    from types import SimpleNamespace
    import torch.nn as nn
    m = _tiny_composite()
    C, V = 8, 64

    class FakeLLM(nn.Module):
        def __init__(self):
            super().__init__()
            self.config = SimpleNamespace(hidden_size=C, vocab_size=V)
            self.emb = nn.Embedding(V, C)
            self.head = nn.Linear(C, V, bias=False)
            self.dtype = torch.float32

        def get_input_embeddings(self):
            return self.emb

        def forward(self, attention_mask=None, inputs_embeds=None, output_hidden_states=True):
            h = torch.tanh(inputs_embeds)
            return SimpleNamespace(hidden_states=(inputs_embeds, h), logits=self.head(h).float())

    from visionllm_b200.modeling import B200VisionLLMv2Model
    torch.manual_seed(3)
    return B200VisionLLMv2Model(m.config, m.vis_encoder, FakeLLM(), unipose=unipose).eval()


def test_labels_give_the_reference_loss(monkeypatch):
    """mv2.py:740-757: [EMB] labels -> IGNORE_INDEX (in place, like the reference), shift, flatten, CrossEntropyLoss()."""
    import torch.nn.functional as F
    import visionllm_b200.ops as ops
    from oracle import torch_kernels as TK
    monkeypatch.setattr(ops, "ce_loss", TK.ce_loss)
    m = _tiny_composite_with_head()
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 30, (2, 16), generator=g)
    ids[0, 3] = 50; ids[0, 4:8] = 45                                                    # [DET] + 4 [EMB] slots
    labels = ids.clone()
    labels[:, :3] = -100                                                                  # prompt positions
    labels[0, 4:8] = torch.tensor([45, 46, 47, 48])                                       # collator leaves the [EMB] ids in labels
    labels_in = labels.clone()
    out = m(input_ids=ids, labels=labels)
    # the reference's lines, literally
    ref_labels = labels_in.clone()
    ref_labels[(ref_labels >= 45) & (ref_labels <= 48)] = -100
    shift_logits = out.logits[..., :-1, :].contiguous().view(-1, 64)
    shift_labels = ref_labels[..., 1:].contiguous().view(-1)
    ref = F.cross_entropy(shift_logits, shift_labels)
    assert out.loss.dtype == torch.float32 and out.loss.dim() == 0
    assert abs(out.loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert torch.equal(labels, ref_labels)                                                # caller's labels masked in place
    tup = m(input_ids=ids, labels=labels_in.clone(), return_dict=False)                   # mv2.py:873-875: (loss,) + output
    assert tup[0].item() == out.loss.item() and tup[1] is not None and len(tup) == 5
    assert len(m(input_ids=ids, return_dict=False)) == 4
    import pytest
    with pytest.raises(NotImplementedError):
        m(input_ids=ids, targets=[{}])
    with pytest.raises(ValueError):
        m(input_ids=ids, labels=labels_in.clone(), logits_rows=torch.tensor([0, 1]))


def test_pad_images_aug_mask_matches_reference_nested_tensor():
    from visionllm_b200.modeling import pad_images_aug
    g = torch.Generator().manual_seed(1)
    ragged = [torch.randn(3, 50, 67, generator=g), torch.randn(6, 64, 40, generator=g)]       # a 6-channel entry splits in two
    t, mask = pad_images_aug(ragged, 32, return_mask=True)
    assert torch.equal(t, _ref_nested_tensor(ragged, 32)) and mask.shape == (3, 64, 96) and mask.dtype == torch.bool
    ref_mask = torch.ones(3, 64, 96, dtype=torch.bool)                                         # util/misc.py:310-313
    ref_mask[0, :50, :67] = False; ref_mask[1, :64, :40] = False; ref_mask[2, :64, :40] = False
    assert torch.equal(mask, ref_mask)
    same = torch.randn(2, 3, 64, 96, generator=g)
    t2, m2 = pad_images_aug(same, 32, return_mask=True)
    assert torch.equal(t2, same) and not m2.any()


def test_pose_branch_routes_emb_states_to_unipose():
    """mv2.py:795-836: for task 'pose' the per-sample [EMB] patches split into the first len(id2index) object-class patches
    and the keypoint patches (100 zero-padded slots each), and UniPose gets the /32-padded NestedTensor with its own mask."""
    import torch.nn as nn
    from types import SimpleNamespace
    seen = {}

    class FakeUniPose(nn.Module):
        def forward_samples(self, tensors, mask, text_query):
            seen.update(tensors=tensors, mask=mask, tq=text_query)
            return SimpleNamespace(pred_boxes="boxes")

    m = _tiny_composite_with_head(unipose=FakeUniPose())
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, 30, (3, 40), generator=g)
    # sample 0: 1 object class + 3 keypoint classes; sample 1: 2 + 2; sample 2: object classes only (skipped by the reference)
    n_patch, n_obj = [4, 4, 2], [1, 2, 2]
    for b, n in enumerate(n_patch):
        for j in range(n):
            p = 2 + 5 * j
            ids[b, p] = 51                                                                # [POSE] tool token
            ids[b, p + 1:p + 5] = 45                                                      # its 4 [EMB] slots
    aug = [torch.randn(3, 40, 50, generator=g), torch.randn(3, 33, 64, generator=g), torch.randn(3, 20, 20, generator=g)]
    metas = [{"task": "pose", "id2index": {i: i for i in range(n)}} for n in n_obj]
    out = m(input_ids=ids, images_aug=aug, img_metas=metas)
    assert out.unipose_outputs.pred_boxes == "boxes" and out.gdino_outputs is None
    # the reference's loop, literally (mv2.py:801-823), on the same hidden states / rewritten ids
    hidden, new_ids = out.last_hidden_state, out.input_ids
    emb_select = (new_ids >= 45) & (new_ids <= 48)
    num_patches = emb_select.sum(-1) // 4
    obj = torch.zeros(3, 100, 4, 8); objm = torch.zeros(3, 100, dtype=torch.bool)
    kpt = torch.zeros(3, 100, 4, 8); kptm = torch.zeros(3, 100, dtype=torch.bool)
    for b in range(3):
        num_objcls = len(metas[b]["id2index"])
        num_kpts = int(num_patches[b]) - num_objcls
        if num_objcls != 0 and num_kpts != 0:
            tq_i = hidden[b, emb_select[b], :].reshape(-1, 4, 8)
            obj[b, :num_objcls] = tq_i[:num_objcls]; objm[b, :num_objcls] = 1
            kpt[b, :num_kpts] = tq_i[num_objcls:]; kptm[b, :num_kpts] = 1
    tq = seen["tq"]
    assert torch.equal(tq["obj_querys"], obj) and torch.equal(tq["obj_query_masks"], objm)
    assert torch.equal(tq["kpt_querys"], kpt) and torch.equal(tq["kpt_query_masks"], kptm)
    assert not objm[2].any() and objm[0].sum() == 1 and kptm[0].sum() == 3
    assert torch.equal(seen["tensors"], _ref_nested_tensor(aug, 32)) and seen["tensors"].shape == (3, 3, 64, 64)
    assert seen["mask"][0, :40, :50].logical_not().all() and seen["mask"][0, 40:].all() and seen["mask"][0, :, 50:].all()
    # no [EMB] tokens at all: UniPose is not called
    seen.clear()
    plain = torch.randint(0, 30, (1, 12), generator=g)
    assert m(input_ids=plain, images_aug=aug[:1], img_metas=metas[:1]).unipose_outputs is None and not seen
    # other tasks never reach it
    assert m(input_ids=ids, images_aug=aug, img_metas=[{"task": "det"}] * 3).unipose_outputs is None and not seen


def _ref_insert_loop(input_ids, inputs_embeds, det_ids, pose_id, emb_token_id, num_embs, emb_det, emb_pose):
    """Pure-python oracle of mv2.py:436-527 with gap_len = 0: python-list splices at `position + 1`, positions read from the
    ORIGINAL row (first the det / seg / grd hits concatenated in that order, then the pose hits), never shifted."""
    out_ids, out_emb = [], []
    for row_ids, row_emb in zip(input_ids.tolist(), inputs_embeds):
        ids, emb = list(row_ids), [e for e in row_emb]
        hits_det = [i for t in det_ids for i, v in enumerate(row_ids) if v == t]
        hits_pose = [i for i, v in enumerate(row_ids) if v == pose_id]
        for hits, table in ((hits_det, emb_det), (hits_pose, emb_pose)):
            for p0 in hits:
                ids[p0 + 1:p0 + 1] = list(range(emb_token_id, emb_token_id + num_embs))
                emb[p0 + 1:p0 + 1] = [table[j] for j in range(num_embs)]
        out_ids.append(torch.tensor(ids, dtype=torch.long))
        out_emb.append(torch.stack(emb))
    return torch.stack(out_ids), torch.stack(out_emb)


def test_insert_form_matches_the_reference_loop():
    """gap_len == 0 (mv2.py:428-429: a tool token in row 0 and no [EMB] id there -- generation, multi-round chat): [EMB] ids and
    emb_embeddings are inserted after every tool token, with the reference's unshifted positions; the attention mask grows by
    ones (mv2.py:539-545)."""
    import pytest
    m = _tiny_composite()
    g = torch.Generator().manual_seed(5)
    emb_det, emb_pose = m.emb_embeddings_det.weight.detach(), m.emb_embeddings_pose.weight.detach()
    cases = []
    a = torch.randint(0, 30, (1, 12), generator=g); a[0, 11] = 50                       # generation: [DET] is the last token
    cases.append(a)
    b = torch.randint(0, 30, (2, 14), generator=g); b[0, 3] = 50; b[0, 9] = 50; b[1, 2] = 50; b[1, 12] = 50   # two per row
    cases.append(b)
    c = torch.randint(0, 30, (2, 14), generator=g); c[0, 8] = 51; c[0, 2] = 50; c[1, 5] = 50; c[1, 6] = 51    # det + pose
    cases.append(c)
    for ids in cases:
        emb = m.llm.get_input_embeddings()(ids).detach()
        want_ids, want_emb = _ref_insert_loop(ids, emb, [50], 51, 45, 4, emb_det, emb_pose)
        assert m.uses_insert_form(ids)
        got_ids, got_emb = m.inject_emb(ids, emb)
        assert torch.equal(got_ids, want_ids) and torch.equal(got_emb, want_emb)
        am = torch.ones_like(ids); am[:, :2] = 0
        out = m(input_ids=ids, attention_mask=am)
        assert torch.equal(out.input_ids, want_ids) and torch.equal(out.last_hidden_state, want_emb)
    ragged = torch.randint(0, 30, (2, 10), generator=g); ragged[0, 3] = 50; ragged[1, 3] = 50; ragged[1, 7] = 50
    with pytest.raises(RuntimeError):
        m(input_ids=ragged)
    # row 0 decides (mv2.py:425-431): [EMB] ids in row 0 -> overwrite form -> a slot-less tool token elsewhere is refused
    mixed = torch.randint(0, 30, (2, 12), generator=g); mixed[0, 2] = 50; mixed[0, 3:7] = 45; mixed[1, 4] = 50
    assert not m.uses_insert_form(mixed)
    with pytest.raises(NotImplementedError):
        m(input_ids=mixed)

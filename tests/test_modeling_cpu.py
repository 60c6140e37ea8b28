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

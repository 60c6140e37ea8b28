"""CPU: the host-side post-processors that follow the UniPose and Grounding-DINO paths (`unipose.post_process_pose`,
`gdino_heads.post_process_sem_seg`) against independent oracles of the reference's functions
(visionllmv2/eval/eval_pose.py:19-86, eval_semseg.py:16-62: one detection / one class map at a time, python loops) on the same
random head outputs."""
import pytest
import torch
import torch.nn.functional as F


def _ref_post_process_pose(out_logits, out_bbox, out_keypoints, target_sizes, num_classes=1, topk=100, num_body_points=17,
                           id_mapping=None, threshold=0.):
    """Pure-python oracle of eval_pose.py:19-86, one detection at a time: rank the (query, class) sigmoid scores, then for each
    kept pair build the label through id_mapping, the xyxy box scaled by (w, h, w, h) and the (x * w, y * h, 1) keypoint triples
    of the first `num_body_points` points of the query's xyxy..zz keypoint vector."""
    results = []
    for b in range(out_logits.shape[0]):
        K = num_classes
        prob = out_logits[b, :, :K].sigmoid().reshape(-1)
        vals, order = torch.topk(prob, min(topk, prob.numel()))
        h, w = float(target_sizes[b][0]), float(target_sizes[b][1])
        scores, labels, boxes, kpts = [], [], [], []
        for v, flat in zip(vals, order.tolist()):
            if not v > threshold:
                continue
            q, c = flat // K, flat % K
            cx, cy, bw, bh = out_bbox[b, q]
            xy = out_keypoints[b, q, :num_body_points * 2]
            trip = []
            for j in range(num_body_points):
                trip += [xy[2 * j] * w, xy[2 * j + 1] * h, torch.tensor(1.0)]
            scores.append(v); labels.append(id_mapping[c])
            boxes.append(torch.stack([(cx - 0.5 * bw) * w, (cy - 0.5 * bh) * h, (cx + 0.5 * bw) * w, (cy + 0.5 * bh) * h]))
            kpts.append(torch.stack(trip))
        n = len(scores)
        results.append({"scores": torch.stack(scores) if n else torch.zeros(0),
                        "labels": torch.tensor(labels, dtype=torch.long),
                        "boxes": torch.stack(boxes) if n else torch.zeros(0, 4),
                        "keypoints": torch.stack(kpts) if n else torch.zeros(0, num_body_points * 3)})
    return results


@pytest.mark.parametrize("num_classes,nbp,topk,threshold", [(1, 17, 100, 0.0), (3, 17, 20, 0.3), (2, 21, 500, 0.0)])
def test_post_process_pose_matches_reference(num_classes, nbp, topk, threshold):
    from visionllm_b200.unipose import post_process_pose
    g = torch.Generator().manual_seed(num_classes * 10 + nbp)
    bs, nq = 3, 50
    logits = torch.randn(bs, nq, 100, generator=g)
    logits[:, :, num_classes:] = float("-inf")                        # padded class slots, as ContrastiveAssign leaves them
    boxes = torch.rand(bs, nq, 4, generator=g)
    kpts = torch.rand(bs, nq, 68 * 3, generator=g)
    sizes = torch.tensor([[480., 640.], [333., 500.], [1024., 768.]])
    id_mapping = {i: 7 * i + 1 for i in range(num_classes)}
    want = _ref_post_process_pose(logits, boxes, kpts, sizes, num_classes, topk, nbp, id_mapping, threshold)
    got = post_process_pose(logits, boxes, kpts, sizes, num_classes, topk, nbp, id_mapping, threshold)
    got_list = post_process_pose(logits, boxes, kpts, [tuple(s.tolist()) for s in sizes], num_classes, topk, nbp, id_mapping,
                                 threshold)
    assert len(got) == len(want) == bs
    for a, b, c in zip(got, want, got_list):
        for k in ("scores", "labels", "boxes", "keypoints"):
            assert torch.equal(a[k], b[k]), k
            assert torch.equal(c[k], b[k]), k
        assert a["keypoints"].shape[-1] == nbp * 3
    with pytest.raises(ValueError):
        post_process_pose(logits, boxes, kpts, sizes, num_classes, topk, nbp, None, threshold)
    with pytest.raises(ValueError):
        post_process_pose(logits, boxes, kpts, sizes[:2], num_classes, topk, nbp, id_mapping, threshold)


def _resize_chain(x, image_size, target_size):
    """[n, h/4, w/4] -> x4 bilinear -> crop of the batch padding -> bilinear to the original size (eval_semseg.py:22-25 / :31-34)."""
    n, h, w = x.shape
    x = F.interpolate(x.unsqueeze(1), size=(4 * h, 4 * w), mode="bilinear", align_corners=False)
    x = x[..., :image_size[0], :image_size[1]]
    return F.interpolate(x, size=tuple(target_size[:2]), mode="bilinear", align_corners=False).squeeze(1)


def _ref_process_seg_result(mask_cls, mask_pred, image_size, target_size, before=True):
    """Oracle of eval_semseg.py:16-37: class-weighted sum of the query mask probabilities, per pixel arg-max; the weighting happens
    after the resize chain (`before`) or before it."""
    prob, m = torch.sigmoid(mask_cls), torch.sigmoid(mask_pred)
    if before:
        m = _resize_chain(m, image_size, target_size)
        per_class = (prob.t()[:, :, None, None] * m[None]).sum(1) if m.numel() < 2e6 else torch.einsum("qc,qhw->chw", prob, m)
    else:
        per_class = _resize_chain(torch.einsum("qc,qhw->chw", prob, m), image_size, target_size)
    return per_class.argmax(0)


@pytest.mark.parametrize("before", [True, False])
def test_post_process_sem_seg_matches_reference(before):
    from visionllm_b200.gdino_heads import post_process_sem_seg
    g = torch.Generator().manual_seed(3)
    nq, K, num_classes = 30, 24, 19
    logits = torch.randn(2, nq, K, generator=g)
    masks = torch.randn(2, nq, 16, 24, generator=g) * 3
    image_sizes, target_sizes = [(60, 90), (64, 96)], [(120, 200, 3), (97, 131)]
    got = post_process_sem_seg(logits, masks, target_sizes, image_sizes, num_classes=num_classes,
                               sem_seg_postprocess_before_inference=before)
    for i in range(2):
        want = _ref_process_seg_result(logits[i][..., :num_classes], masks[i], image_sizes[i], target_sizes[i], before)
        assert got[i].dtype == torch.int64 and got[i].shape == tuple(target_sizes[i][:2])
        assert torch.equal(got[i], want)
        assert int(got[i].max()) < num_classes
    with pytest.raises(ValueError):
        post_process_sem_seg(logits, masks, target_sizes[:1], image_sizes)

"""CPU: the host-side post-processors that follow the UniPose and Grounding-DINO paths (`unipose.post_process_pose`,
`gdino_heads.post_process_sem_seg`) against statement-for-statement transcriptions of the reference's functions
(visionllmv2/eval/eval_pose.py:19-86, eval_semseg.py:16-62) on the same random head outputs."""
import pytest
import torch
import torch.nn.functional as F


def _box_cxcywh_to_xyxy(x):
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([(x_c - 0.5 * w), (y_c - 0.5 * h), (x_c + 0.5 * w), (y_c + 0.5 * h)], dim=-1)


def _ref_post_process_pose(out_logits, out_bbox, out_keypoints, target_sizes, num_classes=1, topk=100, num_body_points=17,
                           id_mapping=None, threshold=0.):
    out_logits = out_logits[:, :, :num_classes]
    prob = out_logits.sigmoid()
    prob = prob.view(out_logits.shape[0], -1)
    k_value = min(topk, prob.size(1))
    topk_values, topk_indexes = torch.topk(prob, k_value, dim=1)
    scores = topk_values
    topk_boxes = torch.div(topk_indexes, out_logits.shape[2], rounding_mode="floor")
    labels = topk_indexes % out_logits.shape[2]
    new_labels = torch.zeros_like(labels)
    for batch_idx in range(len(labels)):
        for j in range(labels.shape[-1]):
            new_labels[batch_idx, j] = id_mapping[labels[batch_idx, j].item()]
    labels = new_labels
    boxes = _box_cxcywh_to_xyxy(out_bbox)
    boxes = torch.gather(boxes, 1, topk_boxes.unsqueeze(-1).repeat(1, 1, 4))
    img_h, img_w = target_sizes.unbind(1)
    scale_fct = torch.stack([img_w, img_h, img_w, img_h], dim=1).to(boxes.device)
    boxes = boxes * scale_fct[:, None, :]
    topk_keypoints = torch.div(topk_indexes, out_logits.shape[2], rounding_mode="floor")
    keypoints = torch.gather(out_keypoints, 1, topk_keypoints.unsqueeze(-1).repeat(1, 1, 68 * 3))
    Z_pred = keypoints[:, :, :(num_body_points * 2)]
    V_pred = torch.ones_like(Z_pred)[:, :, :num_body_points]
    img_h, img_w = target_sizes.unbind(1)
    scale_fct = torch.stack([img_w, img_h], dim=1).repeat(1, num_body_points)[:, None, :].to(Z_pred.device)
    Z_pred = Z_pred * scale_fct
    keypoints = torch.cat([Z_pred, V_pred], dim=-1)
    keypoints_res = torch.zeros_like(keypoints)
    keypoints_res[..., 0::3] = Z_pred[..., 0::2]
    keypoints_res[..., 1::3] = Z_pred[..., 1::2]
    keypoints_res[..., 2::3] = V_pred[..., 0::1]
    results = []
    for s, l, b, k in zip(scores, labels, boxes, keypoints_res):
        results.append({"scores": s[s > threshold], "labels": l[s > threshold], "boxes": b[s > threshold],
                        "keypoints": k[s > threshold]})
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


def _ref_process_seg_result(mask_cls, mask_pred, image_size, target_size, before=True):
    prob = mask_cls.sigmoid()
    mask_pred = mask_pred.sigmoid()
    H, W = mask_pred.shape[-2:]
    if before:
        mask_pred = F.interpolate(mask_pred[:, None], size=(H * 4, W * 4), mode="bilinear", align_corners=False)
        mask_pred = mask_pred[:, :, :image_size[0], :image_size[1]]
        mask_pred = F.interpolate(mask_pred, size=target_size[:2], mode="bilinear", align_corners=False)[:, 0]
        semseg = torch.einsum('qc,qhw->chw', prob, mask_pred)
        semantic_map = semseg.argmax(dim=0)
    else:
        mask_pred = torch.einsum('qc,qhw->chw', prob, mask_pred)
        mask_pred = F.interpolate(mask_pred[:, None], size=(H * 4, W * 4), mode="bilinear", align_corners=False)
        mask_pred = mask_pred[:, :, :image_size[0], :image_size[1]].cpu()
        mask_pred = F.interpolate(mask_pred, size=target_size[:2], mode="bilinear", align_corners=False)[:, 0]
        semantic_map = mask_pred.argmax(dim=0)
    return semantic_map


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

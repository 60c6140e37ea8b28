"""Grounding-DINO query-selection, heads and post-processing of the region/mask decoder (SURVEY 8a a12, a18, a20,
a21), restated from visionllmv2/model/grounding_dino/modeling_ov_grounding_dino_mask_dn.py and
visionllmv2/eval/eval_det.py.  Dense math runs on the tcgen05 GEMM; the integer work (top-k, //, %, gathers)
uses the same torch primitives the reference uses so indices -- including tie order -- are identical.

  GroundingDinoMLPPredictionHead        gd.py:3704-3720   (bbox / mask-embed / patch2query MLPs, ReLU epilogues)
  GroundingDinoContrastiveEmbedding     gd.py:1410-1428   (q @ text^T, -inf padding to max_text_len, fp32)
  EncoderOutputProposals                gd.py:2228-2276   (per-pixel proposals, validity, inverse sigmoid, enc_output+LN)
  select_topk_proposals                 gd.py:2503-2534   (two-stage top-k query selection)
  forward_seg_heads                     gd.py:2278-2281   (mask einsum 'bqc,bchw->bqhw' as one GEMM per image)
  patch2query_mean                      gd.py:3138        (MLP over the 4 [EMB] states, mean over them)
  post_process_det_gdino / _instseg_    eval_det.py:18-104 (sigmoid, top-k over Q*K, idx//K, idx%K, box scale, masks)
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import msda, ops


class GroundingDinoMLPPredictionHead(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    @torch.no_grad()
    def forward(self, x):
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            if last and layer.out_features % 8:
                # bbox head: 4 outputs -- pad the weight rows to 8 so the output pitch is 16 bytes
                n = layer.out_features
                w = torch.zeros((8, layer.in_features), dtype=layer.weight.dtype, device=layer.weight.device)
                b = torch.zeros((8,), dtype=layer.bias.dtype, device=layer.bias.device)
                w[:n], b[:n] = layer.weight, layer.bias
                x = ops.linear(x, w, bias=b)[..., :n]
            else:
                x = ops.linear(x, layer.weight, bias=layer.bias, act=None if last else "relu")
        return x


class GroundingDinoContrastiveEmbedding(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.max_text_len = config.max_text_len

    @torch.no_grad()
    def forward(self, vision_hidden_state, text_hidden_state, text_token_mask):
        B, Q, C = vision_hidden_state.shape
        T = text_hidden_state.shape[1]
        Tp = (T + 7) // 8 * 8
        out = torch.full((B, Q, self.max_text_len), float("-inf"), device=vision_hidden_state.device)   # fp32
        buf = torch.empty((B, Q, Tp), dtype=vision_hidden_state.dtype, device=vision_hidden_state.device)
        for b in range(B):
            ops.linear(vision_hidden_state[b], text_hidden_state[b].contiguous(), out=buf[b, :, :T])
        logits = buf[..., :T].masked_fill(~text_token_mask[:, None, :], float("-inf"))
        out[..., :T] = logits
        return out


@torch.no_grad()
def gen_encoder_output_proposals(enc_output_linear, enc_output_norm, enc_output, padding_mask, spatial_shapes):
    """OVGroundingDinoModel.gen_encoder_output_proposals (gd.py:2228-2276): one (cx, cy, w, h) proposal per pixel,
    validity window (0.01, 0.99), inverse sigmoid with +inf for padded / invalid pixels, zeroed features through
    enc_output + enc_output_norm."""
    B = enc_output.shape[0]
    dev = enc_output.device
    proposals, pos = [], 0
    for level, (H, W) in enumerate(msda.host_shape_list(spatial_shapes)):
        m = padding_mask[:, pos:pos + H * W].view(B, H, W)
        valid_h = (~m[:, :, 0]).sum(1)
        valid_w = (~m[:, 0, :]).sum(1)
        gy, gx = torch.meshgrid(torch.linspace(0, H - 1, H, dtype=torch.float32, device=dev),
                                torch.linspace(0, W - 1, W, dtype=torch.float32, device=dev), indexing="ij")
        grid = torch.stack((gx, gy), -1)[None].expand(B, -1, -1, -1)
        scale = torch.stack((valid_w, valid_h), 1).view(B, 1, 1, 2)
        grid = (grid + 0.5) / scale
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** level)
        proposals.append(torch.cat((grid, wh), -1).view(B, -1, 4))
        pos += H * W
    prop = torch.cat(proposals, 1)
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    prop = torch.log(prop / (1 - prop))
    prop = prop.masked_fill(padding_mask.unsqueeze(-1), float("inf")).masked_fill(~valid, float("inf"))
    q = enc_output.masked_fill(padding_mask.unsqueeze(-1), 0.0).masked_fill(~valid, 0.0)
    q = ops.linear(q, enc_output_linear.weight, bias=enc_output_linear.bias)
    q = ops.layernorm(q, enc_output_norm.weight, enc_output_norm.bias, enc_output_norm.eps)
    return q, prop


class EncoderOutputProposals(nn.Module):
    """gen_encoder_output_proposals + enc_output / enc_output_norm (same parameter names as OVGroundingDinoModel)."""

    def __init__(self, d_model):
        super().__init__()
        self.enc_output = nn.Linear(d_model, d_model)
        self.enc_output_norm = nn.LayerNorm(d_model)

    def forward(self, enc_output, padding_mask, spatial_shapes):
        return gen_encoder_output_proposals(self.enc_output, self.enc_output_norm, enc_output, padding_mask, spatial_shapes)


@torch.no_grad()
def select_topk_proposals(enc_outputs_class, enc_outputs_coord_logits, object_query_embedding, num_queries):
    """Two-stage query selection: indices from torch.topk on the per-pixel max class logit (gd.py:2521-2534)."""
    topk_logits = enc_outputs_class.max(-1)[0]
    topk_proposals = torch.topk(topk_logits, num_queries, dim=1)[1]
    coords = torch.gather(enc_outputs_coord_logits, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4))
    cls = torch.gather(enc_outputs_class, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, enc_outputs_class.shape[-1]))
    target = torch.gather(object_query_embedding, 1,
                          topk_proposals.unsqueeze(-1).repeat(1, 1, object_query_embedding.shape[-1]))
    return topk_proposals, coords.sigmoid(), coords, cls, target


@torch.no_grad()
def forward_seg_heads(mask_embed_head, output, mask_features):
    """einsum('bqc,bchw->bqhw', mask_embed(output), mask_features) as one [Q,C] x [HW,C]^T GEMM per image."""
    e = mask_embed_head(output)
    if isinstance(mask_features, tuple):                                  # (rows [B, H*W, C], H, W): the neck's own layout
        f, H, W = mask_features
        B = f.shape[0]
    else:
        B, C, H, W = mask_features.shape
        f = mask_features.permute(0, 2, 3, 1).reshape(B, H * W, C)      # free for channels_last features
        if not f.is_contiguous():
            f = f.contiguous()
    out = torch.empty((B, e.shape[1], H * W), dtype=e.dtype, device=e.device)
    for b in range(B):
        ops.linear(e[b].contiguous(), f[b], out=out[b])
    return out.view(B, e.shape[1], H, W)


@torch.no_grad()
def patch2query_mean(patch2query_head, text_query):
    """text_query [bs, n_cls, num_embs, C_llm] -> [bs, n_cls, d_model]: MLP then mean over the [EMB] states."""
    return patch2query_head(text_query).mean(-2)


def box_cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


@torch.no_grad()
def post_process_det_gdino(logits, pred_boxes, target_sizes, num_classes, threshold=0.0, topk=100):
    """eval_det.py:18-56.  Returns per-image dicts (scores, labels, boxes) plus the raw index tensors."""
    logits = logits[:, :, :num_classes]
    B, Q, K = logits.shape
    prob = logits.sigmoid().view(B, -1)
    k = min(topk, prob.size(1))
    scores, idx = torch.topk(prob, k, dim=1)
    box_idx = torch.div(idx, K, rounding_mode="floor")
    labels = idx % K
    boxes = torch.gather(box_cxcywh_to_xyxy(pred_boxes), 1, box_idx.unsqueeze(-1).repeat(1, 1, 4))
    ts = torch.as_tensor(target_sizes, dtype=torch.float32, device=boxes.device)
    img_h, img_w = ts.unbind(1)
    boxes = boxes * torch.stack([img_w, img_h, img_w, img_h], dim=1)[:, None, :]
    res = []
    for s, l, bx in zip(scores, labels, boxes):
        keep = s > threshold
        res.append({"scores": s[keep], "labels": l[keep], "boxes": bx[keep]})
    return res, idx, box_idx


@torch.no_grad()
def post_process_instseg_gdino(logits, pred_boxes, pred_masks, target_sizes, image_sizes, num_classes=80, topk=100,
                               mask_stride=4):
    """eval_det.py:59-104: per image top-k, box scale, mask 4x bilinear -> crop -> resize -> sigmoid > 0.5."""
    logits = logits[:, :, :num_classes]
    res = []
    for i in range(logits.shape[0]):
        K = logits.shape[-1]
        prob = logits[i].sigmoid().view(-1)
        k = min(topk, prob.size(0))
        scores, idx = torch.topk(prob, k, dim=0)
        box_idx = torch.div(idx, K, rounding_mode="floor")
        labels = idx % K
        ori_h, ori_w = target_sizes[i][:2]
        boxes = box_cxcywh_to_xyxy(pred_boxes[i][box_idx]) * torch.as_tensor(
            [ori_w, ori_h, ori_w, ori_h], dtype=torch.float32, device=pred_boxes.device)[None, :]
        m = pred_masks[i][box_idx]
        H, W = m.shape[-2:]
        m = F.interpolate(m[:, None], size=(H * mask_stride, W * mask_stride), mode="bilinear", align_corners=False)
        m = m[:, :, :image_sizes[i][0], :image_sizes[i][1]]
        m = F.interpolate(m, size=(ori_h, ori_w), mode="bilinear", align_corners=False)[:, 0]
        res.append({"scores": scores, "labels": labels, "boxes": boxes, "masks": m.sigmoid() > 0.5,
                    "topk_indexes": idx, "topk_boxes": box_idx})
    return res

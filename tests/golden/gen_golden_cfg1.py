"""BASELINE cfg 1 end to end from the REFERENCE's own modules (build container only; needs /root/reference):
reference `InternVisionModel` (ViT-B size) -> `mlp2x_gelu` bridge -> HF `LlamaForCausalLM` (1 layer; the reference's
LLM is third-party transformers) -> [EMB] hidden states -> reference `OVGroundingDinoForObjectDetection.forward_test`
(Swin backbone, 6 + 6 layers, 100 queries, S = 1045, grid_sample MSDA = the pure-PyTorch fallback), composed in the
ORDER of `VisionLLMv2Model.forward` (modeling_visionllmv2.py:419-468 [EMB] injection, :559-605 ViT/bridge/scatter,
:724-738 LLM + fp32 logits, :769-791 text_query gather + gdino) -- the whole class cannot be imported here (peft /
diffusers / mmdet / detectron2), so its forward is restated with the reference's own sub-modules.

Stores fp32 outputs, the same pipeline run in bf16 (the reference's deployed precision; two-stage top-k pinned to the
fp32 leg's indices) and the CPU wall time of the fp32 leg (the cfg-1 "CPU reference fwd" number).

    python tests/golden/gen_golden_cfg1.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import cfg1_common as C  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402


MASK_SUB = 4        # keep every 4th mask logit (fixture size)


def build():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfgv, modv = ref_shim.load_internvit()
    vit = modv.InternVisionModel(cfgv.InternVisionConfig(use_flash_attn=False, drop_path_rate=0.0, **C.VIT)).eval()
    vit.load_state_dict(seeded_state_dict(vit, C.SEEDS["vit"]))
    bridge = C.bridge_module().eval()
    bridge.load_state_dict(seeded_state_dict(bridge, C.SEEDS["bridge"]))
    llm = LlamaForCausalLM(LlamaConfig(attn_implementation="eager", **C.LLM)).eval()
    llm.load_state_dict(seeded_state_dict(llm, C.SEEDS["llm"]))
    emb_det = torch.nn.Embedding(C.NUM_EMBS, C.L_HIDDEN)
    emb_det.load_state_dict(seeded_state_dict(emb_det, C.SEEDS["emb"]))
    cfgm, gd = ref_shim.load_gdino()
    gcfg = cfgm.GroundingDinoConfig(backbone_config=C.swin_config(), fusion_dropout=0., fusion_droppath=0.,
                                    text_enhancer_dropout=0., disable_custom_kernels=True, **C.GDINO)
    gdino = gd.OVGroundingDinoForObjectDetection(gcfg).eval()
    sd = seeded_state_dict(gdino, C.SEEDS["gdino"])
    for k in sd:
        if k.endswith("vision_param") or k.endswith("text_param"):
            sd[k] = sd[k] * 0 + 0.5
    gdino.load_state_dict(sd)
    return vit, bridge, llm, emb_det, gdino, gd


def forward(mods, ids, image, aug, dtype, force_topk=None):
    """VisionLLMv2Model.forward restated (line numbers: modeling_visionllmv2.py)."""
    vit, bridge, llm, emb_det, gdino, gd = mods
    with torch.no_grad():
        input_ids = ids.clone()
        inputs_embeds = llm.get_input_embeddings()(input_ids)                                    # :420
        emb_ids = torch.arange(C.EMB, C.EMB + C.NUM_EMBS)                                        # :432
        gap_len = C.NUM_EMBS                                                                     # :430 ([EMB] present)
        new_ids, new_emb = [], []
        for cur_ids, cur_emb in zip(input_ids, inputs_embeds):                                   # :441-468
            for start in torch.where(cur_ids == C.DET)[0]:
                cur_ids = torch.cat([cur_ids[: start + 1], emb_ids, cur_ids[start + gap_len + 1:]], 0)
                cur_emb = torch.cat([cur_emb[: start + 1], emb_det.weight.to(cur_emb.dtype),
                                     cur_emb[start + gap_len + 1:]], 0).contiguous()
            new_ids.append(cur_ids); new_emb.append(cur_emb)
        input_ids, inputs_embeds = torch.stack(new_ids), torch.stack(new_emb)
        outs = vit(image.to(dtype), output_hidden_states=True)                                   # :571
        feats = outs.hidden_states[-1][:, 1:].to(dtype)                                          # :572-574 (vis_output_layer -1)
        image_features = bridge(feats).to(inputs_embeds.dtype)                                   # :579
        B, L, Cc = inputs_embeds.shape
        flat = inputs_embeds.reshape(B * L, Cc)
        selected = (input_ids == C.IMP).reshape(-1)                                              # :584-590
        flat[selected] = flat[selected] * 0.0 + image_features.reshape(-1, Cc)                   # :594
        inputs_embeds = flat.reshape(B, L, Cc)
        out = llm(attention_mask=torch.ones_like(input_ids), inputs_embeds=inputs_embeds, output_hidden_states=True,
                  return_dict=True)                                                              # :724-732
        hidden = out.hidden_states[-1]
        logits = llm.lm_head(hidden).float()                                                     # :737-738
        pixel_values = aug.to(dtype)                                                             # :771 (224 % 32 == 0)
        pixel_mask = pixel_values[:, 0, :, :] != 0                                               # :773
        emb_select = (input_ids >= C.EMB) & (input_ids <= C.EMB + C.NUM_EMBS - 1)                # :776
        num_patches = emb_select.sum(-1) // C.NUM_EMBS
        text_query = torch.zeros((B, int(num_patches.max()), C.NUM_EMBS, Cc), dtype=hidden.dtype)
        text_query_masks = torch.zeros(B, int(num_patches.max()), dtype=torch.bool)
        for b in range(B):                                                                       # :783-787
            text_query[b, :num_patches[b]] = hidden[b, emb_select[b], :].reshape(-1, C.NUM_EMBS, Cc)
            text_query_masks[b, :num_patches[b]] = 1
        real_topk = torch.topk
        if force_topk is not None:
            def topk(inp, k, dim=-1, **kw):
                if inp.dim() == 2 and inp.shape[0] == force_topk.shape[0] and k == force_topk.shape[1]:
                    return torch.gather(inp, 1, force_topk), force_topk
                return real_topk(inp, k, dim=dim, **kw)
            torch.topk = topk
        try:
            text = gdino.patch2query(text_query).mean(-2)
            mo = gdino.model(pixel_values=pixel_values, pixel_mask=pixel_mask, text_query=text,
                             text_query_masks=text_query_masks, return_dict=True)
            g = gdino(pixel_values, pixel_mask=pixel_mask, text_query=text_query, text_query_masks=text_query_masks,
                      img_metas=[{"task": "det"}], labels=None)                                  # :788
        finally:
            torch.topk = real_topk
        topk_idx = real_topk(mo.enc_outputs_class.max(-1)[0].float(), gdino.config.num_queries, dim=1)[1]
    return dict(llm_logits=logits, llm_hidden=hidden.float(), text_query=text_query.float(),
                enc_class_max=mo.enc_outputs_class.float().max(-1)[0],
                gd_logits=g.logits.float(), gd_boxes=g.pred_boxes.float(),
                gd_masks=g.pred_masks.float().reshape(B, -1)[:, ::MASK_SUB]), input_ids, topk_idx


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mods = build()
    vit, bridge, llm, emb_det, gdino, gd = mods
    ids, image, aug = C.inputs()
    forward(mods, ids, image, aug, torch.float32)                     # warm-up
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        o32, new_ids, idx = forward(mods, ids, image, aug, torch.float32)
        times.append(time.perf_counter() - t0)
    cpu_ms = sorted(times)[1] * 1e3
    print(f"reference CPU fp32 forward (8 threads): {cpu_ms:.0f} ms")
    # bf16 leg: the deployed dataflow (custom-kernel MSDA branch emulated by the reference's own fp32 function)
    class _Ext:
        @staticmethod
        def ms_deform_attn_forward(value, shapes, lsi, loc, w, step):
            return gd.multi_scale_deformable_attention(value, shapes, loc, w)
    gd.MultiScaleDeformableAttention = _Ext
    for m in (vit, bridge, llm, emb_det, gdino):
        m.bfloat16()
    for mm in gdino.modules():
        if hasattr(mm, "disable_custom_kernels"):
            mm.disable_custom_kernels = False
    o16, _, _ = forward(mods, ids, image, aug, torch.bfloat16, force_topk=idx)
    _, _, idx16 = forward(mods, ids, image, aug, torch.bfloat16)
    print("bf16 reference, unforced top-k overlap:", len(set(idx[0].tolist()) & set(idx16[0].tolist())) / idx.shape[1])
    for m in (vit, bridge, llm, emb_det, gdino):
        m.float()
    # detection post-processing of the fp32 leg (eval_det.py:18-56 restated by the reference's own primitives)
    K = C.N_CLS
    prob = o32["gd_logits"][:, :, :K].sigmoid().view(1, -1)
    tv, ti = torch.topk(prob, min(100, prob.shape[1]), dim=1)
    arrs = dict(mask_sub=np.array(MASK_SUB), new_input_ids=new_ids.numpy(), topk=idx.numpy(), det_topk_indexes=ti.numpy(),
                det_box_idx=torch.div(ti, K, rounding_mode="floor").numpy(), det_labels=(ti % K).numpy(),
                det_scores=tv.numpy(), cpu_ms_fp32_8threads=np.array(cpu_ms),
                keys_vit=np.array(json.dumps(key_shapes(vit))), keys_llm=np.array(json.dumps(key_shapes(llm))),
                keys_gdino=np.array(json.dumps(key_shapes(gdino))))
    for k, v in o32.items():
        arrs[k + "_f32"] = v.numpy()
        arrs[k + "_refbf16"] = o16[k].float().numpy()
        fin = torch.isfinite(v)
        print(f"  {k:14s} {tuple(v.shape)}  rel_l2(ref bf16 vs fp32) = {float((o16[k].float() - v)[fin].norm() / v[fin].norm()):.4f}")
    path = os.path.join(HERE, "cfg1_e2e.npz")
    np.savez_compressed(path, **arrs)
    print("wrote cfg1_e2e.npz", os.path.getsize(path))


if __name__ == "__main__":
    main()

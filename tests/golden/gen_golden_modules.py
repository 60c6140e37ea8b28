"""Golden vectors for module-level parity, produced by RUNNING THE REFERENCE'S OWN CLASSES on CPU
(build container only; needs /root/reference).   python tests/golden/gen_golden_modules.py

For every case we store: the (bf16-representable) inputs, the reference output in fp32, the reference's own
output when the module runs in bf16 (the precision the reference deploys in), and the state-dict key/shape list.
Weights are regenerated from a seed by tests/golden/weights_util.py on both sides.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **{k: (v.detach().float().numpy() if torch.is_tensor(v) else v)
                                                     for k, v in arrs.items()})
    print("wrote", name, {k: tuple(v.shape) for k, v in arrs.items() if hasattr(v, "shape")})


def run_both(module, fn, refmod=None):
    """fn(module, dtype) -> tensor; returns (fp32 output, output of the reference run in bf16).

    bf16 leg of MSDA modules: the reference's grid_sample fallback cannot run in bf16 (grid is fp32); its
    deployed bf16 dataflow is the custom-kernel branch (upcast value/weights to fp32, call the op,
    gd.py:763-776).  We take that branch with the op emulated by the reference's OWN fp32 pytorch function."""
    module.float().eval()
    with torch.no_grad():
        o32 = fn(module, torch.float32)
        module.bfloat16()
        if refmod is not None:
            class _Ext:
                @staticmethod
                def ms_deform_attn_forward(value, shapes, lsi, loc, w, step):
                    return refmod.multi_scale_deformable_attention(value, shapes, loc, w)
            refmod.MultiScaleDeformableAttention = _Ext
            for mm in module.modules():
                if hasattr(mm, "disable_custom_kernels"):
                    mm.disable_custom_kernels = False
        o16 = fn(module, torch.bfloat16).float()
        module.float()
    return o32, o16


def internvit():
    cfgm, mod = ref_shim.load_internvit()
    cfg = cfgm.InternVisionConfig(hidden_size=256, num_attention_heads=2, num_hidden_layers=2, intermediate_size=512,
                                  image_size=56, patch_size=14, qk_normalization=True, use_flash_attn=False,
                                  qkv_bias=False, drop_path_rate=0.0)
    m = mod.InternVisionModel(cfg)
    m.load_state_dict(seeded_state_dict(m, 101))
    x = bf16r(torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(5)))

    def fn(mm, dt):
        o = mm(pixel_values=x.to(dt), output_hidden_states=True, return_dict=True)
        return torch.stack([o.hidden_states[-1], o.hidden_states[-2], o.hidden_states[0]]).float()

    o32, o16 = run_both(m, fn)
    save("mod_internvit_small.npz", pixel_values=x, out_f32=o32, out_refbf16=o16,
         keys=np.array(json.dumps(key_shapes(m))))


def gdino():
    cfgm, mod = ref_shim.load_gdino()
    cfg = cfgm.GroundingDinoConfig(d_model=256, encoder_attention_heads=8, decoder_attention_heads=8,
                                   encoder_ffn_dim=512, decoder_ffn_dim=512, num_feature_levels=4, encoder_n_points=4,
                                   decoder_n_points=4, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                                   disable_custom_kernels=True)
    shapes_l = [(12, 16), (6, 8), (3, 4), (2, 2)]
    shapes = torch.tensor(shapes_l, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    B = 2
    g = torch.Generator().manual_seed(7)
    src = bf16r(torch.randn(B, S, 256, generator=g))
    pos = bf16r(torch.randn(B, S, 256, generator=g) * 0.5)
    refs = []
    for (H, W) in shapes_l:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        refs.append(torch.stack(((xs + 0.5) / W, (ys + 0.5) / H), -1).reshape(-1, 2))
    ref2 = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1)          # [B, S, L, 2]  (fp32, like valid_ratios)
    mask = torch.ones(B, S, dtype=torch.bool)
    mask[1, -5:] = False                                                     # a few padded pixels

    lay = mod.GroundingDinoDeformableLayer(cfg)
    lay.load_state_dict(seeded_state_dict(lay, 202))

    def fn_enc(mm, dt):
        return mm(src.to(dt), mask, position_embeddings=pos.to(dt), reference_points=ref2, spatial_shapes=shapes,
                  level_start_index=lsi)[0].float()

    o32, o16 = run_both(lay, fn_enc, mod)
    save("mod_gdino_deformable_layer.npz", src=src, pos=pos, ref=ref2, mask=mask.numpy(), shapes=shapes.numpy(),
         lsi=lsi.numpy(), out_f32=o32, out_refbf16=o16, keys=np.array(json.dumps(key_shapes(lay))))

    Q, T = 37, 11
    dec = mod.GroundingDinoDecoderLayer(cfg)
    dec.load_state_dict(seeded_state_dict(dec, 303))
    hs = bf16r(torch.randn(B, Q, 256, generator=g))
    qpos = bf16r(torch.randn(B, Q, 256, generator=g) * 0.5)
    boxes = torch.rand(B, Q, 4, generator=g) * 0.5 + 0.2
    ref4 = boxes[:, :, None, :].repeat(1, 1, 4, 1)                          # [B, Q, L, 4]
    text = bf16r(torch.randn(B, T, 256, generator=g))
    tpad = torch.zeros(B, T, dtype=torch.bool)
    tpad[1, 7:] = True                                                       # key_padding_mask: True = ignore

    def fn_dec(mm, dt):
        return mm(hs.to(dt), position_embeddings=qpos.to(dt), reference_points=ref4, spatial_shapes=shapes,
                  level_start_index=lsi, vision_encoder_hidden_states=src.to(dt), vision_encoder_attention_mask=mask,
                  text_encoder_hidden_states=text.to(dt), text_encoder_attention_mask=tpad)[0].float()

    o32, o16 = run_both(dec, fn_dec, mod)
    save("mod_gdino_decoder_layer.npz", hs=hs, qpos=qpos, ref=ref4, text=text, tpad=tpad.numpy(), src=src,
         mask=mask.numpy(), shapes=shapes.numpy(), lsi=lsi.numpy(), out_f32=o32, out_refbf16=o16,
         keys=np.array(json.dumps(key_shapes(dec))))


def gdino_encoder_layer():
    """Full GroundingDinoEncoderLayer (fusion bi-attention + text enhancer + deformable layer), gd.py:1216-1289."""
    cfgm, mod = ref_shim.load_gdino()
    cfg = cfgm.GroundingDinoConfig(d_model=256, encoder_attention_heads=8, decoder_attention_heads=8,
                                   encoder_ffn_dim=512, decoder_ffn_dim=512, num_feature_levels=4, encoder_n_points=4,
                                   decoder_n_points=4, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                                   fusion_dropout=0.0, fusion_droppath=0.0, text_enhancer_dropout=0.0,
                                   disable_custom_kernels=True)
    shapes_l = [(12, 16), (6, 8), (3, 4), (2, 2)]
    shapes = torch.tensor(shapes_l, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    B, T = 2, 9
    g = torch.Generator().manual_seed(17)
    src = bf16r(torch.randn(B, S, 256, generator=g))
    pos = bf16r(torch.randn(B, S, 256, generator=g) * 0.5)
    text = bf16r(torch.randn(B, T, 256, generator=g))
    refs = []
    for (H, W) in shapes_l:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        refs.append(torch.stack(((xs + 0.5) / W, (ys + 0.5) / H), -1).reshape(-1, 2))
    ref2 = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1)
    kpm = torch.zeros(B, S, dtype=torch.bool)                    # key_padding_mask: True = padded pixel
    kpm[1, 150:192] = True; kpm[1, -3:] = True
    tq_mask = torch.ones(B, T, dtype=torch.bool); tq_mask[1, 5:] = False     # text_query_masks: valid = 1
    tsa, pids = mod.generate_masks_with_text_query_masks(tq_mask)
    lay = mod.GroundingDinoEncoderLayer(cfg)
    sd = seeded_state_dict(lay, 505)
    for k in sd:                                                  # LayerScale params big enough to matter
        if k.endswith("vision_param") or k.endswith("text_param"):
            sd[k] = sd[k] * 0 + 0.5
    lay.load_state_dict(sd)

    def fn(mm, dt):
        (v, t), _ = mm(vision_features=src.to(dt), vision_position_embedding=pos.to(dt), spatial_shapes=shapes,
                       level_start_index=lsi, key_padding_mask=kpm, reference_points=ref2, text_features=text.to(dt),
                       text_attention_mask=~tq_mask, text_position_embedding=None, text_self_attention_masks=tsa,
                       text_position_ids=pids)
        return torch.cat([v.float().flatten(1), t.float().flatten(1)], 1)

    o32, o16 = run_both(lay, fn, mod)
    save("mod_gdino_encoder_layer.npz", src=src, pos=pos, text=text, ref=ref2, kpm=kpm.numpy(), tq_mask=tq_mask.numpy(),
         tsa=tsa.numpy(), pids=pids.numpy(), shapes=shapes.numpy(), lsi=lsi.numpy(), out_f32=o32, out_refbf16=o16,
         keys=np.array(json.dumps(key_shapes(lay))))


def internlm2():
    """Vendored InternLM2ForCausalLM (GQA, fused wqkv), eager attention, fp32 + bf16 legs."""
    cfgm, mod = ref_shim.load_internlm2()
    cfg = cfgm.InternLM2Config(vocab_size=512, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                               num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-5,
                               max_position_embeddings=256, attn_implementation="eager", bias=False)
    cfg.rope_scaling = None          # transformers 5.x rewrites the field into a dict the 4.34-era code cannot read
    m = mod.InternLM2ForCausalLM(cfg)
    m.load_state_dict(seeded_state_dict(m, 606))
    B, T = 2, 33
    g = torch.Generator().manual_seed(9)
    emb = bf16r(torch.randn(B, T, 512, generator=g) * 0.5)
    am = torch.ones(B, T, dtype=torch.long); am[1, 20:] = 0

    def fn(mm, dt):
        o = mm(inputs_embeds=emb.to(dt), attention_mask=am, output_hidden_states=True, return_dict=True, use_cache=False)
        return torch.cat([o.hidden_states[-1].float(), o.hidden_states[1].float(), o.logits.float()], -1)

    o32, o16 = run_both(m, fn)
    save("mod_internlm2_small.npz", emb=emb, am=am.numpy(), out_f32=o32, out_refbf16=o16,
         keys=np.array(json.dumps(key_shapes(m))))


if __name__ == "__main__":
    which = sys.argv[1:] or ["internvit", "gdino", "gdino_encoder_layer", "internlm2"]
    for w in which:
        globals()[w]()

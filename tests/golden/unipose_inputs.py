"""Deterministic inputs of the UniPose layer parity tests (CPU generator, bf16-representable), shared by the golden
generator (build container) and the tests (GPU box) so the golden file only carries outputs."""
import torch

SHAPES = [(12, 16), (6, 8), (3, 4), (2, 2)]


def inputs():
    g = torch.Generator().manual_seed(9)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).float()  # noqa: E731
    S = sum(h * w for h, w in SHAPES)
    bs, nq, ntok, d = 2, 19, 7, 256
    shapes = torch.tensor(SHAPES, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    pad = torch.zeros(bs, S, dtype=torch.bool)
    pad[1, 150:192] = True
    ref2 = torch.rand(bs, S, 4, 2, generator=g)
    ref4 = torch.cat((torch.rand(nq, bs, 4, 2, generator=g), torch.rand(nq, bs, 4, 2, generator=g) * 0.4 + 0.05), -1)
    attn_mask = torch.zeros(nq, nq, dtype=torch.bool)
    attn_mask[:9, 9:] = True
    attn_mask[9:, :9] = True
    text_mask = torch.zeros(bs, ntok, dtype=torch.bool)
    text_mask[1, 5:] = True
    return dict(src=r(bs, S, d), pos=r(bs, S, d), ref2=ref2, shapes=shapes, lsi=lsi, pad=pad, tgt=r(nq, bs, d),
                qpos=r(nq, bs, d), ref4=ref4, memory=r(S, bs, d), memory_text=r(bs, ntok, d), text_mask=text_mask,
                attn_mask=attn_mask)


# ---- two-stage keypoint decoder (modeling_unipose.py:2869-3130) -------------------------------------------------------
DEC = dict(d_model=256, d_ffn=512, n_heads=8, num_layers=4, num_box_decoder_layers=2, num_body_points=19, nq=60, bs=2,
           ntok=6)


def decoder_inputs():
    g = torch.Generator().manual_seed(23)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).float()  # noqa: E731
    S = sum(h * w for h, w in SHAPES)
    c = DEC
    bs, nq, d, nbp, ntok = c["bs"], c["nq"], c["d_model"], c["num_body_points"], c["ntok"]
    shapes = torch.tensor(SHAPES, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    pad = torch.zeros(bs, S, dtype=torch.bool)
    pad[1, 160:192] = True
    ref_unsig = torch.cat((torch.randn(nq, bs, 2, generator=g), torch.randn(nq, bs, 2, generator=g) * 0.5 - 1.5), -1)
    valid_ratios = (0.75 + 0.25 * torch.rand(bs, len(SHAPES), 2, generator=g))
    text_mask = torch.zeros(bs, ntok, dtype=torch.bool)          # True = padding (layer cross-attention)
    text_mask[1, 4:] = True
    kpt_vis = torch.zeros(bs, nbp, dtype=torch.long)             # kpt_query_masks: 1 = a real keypoint class
    kpt_vis[0, :17] = 1
    kpt_vis[1, :12] = 1
    return dict(tgt=r(nq, bs, d), memory=r(S, bs, d), pad=pad, shapes=shapes, lsi=lsi,
                ref_unsig=ref_unsig.to(torch.bfloat16).float(), valid_ratios=valid_ratios.to(torch.bfloat16).float(),
                memory_text=r(bs, ntok, d), text_mask=text_mask, encoded_text=r(bs, ntok, d) * 2.0,
                kpt_embed=r(bs, nbp, d), kpt_vis=kpt_vis)


# ---- the whole transformer: fused encoder + two-stage selection + keypoint decoder (modeling_unipose.py:2206-2700) ----------
TR = dict(d_model=256, nhead=8, num_queries=60, num_encoder_layers=2, num_decoder_layers=4, dim_feedforward=512,
          num_feature_levels=4, num_box_decoder_layers=2, num_body_points=19, bs=2, ntok=6)


def transformer_kwargs():
    c = TR
    return dict(d_model=c["d_model"], nhead=c["nhead"], num_queries=c["num_queries"], num_encoder_layers=c["num_encoder_layers"],
                num_decoder_layers=c["num_decoder_layers"], dim_feedforward=c["dim_feedforward"], dropout=0.0,
                return_intermediate_dec=True, query_dim=4, deformable_encoder=True, deformable_decoder=True,
                num_feature_levels=c["num_feature_levels"], enc_n_points=4, dec_n_points=4, learnable_tgt_init=True,
                two_stage_type="standard", embed_init_tgt=True, use_text_enhancer=True, use_fusion_layer=True,
                use_text_cross_attention=True, text_dropout=0.0, fusion_dropout=0.0, fusion_droppath=0.0, decoder_sa_type="sa")


def transformer_inputs():
    g = torch.Generator().manual_seed(31)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).float()  # noqa: E731
    c = TR
    bs, d, nbp, ntok = c["bs"], c["d_model"], c["num_body_points"], c["ntok"]
    srcs = [r(bs, d, h, w) for h, w in SHAPES]
    poss = [r(bs, d, h, w) for h, w in SHAPES]
    masks = []
    for h, w in SHAPES:                                          # batch entry 1 is padded on the right / bottom
        m = torch.zeros(bs, h, w, dtype=torch.bool)
        m[1, :, int(w * 0.75):] = True
        m[1, int(h * 0.8):, :] = True
        masks.append(m)
    obj_mask = torch.zeros(bs, ntok, dtype=torch.long)           # obj_query_masks: 1 = a real class token
    obj_mask[0, :5] = 1
    obj_mask[1, :3] = 1
    kpt_vis = torch.zeros(bs, nbp, dtype=torch.long)
    kpt_vis[0, :17] = 1
    kpt_vis[1, :12] = 1
    return dict(srcs=srcs, poss=poss, masks=masks, obj_mask=obj_mask, encoded_text=r(bs, ntok, d) * 2.0, kpt_embed=r(bs, nbp, d),
                kpt_vis=kpt_vis)


# ---- the model behind its backbone (modeling_unipose.py:69-655) -------------------------------------------------------------
MODEL = dict(l_hidden=64, backbone_channels=(32, 48, 64), n_obj=6, n_kpt=19, n_emb=4)


def model_inputs():
    """Backbone outputs of a padded 2-image batch (3 levels; the 4th is derived by input_proj[3]) + LLM [EMB] states."""
    g = torch.Generator().manual_seed(37)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).float()  # noqa: E731
    bs = TR["bs"]
    Hs, Ws = 96, 128                                              # padded batch size in pixels (stride 8 -> 12 x 16)
    sample_mask = torch.zeros(bs, Hs, Ws, dtype=torch.bool)
    sample_mask[1, :, 96:] = True
    sample_mask[1, 80:, :] = True
    feats, poss = [], []
    for (h, w), c in zip(SHAPES[:3], MODEL["backbone_channels"]):
        m = torch.nn.functional.interpolate(sample_mask[None].float(), size=(h, w)).to(torch.bool)[0]
        feats.append((r(bs, c, h, w), m))
        poss.append(r(bs, TR["d_model"], h, w))
    obj_mask = torch.zeros(bs, MODEL["n_obj"], dtype=torch.long)
    obj_mask[0, :5] = 1
    obj_mask[1, :3] = 1
    kpt_mask = torch.zeros(bs, MODEL["n_kpt"], dtype=torch.long)
    kpt_mask[0, :17] = 1
    kpt_mask[1, :12] = 1
    tq = dict(obj_querys=r(bs, MODEL["n_obj"], MODEL["n_emb"], MODEL["l_hidden"]) * 2, obj_query_masks=obj_mask,
              kpt_querys=r(bs, MODEL["n_kpt"], MODEL["n_emb"], MODEL["l_hidden"]) * 2, kpt_query_masks=kpt_mask)
    return dict(feats=feats, poss=poss, sample_mask=sample_mask, text_query=tq)


# ---- the image backbone: Joiner(SwinTransformer, PositionEmbeddingSineHW) (modeling_unipose.py:1212-1226, 1638-1858) --------
BACKBONE = dict(embed_dim=64, depths=[2, 2, 2, 2], num_heads=[2, 4, 8, 16], window_size=7, out_indices=[1, 2, 3],
                hidden_dim=256)


def backbone_inputs():
    """A padded 2-image batch as `nested_tensor_from_tensor_list` builds it: image 1 is 80 x 110 inside 100 x 138 (zeros and
    mask = True outside).  138 is not a multiple of the patch size (pad in PatchEmbed); 25 x 35 tokens pad to 28 x 35 windows."""
    g = torch.Generator().manual_seed(41)
    bs, H, W = 2, 100, 138
    x = torch.randn(bs, 3, H, W, generator=g).to(torch.bfloat16).float()
    mask = torch.zeros(bs, H, W, dtype=torch.bool)
    mask[1, 80:, :] = True
    mask[1, :, 110:] = True
    x[1] = x[1] * (~mask[1]).float()
    return x, mask

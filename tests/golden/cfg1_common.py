"""BASELINE cfg 1: "single 224x224 image + 16-token prompt, ViT-B + 1-layer LLM stub, CPU reference fwd,
world_size=1" -- the shapes, ids and seeds shared by the golden generator (reference modules, build container) and the
GPU parity test / CPU timing (our modules, GPU box).  No reference code here."""
import torch

IMP, DET, EMB = 990, 989, 991           # <im_patch>, [DET], [EMB]..[EMB4] inside the 1000-word stub vocabulary
NUM_EMBS, N_CLS, N_TEXT, VOCAB = 4, 5, 16, 1000
L_HIDDEN = 512
VIT = dict(hidden_size=768, num_attention_heads=12, num_hidden_layers=12, intermediate_size=3072, image_size=224,
           patch_size=14, qk_normalization=True, qkv_bias=False)          # ViT-B-sized InternViT: 16 x 16 + 1 tokens
LLM = dict(hidden_size=L_HIDDEN, intermediate_size=1376, num_hidden_layers=1, num_attention_heads=8,
           num_key_value_heads=8, vocab_size=VOCAB, rms_norm_eps=1e-5, max_position_embeddings=1024)
SEEDS = dict(vit=7101, bridge=7102, llm=7103, emb=7104, gdino=7105, data=7106)


def swin_config():
    from transformers import SwinConfig
    return SwinConfig(image_size=224, embed_dim=96, depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=7,
                      out_features=["stage1", "stage2", "stage3", "stage4"])


GDINO = dict(d_model=256, encoder_layers=6, decoder_layers=6, encoder_ffn_dim=2048, decoder_ffn_dim=2048,
             encoder_attention_heads=8, decoder_attention_heads=8, num_queries=100, num_feature_levels=4,
             encoder_n_points=4, decoder_n_points=4, dropout=0., attention_dropout=0., activation_dropout=0.,
             mask_dim=256, norm="GN", l_hidden_size=L_HIDDEN)           # levels 28^2, 14^2, 7^2, 4^2: S = 1045


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def inputs():
    """(input_ids [1, 297], image [1, 3, 224, 224], images_aug [1, 3, 224, 224]) -- bf16-representable floats."""
    g = torch.Generator().manual_seed(SEEDS["data"])
    ids = torch.randint(0, 900, (1, 256 + N_TEXT + N_CLS * (1 + NUM_EMBS)), generator=g)
    ids[0, :256] = IMP
    for c in range(N_CLS):                  # "name[DET][EMB][EMB2][EMB3][EMB4]" (coco_llava.py:230-238), placeholders = [EMB]
        p = 256 + N_TEXT + c * (1 + NUM_EMBS)
        ids[0, p] = DET
        ids[0, p + 1:p + 1 + NUM_EMBS] = EMB
    image = bf16r(torch.randn(1, 3, 224, 224, generator=g))
    aug = bf16r(torch.randn(1, 3, 224, 224, generator=g))
    return ids, image, aug


def bridge_module():
    """mv2.py:174-182 `mlp2x_gelu`: Linear(768 -> 512) + GELU + Linear(512 -> 512) (a plain nn.Sequential)."""
    import torch.nn as nn
    return nn.Sequential(nn.Linear(VIT["hidden_size"], L_HIDDEN), nn.GELU(), nn.Linear(L_HIDDEN, L_HIDDEN))


def build_b200_model(g=None, device="cuda", dtype=torch.bfloat16):
    """The cfg-1 composite on our modules, weights regenerated from the generator's seeds (key lists checked)."""
    import json
    from types import SimpleNamespace
    from transformers import LlamaConfig
    from weights_util import key_shapes, seeded_state_dict
    from visionllm_b200.gdino_model import B200GroundingDinoForObjectDetection
    from visionllm_b200.internvit import B200InternVisionModel, InternVisionConfig
    from visionllm_b200.llama import B200LlamaForCausalLM
    from visionllm_b200.modeling import B200VisionLLMv2Model
    from visionllm_b200.swin import B200SwinBackbone
    vit = B200InternVisionModel(InternVisionConfig(**VIT))
    llm = B200LlamaForCausalLM(LlamaConfig(**LLM))
    gcfg = SimpleNamespace(backbone_config=swin_config(), activation_function="relu", max_text_len=256, query_dim=4,
                           two_stage=True, embedding_init_target=True, two_stage_bbox_embed_share=False,
                           decoder_bbox_embed_share=True, position_embedding_type="sine",
                           positional_embedding_temperature=20, **GDINO)
    gdino = B200GroundingDinoForObjectDetection(gcfg, backbone_model=B200SwinBackbone(gcfg.backbone_config))
    if g is not None:
        for mod, key in ((vit, "keys_vit"), (llm, "keys_llm"), (gdino, "keys_gdino")):
            assert json.loads(str(g[key])) == [list(k) for k in key_shapes(mod)], f"{key}: state-dict keys differ"
    vit.load_state_dict(seeded_state_dict(vit, SEEDS["vit"]))
    llm.load_state_dict(seeded_state_dict(llm, SEEDS["llm"]))
    sd = seeded_state_dict(gdino, SEEDS["gdino"])
    for k in sd:
        if k.endswith("vision_param") or k.endswith("text_param"):
            sd[k] = sd[k] * 0 + 0.5
    gdino.load_state_dict(sd)
    cfg = SimpleNamespace(use_pixelshuffle=False, vl_bridge_type="mlp2x_gelu", vis_output_layer=-1, num_embs=NUM_EMBS,
                          imp_token_id=IMP, emb_token_id=EMB, det_tool_id=DET, seg_tool_id=-1, grd_tool_id=-1,
                          pose_tool_id=-1)
    model = B200VisionLLMv2Model(cfg, vit, llm, gdino=gdino)
    ref_bridge = bridge_module()
    model.vl_bridge.load_state_dict(seeded_state_dict(ref_bridge, SEEDS["bridge"]))      # same keys as nn.Sequential
    model.emb_embeddings_det.load_state_dict(seeded_state_dict(torch.nn.Embedding(NUM_EMBS, L_HIDDEN), SEEDS["emb"]))
    return model.to(device, dtype).eval()

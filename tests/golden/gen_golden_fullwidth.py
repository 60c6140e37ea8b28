"""Full-WIDTH golden for one InternViT-6B encoder layer (hidden 3200, 25 heads x 128, MLP 12800, 1025 tokens,
qk-norm over the flattened 3200-d, LayerScale) from the REFERENCE's own `InternVisionEncoderLayer`
(visionllmv2/model/internvit/modeling_intern_vit.py:182-210) run on CPU in fp32 and in bf16 (build container only;
needs /root/reference).  Weights and input are regenerated from seeds on both sides (weights_util.py); only a subset
of token rows of the two outputs is stored to keep the fixture small.

    python tests/golden/gen_golden_fullwidth.py
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

SEED_W, SEED_X, TOKENS = 6001, 6002, 1025


def layer_input(tokens=TOKENS, hidden=3200):
    g = torch.Generator().manual_seed(SEED_X)
    return (torch.randn(1, tokens, hidden, generator=g)).to(torch.bfloat16).float()


def kept_rows(tokens=TOKENS):
    return np.unique(np.r_[0:8, np.arange(8, tokens, 29), tokens - 8:tokens])


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    cfgm, mod = ref_shim.load_internvit()
    cfg = cfgm.InternVisionConfig(hidden_size=3200, num_attention_heads=25, num_hidden_layers=1, intermediate_size=12800,
                                  image_size=448, patch_size=14, qk_normalization=True, use_flash_attn=False,
                                  qkv_bias=False, drop_path_rate=0.0, layer_norm_eps=1e-6, initializer_factor=0.1)
    layer = mod.InternVisionEncoderLayer(cfg, 0.0).eval()
    layer.load_state_dict(seeded_state_dict(layer, SEED_W))
    x = layer_input()
    rows = kept_rows()
    with torch.no_grad():
        o32 = layer.float()(x)
        o16 = layer.bfloat16()(x.bfloat16()).float()
    rel = float((o16 - o32).norm() / o32.norm())
    print("rel_l2(reference bf16 run vs fp32) =", rel, " |out| max", float(o32.abs().max()))
    np.savez_compressed(os.path.join(HERE, "mod_internvit6b_layer.npz"),
                        rows=rows, out_f32=o32[0, rows].numpy(), out_refbf16=o16[0, rows].numpy(),
                        ref_bf16_rel_l2_all_rows=np.array(rel), seeds=np.array([SEED_W, SEED_X]),
                        keys=np.array(json.dumps(key_shapes(layer))))
    print("wrote mod_internvit6b_layer.npz", os.path.getsize(os.path.join(HERE, "mod_internvit6b_layer.npz")))


if __name__ == "__main__":
    main()

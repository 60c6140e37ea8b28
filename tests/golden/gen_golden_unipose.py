"""Golden vectors for the UniPose deformable encoder / decoder layers from the REFERENCE's own
classes (visionllmv2/model/unipose/modeling_unipose.py:3132-3323) run on CPU in this build container; their MSDA
extension call is served by the reference's own pure-PyTorch core (ref_shim.load_unipose).  fp32 outputs + the
reference's bf16 run (MSDA in fp32 like its `value.dtype != float32` branch, ms_deform_attn.py:142-150)."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from weights_util import key_shapes, seeded_state_dict  # noqa: E402

from unipose_inputs import inputs  # noqa: E402

SUB = 4                      # every 4th encoder token is stored


def run(mu, dtype, x):
    c = lambda t: t.to(dtype) if t.is_floating_point() else t  # noqa: E731
    torch.manual_seed(0)
    enc = mu.DeformableTransformerEncoderLayer(256, 512, 0.0, "relu", 4, 8, 4).eval()
    enc.load_state_dict(seeded_state_dict(enc, 41))
    dec = mu.DeformableTransformerDecoderLayer(256, 512, 0.0, "relu", 4, 8, 4, use_text_cross_attention=True).eval()
    dec.load_state_dict(seeded_state_dict(dec, 42))
    enc, dec = enc.to(dtype), dec.to(dtype)
    with torch.no_grad():
        e = enc(c(x["src"]), c(x["pos"]), c(x["ref2"]), x["shapes"], x["lsi"], x["pad"])
        d = dec(tgt=c(x["tgt"]), tgt_query_pos=c(x["qpos"]), tgt_reference_points=c(x["ref4"]),
                memory_text=c(x["memory_text"]), text_attention_mask=x["text_mask"], memory=c(x["memory"]),
                memory_key_padding_mask=x["pad"], memory_level_start_index=x["lsi"], memory_spatial_shapes=x["shapes"],
                self_attn_mask=x["attn_mask"])
    return enc, dec, e.float(), d.float()


def main():
    mu = ref_shim.load_unipose()
    x = inputs()
    enc, dec, e32, d32 = run(mu, torch.float32, x)
    _, _, e16, d16 = run(mu, torch.bfloat16, x)
    np.savez_compressed(os.path.join(HERE, "mod_unipose_layers.npz"),
                        enc_keys=json.dumps(key_shapes(enc)), dec_keys=json.dumps(key_shapes(dec)),
                        enc_f32=e32[:, ::SUB].numpy(), enc_refbf16=e16[:, ::SUB].numpy(), dec_f32=d32.numpy(),
                        dec_refbf16=d16.numpy(), sub=np.int64(SUB))
    for n, a, b in (("enc", e32, e16), ("dec", d32, d16)):
        print(n, tuple(a.shape), float(a.abs().mean()), "bf16 rel_l2", float((a - b).norm() / a.norm()))


if __name__ == "__main__":
    main()

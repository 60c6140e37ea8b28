"""Run the two GDINO bi-attention shapes (head_dim 256) and one Swin window shape once each -- target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visionllm_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)


def run(B, Tq, Tk, H, D, bias_nb=0):
    q = torch.randn(B, Tq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Tk, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Tk, H, D, device="cuda", generator=g).bfloat16()
    kw = {"attn_bias": torch.randn(bias_nb, H, Tq, Tk, device="cuda", generator=g)} if bias_nb else \
        {"key_mask": torch.ones(B, Tk, dtype=torch.bool, device="cuda")}
    for _ in range(2):
        ops.attention(q, k, v, **kw)
    torch.cuda.synchronize()


run(8, 21760, 80, 4, 256)
run(8, 80, 21760, 4, 256)
run(8 * 1369, 49, 49, 3, 32, bias_nb=1369)

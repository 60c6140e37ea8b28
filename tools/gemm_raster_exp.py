"""GPU experiment: DRAM traffic / time of the GEMM vs rasterisation group size (run under ncu for dram bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visionllm_b200 import ops, _lib
SH = {"vit_fc2": (41000, 3200, 12800), "vit_qkv": (41000, 9600, 3200), "vit_fc1": (41000, 12800, 3200),
      "llm_down": (12288, 4096, 11008), "llm_gateup": (12288, 22016, 4096)}
for name, (M, N, K) in SH.items():
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for gm in (0, 8, 4, 2, 1):
        _lib.lib().vllm_gemm_set_group_m(gm)
        for _ in range(2):
            ops.linear(x, w, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.linear(x, w, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} group_m={gm} {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    _lib.lib().vllm_gemm_set_group_m(0)
    del x, w, out

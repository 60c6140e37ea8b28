"""One launch of the GDINO FFN1 GEMM (174080 x 2048 x 256, + bias + ReLU, cta_group::1 path) between cudaProfilerStart/Stop,
for `ncu --set full --import-source on --profile-from-start off` (source-level stall reasons of the short-K case)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from visionllm_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(174080, 256, device=dev, generator=g) * 0.5).bfloat16()
w = (torch.randn(2048, 256, device=dev, generator=g) * 0.05).bfloat16()
b = torch.zeros(2048, device=dev).bfloat16()
for _ in range(3):
    ops.linear(x, w, bias=b, act="relu")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.linear(x, w, bias=b, act="relu")
e1.record()
torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 10)
torch.cuda.profiler.start()
ops.linear(x, w, bias=b, act="relu")
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""GPU: time every MSDA kernel variant at the BASELINE cfg-2b shapes (CUDA events, inputs > L2)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_workloads as B
import visionllm_b200.msda as ext
from visionllm_b200 import _lib

dev = torch.device("cuda", 0)
N = 8
value, shapes, lsi, loc, attw = B.msda_encoder_inputs(torch, N, dev, 1234)
hs = shapes.cpu()
S = value.shape[1]
res = {}


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


enc_bytes = (value[0].numel() + loc[0].numel() + attw[0].numel() + S * 256) * 4
for sigma in (0.02, 0.005, 0.08):
    _, _, _, loc_s, _ = B.msda_encoder_inputs(torch, N, dev, 1234, sigma=sigma)
    for v in (0, 1, 2, 3, 4):
        _lib.lib().vllm_msda_set_variant(v)
        ms = timeit(lambda: ext.ms_deform_attn_forward(value, shapes, lsi, loc_s, attw, 64, host_shapes=hs))
        res[f"enc_sigma{sigma}_variant{v}"] = {"ms": ms, "us_per_img": ms * 1e3 / N, "GBps": enc_bytes * N / ms / 1e6}
    _lib.lib().vllm_msda_set_variant(0)
    ms = timeit(lambda: ext.ms_deform_attn_forward(value, shapes, lsi, loc_s, attw, 64, flags=1), iters=3)
    res[f"enc_sigma{sigma}_strict"] = {"ms": ms, "us_per_img": ms * 1e3 / N, "GBps": enc_bytes * N / ms / 1e6}
    del loc_s

g = torch.Generator(device=dev).manual_seed(5)
for Lq in (900, 100):
    cxcy = torch.rand(N, Lq, 1, 1, 1, 2, device=dev, generator=g)
    wh = torch.rand(N, Lq, 1, 1, 1, 2, device=dev, generator=g) * 0.45 + 0.05
    off = torch.randn(N, Lq, 8, 4, 4, 2, device=dev, generator=g) * 0.25
    locd = (cxcy + off * wh).contiguous()
    wd = torch.softmax(torch.randn(N, Lq, 8, 16, device=dev, generator=g), -1).view(N, Lq, 8, 4, 4).contiguous()
    b = (value[0].numel() + locd[0].numel() + wd[0].numel() + Lq * 256) * 4
    for v in (0, 1):
        _lib.lib().vllm_msda_set_variant(v)
        ms = timeit(lambda: ext.ms_deform_attn_forward(value, shapes, lsi, locd, wd, 64, host_shapes=hs), iters=20)
        res[f"dec{Lq}_variant{v}"] = {"ms": ms, "us_per_img": ms * 1e3 / N, "GBps_full_value": b * N / ms / 1e6}
_lib.lib().vllm_msda_set_variant(0)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/msda_sweep.json", "w"), indent=1)

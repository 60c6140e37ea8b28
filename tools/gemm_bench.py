"""GPU: TFLOP/s of the tcgen05 GEMM (both variants) at the hot-path shapes; torch.matmul (cuBLAS) beside it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visionllm_b200 import ops, _lib

SHAPES = {"vit_qkv": (41000, 9600, 3200), "vit_proj": (41000, 3200, 3200), "vit_fc1": (41000, 12800, 3200),
          "vit_fc2": (41000, 3200, 12800), "llm_qkv": (12288, 12288, 4096), "llm_gateup": (12288, 22016, 4096),
          "llm_down": (12288, 4096, 11008), "gdino_ffn1": (21760 * 8, 2048, 256), "sq8k": (8192, 8192, 8192)}
res = {}


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, (M, N, K) in SHAPES.items():
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    r = {}
    for v in (1, 2):
        _lib.lib().vllm_gemm_set_variant(v)
        try:
            ms = timeit(lambda: ops.linear(x, w, out=out))
            r[f"cg{v}_tflops"] = fl / ms / 1e9
        except Exception as e:
            r[f"cg{v}_error"] = str(e)
    _lib.lib().vllm_gemm_set_variant(0)
    ms = timeit(lambda: torch.matmul(x, w.T, out=out))
    r["cublas_tflops"] = fl / ms / 1e9
    res[name] = r
    print(name, r, flush=True)
    del x, w, out
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)

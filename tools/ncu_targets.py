"""The launches this round's `ncu --set full` capture is made of, in a fixed order, between cudaProfilerStart/Stop:
    1. tcgen05 GEMM  41000 x 12800 x 3200   (InternViT-6B fc1 of one 8-pair step, + bias + GELU epilogue)
    2. tcgen05 GEMM  41000 x 3200 x 12800   (fc2, + bias + LayerScale + residual epilogue)
    3. tcgen05 GEMM  41000 x 9600 x 3200    (qkv)
    4. MSDA window kernel, fp32 value / fp32 out, cfg-2b encoder shape (N=8, S=Lq=21760)
    5. MSDA window kernel, bf16 value / bf16 out, same shape
    6. MSDA global warp-gather kernel (variant 4), fp32 -- the r1 kernel, for comparison
    7. MSDA global patch kernel (variant 0 = the fp32 default)

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_targets \
      python tools/ncu_targets.py
  python tools/ncu_targets.py --summarise gpurun_out/r2_targets.ncu-rep      (here; writes profiles/r2_*_ncu.json)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LABELS = [("gemm", "41000x12800x3200"), ("gemm", "41000x3200x12800"), ("gemm", "41000x9600x3200"),
          ("msda", "window fp32 (variant 33)"), ("msda", "window bf16 (default)"), ("msda", "global warp-gather fp32 (variant 4)"),
          ("msda", "global patch kernel fp32 (default)")]


def run():
    import torch
    import bench_workloads as B
    import visionllm_b200.msda as ext
    from visionllm_b200 import _lib, ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    M = 41000
    x32 = (torch.randn(M, 3200, device=dev, generator=g) * 0.5).bfloat16()
    x128 = (torch.randn(M, 12800, device=dev, generator=g) * 0.5).bfloat16()
    w1 = (torch.randn(12800, 3200, device=dev, generator=g) * 0.02).bfloat16()
    w2 = (torch.randn(3200, 12800, device=dev, generator=g) * 0.02).bfloat16()
    wq = (torch.randn(9600, 3200, device=dev, generator=g) * 0.02).bfloat16()
    b1, b2 = torch.zeros(12800, device=dev).bfloat16(), torch.zeros(3200, device=dev).bfloat16()
    ls = torch.full((3200,), 0.1, device=dev).bfloat16()
    value, shapes, lsi, loc, attw = B.msda_encoder_inputs(torch, 8, dev, 1234)
    hs = shapes.cpu()
    v16 = value.bfloat16()
    L_ = _lib.lib()

    def targets():
        ops.linear(x32, w1, bias=b1, act="gelu")
        ops.linear(x128, w2, bias=b2, colscale=ls, residual=x32)
        ops.linear(x32, wq)
        L_.vllm_msda_set_variant(33)                           # fp32 rows: the window kernel is opt-in
        ext.ms_deform_attn_forward(value, shapes, lsi, loc, attw, 64, host_shapes=hs)
        L_.vllm_msda_set_variant(0)
        ext.ms_deform_attn_forward_bf16(v16, shapes, lsi, loc, attw)
        L_.vllm_msda_set_variant(4)
        ext.ms_deform_attn_forward(value, shapes, lsi, loc, attw, 64, host_shapes=hs)
        L_.vllm_msda_set_variant(0)
        ext.ms_deform_attn_forward(value, shapes, lsi, loc, attw, 64, host_shapes=hs)

    for _ in range(2):
        targets()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    targets()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def summarise(rep):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ncu_summary
    tmp = os.path.join(ROOT, "gpurun_out", "_targets_summary.json")
    ncu_summary.full(rep, tmp)
    d = json.load(open(tmp))
    ls = d["launches"]
    assert len(ls) == len(LABELS), (len(ls), [x["kernel"] for x in ls])
    out = {"gemm": [], "msda": []}
    for (fam, tag), launch in zip(LABELS, ls):
        launch["shape" if fam == "gemm" else "case"] = tag
        out[fam].append(launch)
    note = ("ncu --set full --clock-control none, one launch each after 2 warm-up rounds (tools/ncu_targets.py); "
            "durations under ncu are not bench values")
    json.dump({"source": rep, "note": note, "launches": out["gemm"]}, open(os.path.join(ROOT, "profiles", "r2_gemm_ncu.json"), "w"), indent=1)
    json.dump({"source": rep, "note": note, "launches": out["msda"]}, open(os.path.join(ROOT, "profiles", "r2_msda_win_ncu.json"), "w"), indent=1)
    print("wrote profiles/r2_gemm_ncu.json, profiles/r2_msda_win_ncu.json")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run()

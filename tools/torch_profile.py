"""Kernel-level breakdown of one bench workload step with torch.profiler (CUDA activities): which kernels -- ours and
the torch glue between them -- the step spends its time in.  Usage: python tools/torch_profile.py <workload> [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench_workloads as W  # noqa: E402


def main():
    name = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    torch.cuda.set_device(0)
    wl = W.WORKLOADS[name](0, 1, torch.device("cuda:0"))
    wl.setup()
    for _ in range(3):
        wl.step_device()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        wl.step_device()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
            rows.append({"kernel": e.key[:140], "calls": e.count, "us": t})
    rows.sort(key=lambda r: -r["us"])
    total = sum(r["us"] for r in rows)
    print(f"total device time {total / 1e3:.2f} ms over {sum(r['calls'] for r in rows)} launches")
    for r in rows[:40]:
        print(f"{r['us'] / 1e3:9.3f} ms {100 * r['us'] / total:5.1f}% x{r['calls']:<5d} {r['kernel']}")
    # the torch-side glue, by the aten op (and input shapes) that launched it: self device time of CPU-side op events
    glue = []
    for e in prof.key_averages(group_by_input_shape=True):
        t = getattr(e, "self_device_time_total", None)
        if t is None:
            t = getattr(e, "self_cuda_time_total", 0)
        if e.device_type == torch.autograd.DeviceType.CPU and t > 0 and e.key.startswith("aten::"):
            glue.append({"op": e.key, "shapes": str(e.input_shapes)[:160], "calls": e.count, "us": t})
    glue.sort(key=lambda r: -r["us"])
    gl_total = sum(r["us"] for r in glue)
    print(f"aten ops: {gl_total / 1e3:.2f} ms self device time")
    for r in glue[:40]:
        print(f"{r['us'] / 1e3:9.3f} ms x{r['calls']:<4d} {r['op']:<28s} {r['shapes']}")
    if out:
        json.dump({"workload": name, "total_ms": total / 1e3, "kernels": rows[:80], "aten_ms": gl_total / 1e3, "aten_ops": glue[:80]},
                  open(out, "w"), indent=1)


if __name__ == "__main__":
    main()

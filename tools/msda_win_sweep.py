"""Time the MSDA encoder shape (BASELINE cfg 2b: N=8, S=Lq=21760, M=8, D=32, L=4, P=4) on the global-memory warp-gather
kernel and on the TMA-staged window kernel for a sweep of window geometries.  CUDA events, L2 flushed between launches.

    python tools/msda_win_sweep.py [--out profiles/r2_msda_window_sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench_workloads as B
    import visionllm_b200.msda as ext
    from visionllm_b200 import _lib
    L_ = _lib.lib()
    dev = torch.device("cuda", 0)
    value, shapes, lsi, loc, attw = B.msda_encoder_inputs(torch, 8, dev, 1234)
    hs = shapes.cpu()
    v16 = value.bfloat16()
    S = value.shape[1]
    peaks = B.measured_peaks()
    alg32 = (value[0].numel() + loc[0].numel() + attw[0].numel() + S * 256) * 4 * 8
    alg16 = (value[0].numel() * 2 + loc[0].numel() * 4 + attw[0].numel() * 4 + S * 256 * 2) * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    f32 = lambda: ext.ms_deform_attn_forward(value, shapes, lsi, loc, attw, 64, host_shapes=hs)  # noqa: E731
    b16 = lambda: ext.ms_deform_attn_forward_bf16(v16, shapes, lsi, loc, attw)  # noqa: E731
    rows = []

    def rec(name, variant, window, fn, alg, tma=0):
        L_.vllm_msda_set_variant(variant)
        L_.vllm_msda_set_window(*window)
        L_.vllm_msda_set_window_fill(tma)
        try:
            ms = t(fn)
        finally:
            L_.vllm_msda_set_variant(0)
            L_.vllm_msda_set_window(0, 0, 0)
            L_.vllm_msda_set_window_fill(-1)
        rows.append({"case": name, "fill_tma_level_mask": tma, "variant": variant, "window_ph_pw_halo": list(window), "ms": ms,
                     "GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peaks["hbm_gbs"]})
        print(rows[-1], flush=True)

    rec("fp32 global warp-gather (r1 kernel, new reduce-scatter)", 4, (0, 0, 0), f32, alg32)
    rec("fp32 global warp-gather, 8x16 patches", 1, (0, 0, 0), f32, alg32)
    for w in ((0, 0, 0), (8, 8, 5), (8, 8, 8), (8, 16, 8)):
        for mask in (15, 0, 1, 14):
            rec("fp32 window", 0, w, f32, alg32, tma=mask)
    rec("bf16 global warp-gather (r1 kernel)", 32, (0, 0, 0), b16, alg16)
    for w in ((0, 0, 0), (8, 16, 6), (8, 16, 10), (16, 16, 8), (8, 32, 8), (16, 32, 8)):
        for mask in (15, 0, 1, 14):
            rec("bf16 window", 0, w, b16, alg16, tma=mask)
    res = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peaks["hbm_gbs"], "shape": "N=8 S=Lq=21760 M=8 D=32 L=4 P=4",
           "alg_bytes_fp32": alg32, "alg_bytes_bf16": alg16, "rows": rows}
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

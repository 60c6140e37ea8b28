"""Workloads for bench.py (measurement infrastructure, not product code).

Each workload is one pass of (a part of) the hot path over one batch of
synthetic input.  `step_device` runs with inputs resident in
HBM; `step_e2e` goes through the same public API with HOST (pinned) buffers,
host<->device copies inside the timed region.
"""
import json
import os
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    # B200_PROFILING.md fallback
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


GDINO_LEVELS_1024 = [(128, 128), (64, 64), (32, 32), (16, 16)]   # strides 8..64 of a 1024x1024 image


def msda_encoder_inputs(torch, N, device, seed, shapes_l=GDINO_LEVELS_1024, M=8, D=32, P=4, sigma=0.02):
    """SURVEY 8d cfg 2b 'enc': queries are the pixels, refs = pixel centres + N(0, sigma) offsets."""
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = torch.tensor(shapes_l, dtype=torch.int64, device=device)
    L = len(shapes_l)
    S = sum(h * w for h, w in shapes_l)
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, D, device=device, generator=g)
    refs = []
    for (H, W) in shapes_l:
        ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                                torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
        refs.append(torch.stack(((xs + 0.5) / W, (ys + 0.5) / H), -1).reshape(-1, 2))
    ref_pts = torch.cat(refs, 0)[None, :, None, None, None, :]
    loc = (ref_pts + torch.randn(N, S, M, L, P, 2, device=device, generator=g) * sigma).contiguous()
    attw = torch.softmax(torch.randn(N, S, M, L * P, device=device, generator=g), -1).view(N, S, M, L, P)
    return value, shapes, lsi, loc, attw.contiguous()


class MsdaEncoderWorkload:
    """MSDA forward at the GDINO 1024^2 encoder shape (BASELINE cfg 2b): N=8, S=Lq=21760, M=8, D=32, L=4, P=4."""
    metric = "msda_encoder_layer_images_per_sec"
    unit = "images/s"
    dtype = "f32"
    N = 8

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device

    def setup(self):
        import torch
        import visionllm_b200.msda as ext
        self.torch, self.ext = torch, ext
        self.value, self.shapes, self.lsi, self.loc, self.attw = msda_encoder_inputs(
            torch, self.N, self.device, 1234 + self.rank)
        self.host_shapes = self.shapes.cpu()
        self.h_in = [t.cpu().pin_memory() for t in (self.value, self.loc, self.attw)]
        self.d_in = [torch.empty_like(t) for t in (self.value, self.loc, self.attw)]
        S = self.value.shape[1]
        self.h_out = torch.empty((self.N, S, 256), dtype=torch.float32).pin_memory()
        self.h2d_bytes = sum(t.numel() * 4 for t in self.h_in)
        self.d2h_bytes = self.h_out.numel() * 4
        # compulsory bytes per image: value + loc + attw once, out once
        self.alg_bytes_per_image = (self.value[0].numel() + self.loc[0].numel() + self.attw[0].numel()
                                    + S * 256) * 4

    def step_device(self):
        self.out = self.ext.ms_deform_attn_forward(self.value, self.shapes, self.lsi, self.loc, self.attw, 64,
                                                   host_shapes=self.host_shapes)

    def step_e2e(self):
        for d, h in zip(self.d_in, self.h_in):
            d.copy_(h, non_blocking=True)
        out = self.ext.ms_deform_attn_forward(self.d_in[0], self.shapes, self.lsi, self.d_in[1], self.d_in[2], 64,
                                              host_shapes=self.host_shapes)
        self.h_out.copy_(out, non_blocking=True)

    def units_per_step(self):
        return self.N

    def dominant_kernel_ms(self, steps):
        torch = self.torch
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); self.step_device(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps

    def roofline(self, kern_ms, peaks):
        ach = self.alg_bytes_per_image * self.N / (kern_ms * 1e-3) / 1e9
        return {"kernel": "msda_fwd_warp_kernel", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                "peak_source": peaks["source"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": self.alg_bytes_per_image * self.N}

    def config(self):
        return {"workload": "msda_fwd encoder shape (BASELINE cfg 2b): N=8 S=Lq=21760 M=8 D=32 L=4 P=4 fp32",
                "l2_policy": "inputs_exceed_l2 (624 MB per step > 126 MB L2)", "parallelism": f"dp{self.world}"}

    def extra(self):
        return {}


WORKLOADS = {"msda_encoder": MsdaEncoderWorkload}
DEFAULT_WORKLOAD = "msda_encoder"


# --------------------------------------------------------------------------------------
# CPU legs: the ONLY place bench code touches oracle/.
# --------------------------------------------------------------------------------------
def _cpu_msda_encoder(steps, warmup):
    import torch
    from oracle import msda_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    value, shapes, lsi, loc, attw = msda_encoder_inputs(torch, 1, torch.device("cpu"), 1234)
    for _ in range(warmup):
        O.forward_grid_sample(value, shapes, loc, attw)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward_grid_sample(value, shapes, loc, attw)
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "1 image per step at the same shape (S=Lq=21760, M=8, D=32, L=4, P=4, fp32), "
                      "oracle.forward_grid_sample = the reference's pure-PyTorch CPU path restated",
            "ms_per_step": dt * 1e3}


_CPU = {"msda_encoder": _cpu_msda_encoder}


def cpu_baseline(name):
    return _CPU[name](steps=3, warmup=1)


def run_reference_arm(name, n_gpus, steps, warmup):
    wl = WORKLOADS[name]
    cb = _CPU[name](steps=max(1, min(steps, 5)), warmup=max(1, min(warmup, 1)))
    return {"impl": "reference", "metric": wl.metric, "value": cb["value"], "unit": wl.unit, "n_gpus": n_gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": cb["sample"]}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

"""Deterministic, reference-free weights for module parity tests: the SAME function fills the reference module
(when generating goldens in the build container) and the B200 module (on the GPU box), keyed by the state-dict
names/shapes, which the golden file records so a key mismatch fails loudly."""
import torch


def seeded_state_dict(module, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, t in sorted(module.state_dict().items()):
        shape = tuple(t.shape)
        if not t.is_floating_point():
            out[name] = t.clone()
            continue
        r = torch.randn(shape, generator=g)
        last = name.split(".")[-1]
        if last in ("ls1", "ls2"):
            v = 0.1 + 0.02 * r
        elif "norm" in name and last == "weight":
            v = 1.0 + 0.1 * r
        elif len(shape) <= 1:
            v = 0.05 * r
        elif last in ("class_embedding", "position_embedding"):
            v = 0.2 * r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = r / fan_in ** 0.5
        out[name] = v.to(torch.bfloat16).to(torch.float32)      # bf16-representable values
    return out


def key_shapes(module):
    return sorted((k, list(v.shape)) for k, v in module.state_dict().items())

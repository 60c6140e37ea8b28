"""TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): fp32 torch stand-ins for the op contracts of
`visionllm_b200.ops` and `visionllm_b200.msda`, and a context manager that swaps them in so that the HOST LOGIC of the
B200 modules (masks, packing, residual wiring, index arithmetic) can run on CPU.

Who may use this: the `-m "not gpu"` tests (host-logic parity against reference goldens) and bench_workloads'
`cpu_baseline` leg of BASELINE cfg 1 (the reference's CPU forward, restated: same module graph, torch fp32 math on the
host cores).  The product package never imports it (tests/test_no_fallback_cpu.py enforces that); on a GPU box every
product op goes through libvllm_b200.so or raises.
"""
import contextlib

import torch
import torch.nn.functional as F

_ACT = {None: lambda z: z, "none": lambda z: z, "relu": torch.relu, "gelu": F.gelu, "silu": F.silu,
        "quick_gelu": lambda z: z * torch.sigmoid(1.702 * z)}


def linear(x, w, bias=None, act=None, colscale=None, residual=None, out_dtype=None, out=None, row_keep=None):
    y = F.linear(x.float(), w.float(), None if bias is None else bias.float())
    if act == "swiglu":                                   # gate / up rows interleaved in `w` (llama.py packed operand)
        y = F.silu(y[..., 0::2]) * y[..., 1::2]
    else:
        y = _ACT[act](y)
    if colscale is not None:
        y = y * colscale.float()
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    if row_keep is not None:                               # rows with False are stored as exact zeros (ops.linear)
        y = y.masked_fill(~row_keep.bool().reshape(y.shape[:-1])[..., None], 0.0)
    if out is not None:
        out.copy_(y)
        return out
    return y


def rmsnorm(x, weight, eps, out=None):
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()
    if out is not None:
        out.copy_(y)
        return out
    return y


def layernorm(x, weight, bias, eps, out=None, gelu=False, residual=None):
    y = F.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps)
    if gelu:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.float()
    if out is not None:
        out.copy_(y)
        return out
    return y


def rope_(x, cos, sin, heads, head_dim):
    """In-place rotate-half RoPE on the first `heads` heads of the packed rows x [tokens, >= heads*head_dim]."""
    t = x.shape[0]
    v = x[:, :heads * head_dim].reshape(t, heads, head_dim).float()
    c, s = cos.float()[:, None, :], sin.float()[:, None, :]
    half = head_dim // 2
    rot = torch.cat((-v[..., half:], v[..., :half]), -1)
    x[:, :heads * head_dim] = (v * c + rot * s).reshape(t, heads * head_dim).to(x.dtype)
    return x


def attention(q, k, v, causal=False, scale=None, seqlens=None, key_mask=None, attn_mask=None, attn_bias=None, out=None):
    B, Tq, H, D = q.shape
    Tk, Hkv = k.shape[1], k.shape[2]
    kf, vf = k.float(), v.float()
    if Hkv != H:
        kf, vf = kf.repeat_interleave(H // Hkv, 2), vf.repeat_interleave(H // Hkv, 2)
    s = (q.float().permute(0, 2, 1, 3) @ kf.permute(0, 2, 3, 1)) * (scale or D ** -0.5)
    if attn_bias is not None:
        s = s + attn_bias[torch.arange(B) % attn_bias.shape[0]]
    if attn_mask is not None:
        s = s.masked_fill(~attn_mask.bool().view(B, H, Tq, Tk), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(~key_mask.bool()[:, None, None, :], float("-inf"))
    if seqlens is not None:
        s = s.masked_fill(torch.arange(Tk)[None, None, None, :] >= seqlens[:, None, None, None], float("-inf"))
    if causal:
        s = s.masked_fill(torch.ones(Tq, Tk, dtype=torch.bool).triu(1 + Tk - Tq), float("-inf"))
    y = (torch.softmax(s, -1) @ vf.permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Tq, H * D)
    if out is not None:
        out.copy_(y)
        return out
    return y


def groupnorm_nhwc(x, w, b, groups, eps, relu=False):
    if x.dim() == 4:                                       # a [B, H, W, C] (possibly strided) view -> rows [B, H*W, C]
        x = x.reshape(x.shape[0], -1, x.shape[-1])
    y = F.group_norm(x.float().transpose(1, 2), groups, w.float(), b.float(), eps).transpose(1, 2)
    return torch.relu(y) if relu else y


def conv2d_s1_rows(x, w_rows, bias, kernel, padding, act=None, prepadded=False):
    Cout, C = w_rows.shape[0], x.shape[-1]
    w = w_rows.float().view(Cout, kernel, kernel, C).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None if bias is None else bias.float(), padding=0 if prepadded else padding)
    return _ACT[act](y.permute(0, 2, 3, 1))


def layernorm_gather(x, index, weight, bias, eps):
    h = F.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps)
    h = torch.cat((h, h.new_zeros(h.shape[0], 1, h.shape[2])), 1)
    return h.index_select(1, index.clamp(max=x.shape[1]))


def gather_rows(src, idx):
    return src[idx]


def ce_loss(logits, labels):
    return F.cross_entropy(logits.float(), labels.reshape(-1), ignore_index=-100)


def upsample_add_nhwc(top, lateral, pad=0):
    up = F.interpolate(top.float().permute(0, 3, 1, 2), size=lateral.shape[1:3], mode="bilinear", align_corners=False)
    y = lateral.float() + up.permute(0, 2, 3, 1)
    return F.pad(y, (0, 0, pad, pad, pad, pad)) if pad else y


def msda_forward(value, shapes, lsi, loc, w, step=64, **kw):
    from . import msda_oracle as O
    return O.forward_grid_sample(value.float(), shapes, loc.float(), w.float())


@contextlib.contextmanager
def patched():
    """Swap the stand-ins into visionllm_b200.ops / .msda for the duration of the block (CPU legs only)."""
    import visionllm_b200.msda as msda
    import visionllm_b200.ops as ops
    table = {"linear": linear, "rmsnorm": rmsnorm, "layernorm": layernorm, "rope_": rope_, "attention": attention,
             "groupnorm_nhwc": groupnorm_nhwc, "conv2d_s1_rows": conv2d_s1_rows, "upsample_add_nhwc": upsample_add_nhwc,
             "layernorm_gather": layernorm_gather, "ce_loss": ce_loss, "gather_rows": gather_rows}
    saved = {k: getattr(ops, k) for k in table}
    saved_msda = (msda.ms_deform_attn_forward, msda.ms_deform_attn_forward_bf16)
    try:
        for k, fn in table.items():
            setattr(ops, k, fn)
        msda.ms_deform_attn_forward = msda_forward
        msda.ms_deform_attn_forward_bf16 = lambda value, shapes, lsi, loc, w, out_dtype=None: msda_forward(value, shapes, lsi, loc, w)
        yield
    finally:
        for k, fn in saved.items():
            setattr(ops, k, fn)
        msda.ms_deform_attn_forward, msda.ms_deform_attn_forward_bf16 = saved_msda

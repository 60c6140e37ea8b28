"""ORACLE (test infrastructure): CPU multi-scale deformable attention.

Two restatements of the reference, both pinned by tests/golden/*.npz
(tests/test_oracle_golden.py):

* ``forward_kernel_semantics`` / ``sample_indices`` -- ctypes over
  oracle/liboracle.so (oracle/msda_oracle.c), the reference CUDA kernel's
  arithmetic (mmcv ms_deform_attn_cuda_kernel.cuh:17-64,200-254).  This is the
  bit-exact checker for sampling indices and for the strict CUDA kernel.
* ``forward_grid_sample`` -- the reference's CPU path in torch ops
  (mmcv/ops/multi_scale_deform_attn.py:100-159 == grounding_dino/
  modeling_ov_grounding_dino_mask_dn.py:607-643): per level, reshape value to
  [N*M, D, H, W], bilinear grid_sample (zeros padding, align_corners=False) at
  2*loc-1, weight and sum.  This is what `bench.py --impl reference` times.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle`")
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(_lib().oracle_num_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward_kernel_semantics(value, shapes, lsi, loc, attw):
    """numpy in / numpy out; fp32 or fp64 (dtype of value)."""
    value = np.ascontiguousarray(value)
    dt = value.dtype
    assert dt in (np.float32, np.float64)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attw = np.ascontiguousarray(attw, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = np.empty((N, Lq, M * D), dtype=dt)
    fn = _lib().oracle_msda_forward_f32 if dt == np.float32 else _lib().oracle_msda_forward_f64
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attw), _p(out), N, S, M, D, L, Lq, P)
    return out


def backward_kernel_semantics(value, shapes, lsi, loc, attw, grad_out):
    """numpy in / out: (grad_value, grad_loc, grad_attw); fp32 or fp64 (dtype of value)."""
    value = np.ascontiguousarray(value)
    dt = value.dtype
    loc = np.ascontiguousarray(loc, dtype=dt); attw = np.ascontiguousarray(attw, dtype=dt)
    grad_out = np.ascontiguousarray(grad_out, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64); lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    gv, gl, gw = np.zeros_like(value), np.empty_like(loc), np.empty_like(attw)
    fn = _lib().oracle_msda_backward_f32 if dt == np.float32 else _lib().oracle_msda_backward_f64
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attw), _p(grad_out), _p(gv), _p(gl), _p(gw), N, S, M, D, L, Lq, P)
    return gv, gl, gw


def sample_indices(shapes, loc):
    loc = np.ascontiguousarray(loc, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    L, P = loc.shape[-3], loc.shape[-2]
    n = loc.size // 2
    out = np.empty(loc.shape[:-1] + (3,), dtype=np.int32)
    _lib().oracle_msda_sample_indices_f32(_p(shapes), _p(loc), _p(out), ctypes.c_longlong(n), L, P)
    return out


def forward_grid_sample(value, shapes, loc, attw):
    """torch CPU tensors in / out -- the reference's pure-PyTorch CPU path, restated."""
    import torch
    import torch.nn.functional as F
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    hw = [(int(h), int(w)) for h, w in shapes.tolist()]
    grids = loc * 2 - 1
    acc = value.new_zeros((N * M, D, Lq))
    start = 0
    for lvl, (H, W) in enumerate(hw):
        feat = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        start += H * W
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        sampled = F.grid_sample(feat, g, mode="bilinear", padding_mode="zeros", align_corners=False)
        w = attw[:, :, :, lvl].permute(0, 2, 1, 3).reshape(N * M, 1, Lq, P)
        acc = acc + (sampled * w).sum(-1)
    return acc.view(N, M * D, Lq).transpose(1, 2).contiguous()

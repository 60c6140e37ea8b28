"""ORACLE (test infrastructure): ctypes wrapper of oracle/dcnv3_oracle.c (the reference DCNv3 CUDA kernel's
arithmetic restated in C).  Pinned by tests/golden/dcnv3_ref_*.npz (tests/test_oracle_golden.py)."""
import ctypes

import numpy as np

from .msda_oracle import _lib, _p


def forward(inp, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, offset_scale):
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    offset = np.ascontiguousarray(offset, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    N, H_in, W_in, _ = inp.shape
    _, H_out, W_out, _ = offset.shape
    out = np.empty((N, H_out, W_out, group * gc), dtype=np.float32)
    _lib().oracle_dcnv3_forward_f32(_p(inp), _p(offset), _p(mask), _p(out), N, H_in, W_in, H_out, W_out, group, gc,
                                    kh, kw, sh, sw, ph, pw, dh, dw, ctypes.c_float(offset_scale))
    return out


def backward(inp, offset, mask, grad_out, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, offset_scale, dtype=np.float32):
    """(grad_input, grad_offset, grad_mask) -- oracle_dcnv3_backward_f32 / _f64."""
    inp, offset, mask, grad_out = (np.ascontiguousarray(a, dtype=dtype) for a in (inp, offset, mask, grad_out))
    N, H_in, W_in, _ = inp.shape
    _, H_out, W_out, _ = offset.shape
    gi, go, gm = np.zeros_like(inp), np.zeros_like(offset), np.zeros_like(mask)
    fn = _lib().oracle_dcnv3_backward_f32 if dtype == np.float32 else _lib().oracle_dcnv3_backward_f64
    sc = ctypes.c_float(offset_scale) if dtype == np.float32 else ctypes.c_double(offset_scale)
    fn(_p(inp), _p(offset), _p(mask), _p(grad_out), _p(gi), _p(go), _p(gm), N, H_in, W_in, H_out, W_out, group, gc,
       kh, kw, sh, sw, ph, pw, dh, dw, sc)
    return gi, go, gm

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

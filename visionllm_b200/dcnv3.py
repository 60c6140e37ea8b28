"""Drop-in for the reference extension module ``DCNv3`` (visionllmv2/model/ops_dcnv3/src/vision.cpp:14-17):

    dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, group, group_channels, offset_scale, im2col_step) -> Tensor[N, H_out, W_out, G*C]

caller: ops_dcnv3/functions/dcnv3_func.py:39-58.  Install with ``sys.modules["DCNv3"] = visionllm_b200.dcnv3``
before ``functions/dcnv3_func.py`` is imported (it does ``import DCNv3`` at :16).  fp32, contiguous NHWC CUDA
tensors like the reference (dcnv3_cuda.cu:28-40 asserts); ``im2col_step`` is validated and ignored (one launch).
"""
import torch

from . import _lib

STRICT = 1


def dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                  group, group_channels, offset_scale, im2col_step=256, *, flags=0):
    for n, t in (("input", input), ("offset", offset), ("mask", mask)):
        if not t.is_contiguous():
            raise RuntimeError(f"{n} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{n} must be a CUDA tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"dcnv3_forward: {n} must be float32 (the reference module upcasts, "
                               "modules/dcnv3.py:331-340)")
    N, H_in, W_in, C = input.shape
    if C != group * group_channels:
        raise RuntimeError("input channels must equal group * group_channels")
    H_out = (H_in + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    W_out = (W_in + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    K = kernel_h * kernel_w
    if tuple(offset.shape) != (N, H_out, W_out, group * K * 2) or tuple(mask.shape) != (N, H_out, W_out, group * K):
        raise RuntimeError("offset / mask shapes do not match the output size")
    step = min(N, int(im2col_step)) if N > 0 else 1
    if step <= 0 or (N > 0 and N % step != 0):
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
    out = torch.empty((N, H_out, W_out, C), dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(input.device):
        rc = _lib.lib().vllm_dcnv3_forward_f32(
            input.data_ptr(), offset.data_ptr(), mask.data_ptr(), out.data_ptr(), N, H_in, W_in, H_out, W_out, group,
            group_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
            float(offset_scale), int(flags), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "dcnv3_forward")
    return out


def dcnv3_backward(*args, **kwargs):
    raise NotImplementedError("dcnv3_backward is a SURVEY 8(f) 'next' row (forward-only hot path this round)")

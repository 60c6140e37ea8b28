"""Drop-in for the reference extension module ``DCNv3`` (visionllmv2/model/ops_dcnv3/src/vision.cpp:14-17):

    dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, group, group_channels, offset_scale, im2col_step) -> Tensor[N, H_out, W_out, G*C]

    dcnv3_backward(same 14 arguments, grad_output, im2col_step) -> [grad_input, grad_offset, grad_mask]

caller: ops_dcnv3/functions/dcnv3_func.py:39-77.  Install with ``sys.modules["DCNv3"] = visionllm_b200.dcnv3``
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
    from . import ops
    with torch.cuda.device(input.device), ops._Prof("dcnv3", 0.0, 4.0 * (input.numel() + offset.numel() + mask.numel()
                                                                          + out.numel())):
        rc = _lib.lib().vllm_dcnv3_forward_f32(
            input.data_ptr(), offset.data_ptr(), mask.data_ptr(), out.data_ptr(), N, H_in, W_in, H_out, W_out, group,
            group_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
            float(offset_scale), int(flags), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "dcnv3_forward")
    return out


def dcnv3_backward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                   group, group_channels, offset_scale, grad_output, im2col_step=256):
    """-> [grad_input, grad_offset, grad_mask] (ops_dcnv3/src/dcnv3.h:40-59; caller functions/dcnv3_func.py:60-77)."""
    for n, t in (("input", input), ("offset", offset), ("mask", mask), ("grad_output", grad_output)):
        if not t.is_contiguous():
            raise RuntimeError(f"{n} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{n} must be a CUDA tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"dcnv3_backward: {n} must be float32")
    N, H_in, W_in, C = input.shape
    if C != group * group_channels:
        raise RuntimeError("input channels must equal group * group_channels")
    H_out = (H_in + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    W_out = (W_in + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1
    K = kernel_h * kernel_w
    if tuple(offset.shape) != (N, H_out, W_out, group * K * 2) or tuple(mask.shape) != (N, H_out, W_out, group * K):
        raise RuntimeError("offset / mask shapes do not match the output size")
    if grad_output.numel() != N * H_out * W_out * C:
        raise RuntimeError("grad_output shape does not match the output size")
    step = min(N, int(im2col_step)) if N > 0 else 1
    if step <= 0 or (N > 0 and N % step != 0):
        raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
    grad_input, grad_offset, grad_mask = torch.zeros_like(input), torch.empty_like(offset), torch.empty_like(mask)
    if grad_output.numel():
        with torch.cuda.device(input.device):
            rc = _lib.lib().vllm_dcnv3_backward_f32(
                input.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_output.data_ptr(), grad_input.data_ptr(),
                grad_offset.data_ptr(), grad_mask.data_ptr(), N, H_in, W_in, H_out, W_out, group, group_channels,
                kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, float(offset_scale),
                torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "dcnv3_backward")
    return [grad_input, grad_offset, grad_mask]


class DCNv3Function(torch.autograd.Function):
    """The autograd wrapper the reference defines (ops_dcnv3/functions/dcnv3_func.py:24-77), on this module's ops."""

    @staticmethod
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                group, group_channels, offset_scale, im2col_step):
        ctx.args = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, group_channels,
                    offset_scale)
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(input, offset, mask)
        return dcnv3_forward(input, offset, mask, *ctx.args, im2col_step)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        gi, go, gm = dcnv3_backward(input, offset, mask, *ctx.args, grad_output.contiguous(), ctx.im2col_step)
        return (gi, go, gm) + (None,) * 12

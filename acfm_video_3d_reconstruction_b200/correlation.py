"""Correlation cost volume — mirror of the reference's `correlation_package/correlation.py` (its only native extension,
used by the frozen MaskFlowNet: /root/reference/multiframe/data/optical_flow/model/MaskFlownet.py:116,416), on
acfm_correlation_fwd/bwd (csrc/correlation.cu).  Same class names, constructor arguments and call signature:

    corr = Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    cost = corr(feat1, feat2)          # (B,C,H,W) x2 -> (B, 81, H, W)

CUDA fp32 tensors only (the reference also dispatches half; its flow network runs in fp32); no CPU fallback.
"""
import ctypes

import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import _lib


def out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    """(channels, outH, outW) of the cost volume (correlation_cuda.cc:24-32)."""
    c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    st = _lib.lib().acfm_correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                               ctypes.byref(c), ctypes.byref(h), ctypes.byref(w))
    _lib.check(st, "acfm_correlation_out_shape")
    return c.value, h.value, w.value


def _check(x, name):
    _lib.require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError(f"Correlation: {name} must be a (B,C,H,W) float32 tensor, got {tuple(x.shape)} {x.dtype}")
    return x.contiguous()


class CorrelationFunction(Function):
    """correlation.py:6-51 (same argument order as the reference's Function.apply call)."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        if corr_multiply != 1:
            raise ValueError("Correlation: corr_multiply != 1 is not implemented (the reference's kernels ignore it too)")
        input1, input2 = _check(input1, "input1"), _check(input2, "input2")
        if input1.shape != input2.shape:
            raise ValueError("Correlation: inputs must have the same shape")
        ctx.save_for_backward(input1, input2)
        ctx.cfg = (int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
        B, C, H, W = input1.shape
        ch, oh, ow = out_shape(H, W, *ctx.cfg)
        output = torch.empty((B, ch, oh, ow), dtype=torch.float32, device=input1.device)
        with torch.cuda.device(input1.device):
            st = _lib.lib().acfm_correlation_fwd(_lib.ptr(input1), _lib.ptr(input2), B, C, H, W, *ctx.cfg, _lib.ptr(output),
                                                 _lib.stream_of(input1))
        _lib.check(st, "acfm_correlation_fwd")
        _lib.count()
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        B, C, H, W = input1.shape
        grad_output = _check(grad_output, "grad_output")
        g1 = torch.empty_like(input1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(input2) if ctx.needs_input_grad[1] else None
        if g1 is not None or g2 is not None:
            with torch.cuda.device(input1.device):
                st = _lib.lib().acfm_correlation_bwd(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output), B, C, H, W, *ctx.cfg,
                                                     _lib.ptr(g1), _lib.ptr(g2), _lib.stream_of(input1))
            _lib.check(st, "acfm_correlation_bwd")
            _lib.count((g1 is not None) + (g2 is not None))
        return g1, g2, None, None, None, None, None, None


class Correlation(Module):
    """correlation.py:54-74."""

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super(Correlation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1.contiguous(), input2.contiguous(), self.pad_size, self.kernel_size,
                                         self.max_displacement, self.stride1, self.stride2, self.corr_multiply)

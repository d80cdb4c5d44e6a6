"""oracle/correlation_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement in plain torch (any device, fp32 or fp64) of the reference's correlation cost volume
(/root/reference/multiframe/data/optical_flow/model/correlation_package/correlation_cuda_kernel.cu:73-147 forward,
host sizes correlation_cuda.cc:18-40).  Differentiable, so its autograd is the gradient truth for the backward.

PINNED on the GPU box against the reference's OWN extension: oracle/build_ref_correlation.sh compiles the reference's
correlation_cuda.cc / correlation_cuda_kernel.cu where they lie under /root/reference (sm_100a, with a three-line
compatibility shim for current PyTorch headers) into oracle/_ref/correlation_cuda.so; tests/test_correlation_gpu.py loads
it with `load_reference_extension()` and compares it with this restatement and with the product kernel.
"""
import importlib.util
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "correlation_cuda.so")


def out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    kr = (kernel_size - 1) // 2
    border = kr + max_displacement
    dr = max_displacement // stride2
    D = 2 * dr + 1
    return D * D, math.ceil((H + 2 * pad_size - 2 * border) / stride1), math.ceil((W + 2 * pad_size - 2 * border) / stride1)


def correlation(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2):
    """(B,C,H,W) x2 -> (B, D*D, outH, outW): mean over the k x k x C window of P1 * shifted P2 (correlation_cuda_kernel.cu:103-144)."""
    B, C, H, W = input1.shape
    kr = (kernel_size - 1) // 2
    dr = max_displacement // stride2
    ch, oh, ow = out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    p1 = torch.nn.functional.pad(input1, (pad_size,) * 4)
    p2 = torch.nn.functional.pad(input2, (pad_size,) * 4)
    ys = torch.arange(oh, device=input1.device) * stride1 + max_displacement
    xs = torch.arange(ow, device=input1.device) * stride1 + max_displacement
    outs = []
    for tj in range(-dr, dr + 1):
        for ti in range(-dr, dr + 1):
            acc = 0
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    a = p1[:, :, (ys + j)[:, None], (xs + i)[None, :]]
                    b = p2[:, :, (ys + j + tj * stride2)[:, None], (xs + i + ti * stride2)[None, :]]
                    acc = acc + (a * b).sum(1)
            outs.append(acc / (kernel_size * kernel_size * C))
    return torch.stack(outs, 1)


def load_reference_extension():
    """The reference's own compiled extension (module with .forward / .backward), or None if it was not built."""
    if not os.path.exists(REF_SO):
        return None
    spec = importlib.util.spec_from_file_location("correlation_cuda", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_forward(mod, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2):
    """correlation.py:19-34 (CorrelationFunction.forward) around the compiled reference extension."""
    rbot1, rbot2, output = input1.new_empty(0), input2.new_empty(0), input1.new_empty(0)
    mod.forward(input1.contiguous(), input2.contiguous(), rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2, 1)
    return output


def reference_backward(mod, input1, input2, grad_output, pad_size, kernel_size, max_displacement, stride1, stride2):
    """correlation.py:36-51 (CorrelationFunction.backward)."""
    rbot1, rbot2 = input1.new_empty(0), input2.new_empty(0)
    g1, g2 = input1.new_empty(0), input2.new_empty(0)
    mod.backward(input1.contiguous(), input2.contiguous(), rbot1, rbot2, grad_output.contiguous(), g1, g2, pad_size, kernel_size,
                 max_displacement, stride1, stride2, 1)
    return g1, g2

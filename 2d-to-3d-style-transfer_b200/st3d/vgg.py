"""VGG-19 feature extractor plumbing.  The convolutions stay on torch / cuDNN (out of scope of the kernels
in this repo, BASELINE.json); this module only picks cuDNN's better entry points:

* channels_last activations: cuDNN's tensor-core convolutions are NHWC kernels, with NCHW tensors torch
  wraps every convolution in nchw<->nhwc transposes (21 % of an iteration);
* cuDNN's fused convolution + bias + ReLU (`torch.cudnn_convolution_relu`,
  cudnnConvolutionBiasActivationForward) instead of convolution, broadcast bias add and in-place ReLU as
  three kernels: bit-identical output, 2.55 ms instead of 4.85 ms per forward of 8 x 512^2 images.

* the four MaxPool2d(2, 2) modules run libst3d's NHWC pooling kernels (csrc/pool.cu): no int64 argmax indices
  written forward and read backward, and the backward also applies the ReLU mask of the layer in front of the
  pool, so that layer's separate ReLU-backward pass disappears (bit-identical gradients).

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
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups, premasked):
        y = torch.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
        ctx.conf = (stride, padding, dilation, groups)
        # the only consumer is a libst3d pool whose backward already applies this layer's ReLU mask
        ctx.premasked = bool(premasked) and _ops().maxpool_supported(y)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, weight, y = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        # ReLU backward from the saved output (skipped when the pool behind this layer has done it already)
        g = grad_y if ctx.premasked else torch.ops.aten.threshold_backward(grad_y, y, 0.0)
        need = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]]
        gx, gw, gb = torch.ops.aten.convolution_backward(g, x, weight, [weight.shape[0]], list(stride), list(padding),
                                                         list(dilation), False, [0, 0], groups, need)
        return gx, gw, gb, None, None, None, None, None


def _ops():
    from . import ops
    return ops


class _ConvReLUStyleTapFn(torch.autograd.Function):
    """A FusedConvReLU whose activation is a style tap (style_transfer.py:21-26 + losses.py:35-39 for that layer):
    returns (y, layer_loss) with y = relu(conv(x) + b) and layer_loss = mean((y y^T - target)^2) / (C^2 H^2).

    Keeping the tap inside the layer's Function lets the backward do in ONE kernel what autograd spreads over
    three: the Gram backward dF = (dG + dG^T) y, its addition to the gradient that reaches y through the rest of
    the network, and the ReLU backward of the sum -- st3d_gram_backward with ST3D_GRAM_ACCUMULATE | ST3D_GRAM_RELU_MASK."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, stride, padding, dilation, groups, precision, acc=None, loss_weight=1.0):
        """acc None: returns (y, this layer's loss term).  acc (1,) float32: the kernel ADDS loss_weight x term to it in
        place and (y, acc) is returned -- a whole perceptual loss is then one accumulator, not a tree of scalar kernels."""
        ops = _ops()
        y = torch.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
        B, C, H, W = y.shape
        scale = float(loss_weight) / (B * C * C) / (float(C) ** 2 * float(H) ** 2)
        loss = acc if acc is not None else torch.zeros(1, device=y.device, dtype=torch.float32)
        dgram, _ = ops.gram_mse_forward(y, target.detach(), scale, loss, precision=precision)
        ctx.conf = (stride, padding, dilation, groups, precision)
        ctx.save_for_backward(x, weight, y, dgram)
        ctx.set_materialize_grads(False)
        if acc is not None:
            ctx.mark_dirty(acc)
            return y, acc
        return y, loss.reshape(())

    @staticmethod
    def backward(ctx, grad_y, grad_loss):
        return _ConvReLUStyleTapFn._backward(ctx, grad_y, grad_loss) + (grad_loss, None)   # acc: passed through

    @staticmethod
    def _backward(ctx, grad_y, grad_loss):
        ops = _ops()
        x, weight, y, dgram = ctx.saved_tensors
        stride, padding, dilation, groups, precision = ctx.conf
        if grad_loss is None and grad_y is None:
            return (None,) * 9
        if grad_loss is None:                       # the tap's loss term is unused: plain ReLU backward
            g = torch.ops.aten.threshold_backward(grad_y, y, 0.0)
        else:
            out = None
            # The Gram kernel's epilogue adds into the INCOMING gradient buffer (no third pass over the largest tensors of
            # the step).  Autograd does not in general allow a backward to write its grad input: the buffer could be
            # shared with y.retain_grad(), a tensor hook, or another consumer's accumulation.  Here y's only other
            # consumer is the next VGG module, whose backward (cuDNN dgrad / k_maxpool_bwd) returns a fresh tensor that
            # nothing else holds, and perceptual_loss_of_images never exposes y -- callers that do hold y (get_features)
            # go through _StyleLayerFn, which allocates.
            if grad_y is not None:                  # accumulate into the incoming gradient when its layout allows it
                if grad_y.stride() == y.stride() and grad_y.dtype == torch.float32:
                    out = grad_y
                else:
                    out = torch.empty_like(y)
                    out.copy_(grad_y)
            g = ops.gram_backward(y, dgram, 1.0, out=out, accumulate=out is not None, precision=precision,
                                  scale_tensor=grad_loss, relu_mask=True, symmetric_dgram=True)
        need = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]]
        gx, gw, gb = torch.ops.aten.convolution_backward(g, x, weight, [weight.shape[0]], list(stride), list(padding),
                                                         list(dilation), False, [0, 0], groups, need)
        return gx, gw, gb, None, None, None, None, None, None


class _ConvReLUContentTapFn(torch.autograd.Function):
    """A FusedConvReLU whose activation is the CONTENT tap (conv4_2; losses.py:31): returns (y, mean((y - c)^2)).
    The backward is one kernel, st3d_mse_tap_backward: the MSE gradient 2 (y - c) / n, its sum with the gradient arriving
    from the layers behind y, and the ReLU backward of the sum -- autograd runs a multiply, an add and a threshold pass
    over the 67 MB activation for the same thing."""

    @staticmethod
    def forward(ctx, x, weight, bias, content, stride, padding, dilation, groups, acc=None, loss_weight=1.0):
        """acc: as in _ConvReLUStyleTapFn.forward."""
        ops = _ops()
        y = torch.cudnn_convolution_relu(x, weight, bias, stride, padding, dilation, groups)
        c = content.detach()
        if c.stride() != y.stride() or c.dtype != torch.float32:       # the kernels walk y and c in one element order
            c = torch.empty_like(y).copy_(c)
        loss = acc if acc is not None else torch.zeros(1, device=y.device, dtype=torch.float32)
        ctx.scale = float(loss_weight) / max(y.numel(), 1)
        ops.mse_forward(y, c, ctx.scale, loss, want_grad=False)
        ctx.conf = (stride, padding, dilation, groups)
        ctx.save_for_backward(x, weight, y, c)
        ctx.set_materialize_grads(False)
        if acc is not None:
            ctx.mark_dirty(acc)
            return y, acc
        return y, loss.reshape(())

    @staticmethod
    def backward(ctx, grad_y, grad_loss):
        ops = _ops()
        x, weight, y, c = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        if grad_loss is None and grad_y is None:
            return (None,) * 10
        if grad_loss is None:                       # the tap's loss term is unused: plain ReLU backward
            g = torch.ops.aten.threshold_backward(grad_y, y, 0.0)
        else:
            if grad_y is not None and (grad_y.stride() != y.stride() or grad_y.dtype != torch.float32):
                grad_y = torch.empty_like(y).copy_(grad_y)
            g = ops.mse_tap_backward(y, c, grad_y, ctx.scale, scale_tensor=grad_loss)
        need = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]]
        gx, gw, gb = torch.ops.aten.convolution_backward(g, x, weight, [weight.shape[0]], list(stride), list(padding),
                                                         list(dilation), False, [0, 0], groups, need)
        return gx, gw, gb, None, None, None, None, None, grad_loss, None     # acc: passed through


class _MaxPool2x2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, relu_mask):
        ctx.save_for_backward(x)
        ctx.relu_mask = relu_mask
        return _ops().maxpool2x2_forward(x)

    @staticmethod
    def backward(ctx, grad_y):
        (x,) = ctx.saved_tensors
        return _ops().maxpool2x2_backward(x, grad_y, ctx.relu_mask), None


class FusedMaxPool(nn.Module):
    """MaxPool2d(2, 2) on libst3d's NHWC kernels when the input allows it (CUDA fp32 channels_last, even sides,
    C % 4 == 0), torch's CUDA pooling for other CUDA shapes; CPU tensors raise.  `after_relu`: the input is a post-ReLU activation, so the backward may
    apply that ReLU's mask in the same pass (idempotent with a ReLU backward done elsewhere)."""

    def __init__(self, pool: nn.MaxPool2d, after_relu: bool = False):
        super().__init__()
        self.pool, self.after_relu = pool, bool(after_relu)

    @staticmethod
    def accepts(pool) -> bool:
        def pair(v):
            return (v, v) if isinstance(v, int) else tuple(v)
        return (isinstance(pool, nn.MaxPool2d) and pair(pool.kernel_size) == (2, 2) and pair(pool.stride) == (2, 2)
                and pair(pool.padding) == (0, 0) and pair(pool.dilation) == (1, 1) and not pool.ceil_mode
                and not pool.return_indices)

    def forward(self, x):
        if _ops().maxpool_supported(x):
            return _MaxPool2x2Fn.apply(x, self.after_relu)
        if not x.is_cuda:
            raise RuntimeError("FusedMaxPool: CPU tensor -- the fused VGG runs on CUDA only (libst3d has no CPU path)")
        return self.pool(x)     # CUDA tensor of a shape / layout the NHWC kernels do not take: torch's CUDA pooling


class FusedConvReLU(nn.Module):
    """Conv2d + ReLU as one cuDNN call; shares (does not copy) the parameters of the wrapped Conv2d."""

    def __init__(self, conv: nn.Conv2d):
        super().__init__()
        if conv.padding_mode != "zeros" or isinstance(conv.padding, str):
            raise ValueError("FusedConvReLU supports zero padding given as integers")
        self.conv = conv
        self.feeds_masking_pool = False   # set by fuse_vgg_features: next real module is a FusedMaxPool(after_relu)
        # cleared by st3d.losses.get_features around the call when it knows nothing else reads this activation; any
        # other caller keeps the conservative default (this layer's ReLU backward is always applied)
        self.tapped = True

    def forward(self, x):
        c = self.conv
        if not x.is_cuda:
            raise RuntimeError("FusedConvReLU: CPU tensor -- the fused VGG runs on CUDA only (cuDNN's fused entry point; "
                               "libst3d has no CPU path)")
        bias = c.bias if c.bias is not None else torch.zeros(c.out_channels, device=x.device, dtype=x.dtype)
        return _ConvBiasReLUFn.apply(x, c.weight, bias, tuple(c.stride), tuple(c.padding), tuple(c.dilation), c.groups,
                                     self.feeds_masking_pool and not self.tapped)

    def forward_with_style_tap(self, x, target_gram, precision=None, acc=None, loss_weight=1.0):
        """(activation, style-loss term of this layer against `target_gram` (1|B,C,C)); CUDA only.
        `target_gram` must be SYMMETRIC -- a Gram matrix or a blend of Gram matrices, what losses.py:24 builds: the
        backward then reads dG = 2 scale (G - target) as the symmetric matrix it is and skips the dG + dG^T pass
        (st3d.functional.style_layer_loss is the entry point for arbitrary targets).
        acc: a (1,) float32 loss accumulator; the kernel adds `loss_weight` x term to it in place and it is returned in
        place of the term (see perceptual_loss_of_images)."""
        c = self.conv
        bias = c.bias if c.bias is not None else torch.zeros(c.out_channels, device=x.device, dtype=x.dtype)
        return _ConvReLUStyleTapFn.apply(x, c.weight, bias, target_gram, tuple(c.stride), tuple(c.padding),
                                         tuple(c.dilation), c.groups, precision, acc, loss_weight)


def _content_tap(self, x, content_feat, acc=None, loss_weight=1.0):
    """(activation, mean((activation - content_feat)^2)) with the fused one-kernel backward; CUDA only.  acc / loss_weight as
    in forward_with_style_tap."""
    c = self.conv
    bias = c.bias if c.bias is not None else torch.zeros(c.out_channels, device=x.device, dtype=x.dtype)
    return _ConvReLUContentTapFn.apply(x, c.weight, bias, content_feat, tuple(c.stride), tuple(c.padding),
                                       tuple(c.dilation), c.groups, acc, loss_weight)


FusedConvReLU.forward_with_content_tap = _content_tap


def fuse_vgg_features(features: nn.Sequential, channels_last: bool = True, fuse_pool: bool = True) -> nn.Sequential:
    """Same module names as `features`; every Conv2d followed by a ReLU becomes one FusedConvReLU + Identity, every
    MaxPool2d(2, 2) a FusedMaxPool (libst3d kernels) that also carries the ReLU mask of a FusedConvReLU before it."""
    items = list(features._modules.items())
    out = OrderedDict()
    skip = False
    last_fused = None       # the FusedConvReLU whose activation is the current tensor, if any
    for i, (name, m) in enumerate(items):
        if skip:
            out[name] = nn.Identity()
            skip = False
            continue
        nxt = items[i + 1][1] if i + 1 < len(items) else None
        if isinstance(m, nn.Conv2d) and isinstance(nxt, nn.ReLU):
            out[name] = last_fused = FusedConvReLU(m)
            skip = True
            continue
        if fuse_pool and channels_last and FusedMaxPool.accepts(m):
            out[name] = FusedMaxPool(m, after_relu=last_fused is not None)
            if last_fused is not None:
                last_fused.feeds_masking_pool = True
        else:
            out[name] = m
        last_fused = None
    fused = nn.Sequential(out).eval()
    if channels_last:
        fused = fused.to(memory_format=torch.channels_last)
    fused._st3d_channels_last = bool(channels_last)     # get_features converts its input accordingly
    return fused

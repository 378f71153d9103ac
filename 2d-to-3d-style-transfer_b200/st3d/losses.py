"""Fused-kernel versions of the reference's loss bodies (losses.py:12-44, 68-98; style_transfer.py:10-35).

The VGG-19 convolutions stay on torch/cuDNN (out of scope per BASELINE.json); everything after the
feature taps -- Gram products, the MSE against the style Grams, the content MSE, the masked image MSE
and all their backwards -- runs in libst3d kernels."""
from __future__ import annotations

import torch

from . import functional as Fn

# VGG-19 `.features` taps of style_transfer.py:12-19.  The ReLUs are in-place, so every tapped tensor is
# the post-ReLU activation by the time it is used.
VGG_TAPS = {"0": "conv1_1", "5": "conv2_1", "10": "conv3_1", "19": "conv4_1", "21": "conv4_2", "28": "conv5_1"}
CONTENT_LAYER = "conv4_2"
LAST_TAP = 28


def get_features(image: torch.Tensor, model, layers=None, stop_after_last_tap: bool = True):
    """style_transfer.py:10-27.  Walks `model` (VGG-19 `.features`) and returns {layer name: activation}.
    The reference keeps walking modules 30-36 whose outputs nobody reads; this stops after the last tap
    and the in-place ReLU that follows it (that ReLU rewrites the tapped tensor, so it is part of the tap)."""
    layers = VGG_TAPS if layers is None else layers
    last = max((int(k) for k in layers if str(k).isdigit()), default=None)
    if getattr(model, "_st3d_channels_last", False) and image.dim() == 4 and image.is_cuda:
        image = image.contiguous(memory_format=torch.channels_last)
    feats, x, done = {}, image, False
    for name, module in model._modules.items():
        if done and not _rewrites_the_tap(module):
            break
        if hasattr(module, "tapped"):       # FusedConvReLU: a tapped activation has a second consumer (st3d/vgg.py)
            module.tapped = name in layers
            x = module(x)
            module.tapped = True
        else:
            x = module(x)
        if done:
            break
        if name in layers:
            feats[layers[name]] = x
        done = stop_after_last_tap and last is not None and name == str(last)
    return feats


def _rewrites_the_tap(module) -> bool:
    """Modules that leave the tapped tensor's identity intact: the in-place ReLU behind a tapped convolution
    (it overwrites the stored tensor, style_transfer.py:21-26) and the Identity a fused model keeps in its place."""
    return (isinstance(module, torch.nn.ReLU) and module.inplace) or isinstance(module, torch.nn.Identity)


def style_targets(style_imgs: torch.Tensor, model, precision=None):
    """Gram matrices of the style image's features (losses.py:19-25), conv4_2 excluded."""
    with torch.no_grad():
        feats = get_features(style_imgs, model)
        return {k: Fn.gram_matrix(v, precision) for k, v in feats.items() if k != CONTENT_LAYER}


def blended_style_targets(style_imgs, weights, model, precision=None):
    """Multi-style targets (BASELINE configs[3]): Gs = sum_j w_j * Gram(style_j) per layer, as (1,C,C) tensors.
    `style_imgs` is (J,3,H,W); the reference itself uses a single style image (second_approach.py:157)."""
    w = torch.as_tensor(weights, dtype=torch.float32, device=style_imgs.device).reshape(-1, 1, 1)
    if w.shape[0] != style_imgs.shape[0]:
        raise ValueError("one weight per style image")
    grams = style_targets(style_imgs, model, precision)              # layer -> (J,C,C)
    return {k: (g * w).sum(dim=0, keepdim=True) for k, g in grams.items()}


def content_and_style_constants(content_imgs, style_imgs, model, precision=None, style_weights=None):
    """The two constant branches of losses.py:18-25 in ONE walk of the VGG: the content images and the style image(s)
    share a batch up to conv4_2 (a one-image pass leaves most of the GPU idle), the style image(s) continue alone to
    conv5_1.  Returns (content feature conv4_2 of the content images, {layer: style Gram (J|1,C,C)}).
    style_weights: blend the J style Grams into one target per layer (BASELINE configs[3])."""
    if content_imgs.shape[1:] != style_imgs.shape[1:]:
        with torch.no_grad():
            content = get_features(content_imgs, model, {"21": CONTENT_LAYER})[CONTENT_LAYER]
        grams = (blended_style_targets(style_imgs, style_weights, model, precision) if style_weights is not None
                 else style_targets(style_imgs, model, precision))
        return content, grams
    B = content_imgs.shape[0]
    with torch.no_grad():
        if getattr(model, "_st3d_channels_last", False) and content_imgs.is_cuda:
            # both channels_last BEFORE the cat (the renderer can deliver its views that way): the cat then writes
            # channels_last storage directly instead of an NCHW batch that needs a second, full-size layout copy
            content_imgs = content_imgs.contiguous(memory_format=torch.channels_last)
            style_imgs = style_imgs.contiguous(memory_format=torch.channels_last)
        x = torch.cat([content_imgs, style_imgs], dim=0)
        if getattr(model, "_st3d_channels_last", False) and x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        state = {"content": None, "grams": {}}

        def take(layer, x):
            """Records the tap `layer` from the activation x and returns what continues down the network."""
            if layer == CONTENT_LAYER:
                state["content"] = x[:B]
                return x[B:]                    # only the style image(s) go on to conv5_1
            state["grams"][layer] = Fn.gram_matrix(x[B:] if state["content"] is None else x, precision)
            return x

        modules = list(model._modules.items())
        pending, done = None, False
        for i, (name, module) in enumerate(modules):
            if done and not _rewrites_the_tap(module):
                break
            x = module(x)
            if pending is not None:             # the in-place ReLU behind a tapped Conv2d has run: NOW it is the tap
                x, pending = take(pending, x), None
            if done:
                break
            layer = VGG_TAPS.get(name)
            if layer is not None:
                # style_transfer.py:21-26 stores the tensor a tapped module returns, and the in-place ReLU that
                # follows rewrites that very tensor: every tap is the POST-ReLU activation (SURVEY section 8 row
                # a10).  A FusedConvReLU already returns it; behind a plain Conv2d the tap waits for the ReLU.
                nxt = modules[i + 1][1] if i + 1 < len(modules) else None
                if nxt is not None and _rewrites_the_tap(nxt) and not isinstance(nxt, torch.nn.Identity):
                    pending = layer
                else:
                    x = take(layer, x)
            done = name == str(LAST_TAP)
        content, grams = state["content"], state["grams"]
        if style_weights is not None:
            w = torch.as_tensor(style_weights, dtype=torch.float32, device=style_imgs.device).reshape(-1, 1, 1)
            if w.shape[0] != style_imgs.shape[0]:
                raise ValueError("one weight per style image")
            grams = {k: (g * w).sum(dim=0, keepdim=True) for k, g in grams.items()}
    return content, grams


def perceptual_loss_from_features(cur_feats, content_feat, style_grams, style_weight=1e6, content_weight=1.0,
                                  precision=None):
    """losses.py:28-42 given the three sets of features."""
    content_loss = Fn.mse_loss(cur_feats[CONTENT_LAYER], content_feat)
    style_loss = None
    for layer, target in style_grams.items():
        term = Fn.style_layer_loss(cur_feats[layer], target, 1.0, precision)
        style_loss = term if style_loss is None else style_loss + term
    return content_weight * content_loss + style_weight * style_loss


def perceptual_loss_of_images(current_imgs, model, content_feat, style_grams, style_weight=1e6, content_weight=1.0,
                              precision=None, loss_scale: float = 1.0):
    """losses.py:26-42 from the images on: walks `model` like get_features and evaluates the loss on the way.
    With a fused model (st3d.vgg.fuse_vgg_features) every tap is evaluated INSIDE its conv + ReLU layer, so the layer's
    backward adds the tap's gradient to the gradient arriving from deeper layers and applies the ReLU mask in one kernel,
    and every tap ADDS its weighted term to ONE loss accumulator from inside its kernel (content_weight x content +
    style_weight x sum of the layer terms, losses.py:42, times `loss_scale`): no per-layer loss tensors, no tree of scalar
    multiply / add kernels in the forward or the backward.  Values and gradients equal get_features +
    perceptual_loss_from_features."""
    from .vgg import FusedConvReLU
    fused_taps = current_imgs.is_cuda and all(isinstance(model._modules.get(n), FusedConvReLU)
                                              for n, l in VGG_TAPS.items() if l in style_grams)
    if not fused_taps:
        loss = perceptual_loss_from_features(get_features(current_imgs, model), content_feat, style_grams, style_weight,
                                             content_weight, precision)
        return loss if loss_scale == 1.0 else loss_scale * loss
    x = current_imgs
    if getattr(model, "_st3d_channels_last", False) and x.dim() == 4:
        x = x.contiguous(memory_format=torch.channels_last)
    acc = torch.zeros(1, device=x.device, dtype=torch.float32)
    sw, cw = float(style_weight) * float(loss_scale), float(content_weight) * float(loss_scale)
    done = False
    for name, module in model._modules.items():
        if done and not _rewrites_the_tap(module):
            break
        layer = VGG_TAPS.get(name)
        if layer in style_grams:
            x, acc = module.forward_with_style_tap(x, style_grams[layer], precision, acc=acc, loss_weight=sw)
        elif layer == CONTENT_LAYER and isinstance(module, FusedConvReLU) and x.shape[0] == content_feat.shape[0]:
            # the content tap inside its conv + ReLU layer: MSE backward + gradient accumulation + ReLU mask in one kernel
            x, acc = module.forward_with_content_tap(x, content_feat, acc=acc, loss_weight=cw)
        else:
            if hasattr(module, "tapped"):
                module.tapped = layer is not None
                x = module(x)
                module.tapped = True
            else:
                x = module(x)
            if layer == CONTENT_LAYER:
                acc = acc + cw * Fn.mse_loss(x, content_feat)
        if done:
            break
        done = name == str(LAST_TAP)
    return acc.reshape(())


def compute_perceptual_loss(current_imgs, content_imgs, style_imgs, model, style_weight=1e6, content_weight=1,
                            precision=None):
    """losses.py:12-44.  `style_imgs` may have batch 1 (the reference repeats one image B times,
    second_approach.py:157; a single copy gives the same target Gram for every view)."""
    B = current_imgs.shape[0]
    assert content_imgs.shape[0] == B and style_imgs.shape[0] in (1, B)
    with torch.no_grad():
        content_feat = get_features(content_imgs, model, {"21": CONTENT_LAYER})[CONTENT_LAYER]
    grams = style_targets(style_imgs, model, precision)
    return perceptual_loss_of_images(current_imgs, model, content_feat, grams, style_weight, content_weight, precision)


def compute_first_approach_image_loss(rendered, masks, target_rendered):
    """losses.py:71-75 (`texture` target): mse_loss(rendered * masks, target * masks)."""
    return Fn.masked_mse_loss(rendered, target_rendered, masks)

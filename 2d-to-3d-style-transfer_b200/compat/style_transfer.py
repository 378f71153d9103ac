"""Drop-in for the reference's style_transfer.py (same names, same arguments), with the Gram products,
the style MSE and the content MSE running in libst3d kernels.  VGG-19 stays on torch/cuDNN.

    get_features    style_transfer.py:10-27
    gram_matrix     style_transfer.py:31-35
    style_transfer  style_transfer.py:38-84   (2D neural style transfer of a batch of images)
"""
import os

import torch
from tqdm import tqdm

from st3d import functional as _fn
from st3d import losses as _losses

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
_GRAPH_WARMUP = 3


def get_features(image, model, layers=None):
    return _losses.get_features(image, model, layers)


def gram_matrix(tensor):
    return _fn.gram_matrix(tensor)


def style_transfer(initial_optimized_imgs, content_imgs, style_imgs, model, steps=2000, style_weight=1e6,
                   content_weight=1, lr=0.003):
    assert initial_optimized_imgs.shape[0] == content_imgs.shape[0] == style_imgs.shape[0]
    # constants of the loop: content features and style Grams (one Gram per image, as in the reference)
    with torch.no_grad():
        content_feat = _losses.get_features(content_imgs, model, {"21": _losses.CONTENT_LAYER})[_losses.CONTENT_LAYER]
    grams = _losses.style_targets(style_imgs, model)
    images = initial_optimized_imgs.clone().detach().to(device).requires_grad_(True)
    graphed = images.is_cuda and steps > _GRAPH_WARMUP and os.environ.get("ST3D_NST_GRAPH", "1") != "0"
    # capturable: Adam's step counter lives on the device, so the update can be replayed from a CUDA graph
    optimizer = torch.optim.Adam([images], lr=lr, capturable=True, fused=True) if graphed else torch.optim.Adam([images], lr=lr)

    def iteration():
        # features + loss in one walk: style taps are evaluated inside their conv layers on a fused model (st3d.vgg)
        loss = _losses.perceptual_loss_of_images(images, model, content_feat, grams, style_weight, content_weight)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()

    if not graphed:
        for _ in tqdm(range(steps), desc="2D Style Transfer"):
            iteration()
        return images
    # style_transfer.py:59-83 as ONE CUDA graph per step: the first steps run eagerly (they are
    # the warm-up of the capture AND real iterations), every later one is a replay
    from st3d.optimize import CapturedIteration
    step = CapturedIteration(iteration, images.device, warmup=_GRAPH_WARMUP)
    for _ in tqdm(range(steps - _GRAPH_WARMUP), desc="2D Style Transfer", initial=_GRAPH_WARMUP, total=steps):
        step.replay()
    return images

"""The style-optimisation iteration itself -- the body of the reference's hot loop
(second_approach.py:147-190 for the `texture` target), batched over views and running on libst3d
kernels, with view-sharded data parallelism over NCCL (one process per GPU).

One iteration = render the content mesh (no grad) -> render the current mesh -> VGG features of
content / style / current -> content MSE + Gram style loss -> backward into the texture -> (all-reduce
of the texture gradient when sharded) -> Adam step.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import functional as Fn
from . import losses


def allreduce_gradients(params, group=None) -> None:
    """SUM all-reduce of the gradients over the view-sharding ranks as ONE flat fp32 buffer
    (texture S*S*3 [+ verts V*3]); NCCL over NVLink when the process group is NCCL."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class TextureStyleOptimizer:
    """Optimises the UV texture of a mesh so that its renders match a style image (Approach 2)."""

    def __init__(self, verts, faces, verts_uvs, faces_uvs, texture, vgg, image_size, lr: float = 0.01,
                 style_weight: float = 1e6, content_weight: float = 1.0, precision=None,
                 cache_constants: bool = False, world_size: int = 1, group=None, channels_last: bool = True,
                 fuse_conv_relu: bool = True):
        dev = verts.device
        if dev.type != "cuda":
            raise RuntimeError("TextureStyleOptimizer needs CUDA tensors: libst3d has no CPU path")
        self.verts = verts.detach().float().contiguous()
        self.faces = faces.to(torch.int32).contiguous()
        self.face_uvs = verts_uvs.float()[faces_uvs.long()].contiguous()          # (F,3,2)
        self.content_texture = texture.detach().clone().float().contiguous()     # the mesh as loaded
        self.texture = texture.detach().clone().float().contiguous().requires_grad_(True)
        # cuDNN's tensor-core convolutions are NHWC kernels: with NCHW tensors torch wraps every conv in
        # nchw<->nhwc transposes (~20 % of a step).  channels_last keeps activations NHWC end to end.
        self.channels_last = channels_last
        if fuse_conv_relu:          # cuDNN's fused conv + bias + ReLU entry point (same numbers, fewer kernels)
            from .vgg import fuse_vgg_features
            self.vgg = fuse_vgg_features(vgg, channels_last)
        else:
            self.vgg = vgg.to(memory_format=torch.channels_last) if channels_last else vgg
        self.image_size = image_size
        self.style_weight, self.content_weight = style_weight, content_weight
        self.precision = precision
        self.cache_constants = cache_constants
        self.world_size, self.group = world_size, group
        self.optimizer = torch.optim.Adam([self.texture], lr=lr)
        self._cache = {}
        self.last_images: Optional[torch.Tensor] = None

    def _nn_input(self, images):
        return images.contiguous(memory_format=torch.channels_last) if self.channels_last else images

    def _render(self, texture, R, T):
        images, masks, _ = Fn.render_views(self.verts, self.faces, R, T, self.image_size, texture=texture,
                                           face_uvs=self.face_uvs)
        return images, masks

    def step(self, R: torch.Tensor, T: torch.Tensor, style_img: torch.Tensor) -> torch.Tensor:
        """One optimisation iteration over the views (R, T) held by this rank; returns the loss (device scalar)."""
        self.optimizer.zero_grad(set_to_none=True)
        key = (R.data_ptr(), T.data_ptr(), style_img.data_ptr())
        if self.cache_constants and self._cache.get("key") == key:
            content_feat, grams = self._cache["content_feat"], self._cache["grams"]
        else:
            with torch.no_grad():
                content_imgs, _ = self._render(self.content_texture, R, T)                     # second_approach.py:160
                content_feat = losses.get_features(self._nn_input(content_imgs), self.vgg,
                                                   {"21": losses.CONTENT_LAYER})[losses.CONTENT_LAYER]
            grams = losses.style_targets(self._nn_input(style_img), self.vgg, self.precision)                   # losses.py:19-25
            if self.cache_constants:
                self._cache = dict(key=key, content_feat=content_feat, grams=grams)
        current_imgs, _ = self._render(self.texture, R, T)                                      # :165
        cur = losses.get_features(self._nn_input(current_imgs), self.vgg)
        loss = losses.perceptual_loss_from_features(cur, content_feat, grams, self.style_weight, self.content_weight,
                                                    self.precision)
        (loss / self.world_size).backward()                                                     # :188
        allreduce_gradients([self.texture], self.group)
        self.optimizer.step()                                                                   # :189
        self.last_images = current_imgs.detach()
        return loss.detach()

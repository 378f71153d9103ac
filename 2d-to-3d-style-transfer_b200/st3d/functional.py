"""Differentiable (torch.autograd) front end of the libst3d kernels.

Forward and backward of every function below run hand-written sm_100a kernels through the C ABI
(st3d.ops); torch supplies device memory, streams and the autograd tape only.  There is no CPU path.

Mirrors, batched over views, what the reference reaches through `render_meshes` (utils.py:65-77),
`gram_matrix` (style_transfer.py:31-35) and the loss bodies of losses.py:31-39 / 71-75.
"""
from __future__ import annotations

import functools
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


@functools.lru_cache(maxsize=64)
def fov_scales(fov_deg: float = 60.0, aspect: float = 1.0, znear: float = 1.0):
    """K00, K11 of FoVPerspectiveCameras' projection, rounded to fp32 step by step so the
    projected vertices are bit-identical to the oracle's."""
    f32 = np.float32
    tan_half = f32(math.tan(f32(fov_deg) * f32(math.pi / 180.0) / 2.0))
    max_y = f32(tan_half * f32(znear))
    max_x = f32(max_y * f32(aspect))
    k00 = f32(f32(2.0) * f32(znear) / f32(max_x - (-max_x)))
    k11 = f32(f32(2.0) * f32(znear) / f32(max_y - (-max_y)))
    return float(k00), float(k11)


# ------------------------------------------------------------------------------------------------
# fused multi-view renderer
# ------------------------------------------------------------------------------------------------
class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, tex_param, faces, R, T, face_uvs, spec, mode, background_image=None):
        kw = dict(face_uvs=face_uvs, texture=tex_param) if mode == ops.TEX_UV else dict(verts_rgb=tex_param)
        if background_image is not None:
            kw["background_image"] = background_image
        image, mask, p2f, state = ops.render_forward(spec, verts.detach(), faces, R, T,
                                                     **{k: (v.detach() if isinstance(v, torch.Tensor) else v)
                                                        for k, v in kw.items()})
        ctx.state, ctx.mode = state, mode
        ctx.tex_shape = tuple(tex_param.shape)
        ctx.mark_non_differentiable(p2f)
        if mask is not None:
            ctx.mark_non_differentiable(mask)
            return image, mask, p2f
        return image, p2f

    @staticmethod
    def backward(ctx, grad_image, *unused):
        need_verts, need_tex = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        uv = ctx.mode == ops.TEX_UV
        g_tex, g_verts, g_rgb = ops.render_backward(ctx.state, grad_image, need_texture=need_tex and uv,
                                                    need_verts=need_verts, need_verts_rgb=need_tex and not uv)
        g_param = g_tex if uv else g_rgb
        if g_param is not None:
            g_param = g_param.reshape(ctx.tex_shape)
        return g_verts, g_param, None, None, None, None, None, None, None


def render_views(verts: torch.Tensor, faces: torch.Tensor, R: torch.Tensor, T: torch.Tensor, image_size,
                 *, texture: Optional[torch.Tensor] = None, face_uvs: Optional[torch.Tensor] = None,
                 verts_rgb: Optional[torch.Tensor] = None, fov: float = 60.0, aspect: float = 1.0,
                 znear: float = 1.0, zfar: float = 100.0, blur_radius: float = 0.0, cull_backfaces: bool = False,
                 ambient: Sequence[float] = (1.0, 1.0, 1.0), background: Sequence[float] = (1.0, 1.0, 1.0),
                 sigma: float = 1e-4, gamma: float = 1e-4, planar: bool = True, z_clip: Optional[float] = None,
                 lights: Optional[dict] = None, background_image: Optional[torch.Tensor] = None,
                 cull_to_frustum: bool = False, channels_last: bool = False):
    """Render N camera views of one mesh in one launch sequence (faces_per_pixel = 1).

    lights: None = AmbientLights with the colour `ambient` (what the reference uses); or a dict(kind='point' |
            'directional', location= | direction=, diffuse=, specular= (light colour x material colour), shininess=)
            for phong_shading in the same epilogue -- differentiable w.r.t. the texture / vertex colours, not `verts`.
    background_image: (1|N,3,H,W); uncovered pixels take its colour (apply_background of utils.py:19-30, fused).

    planar=True  -> (images (N,3,H,W), masks (N,1,H,W), pix_to_face (N,H,W) int32): what
                    `render_meshes` (utils.py:65-77) returns, without the per-view Python loop.
                    channels_last=True returns the same (N,3,H,W) images in torch channels_last storage (N,H,W,3) -- what
                    a channels_last VGG takes without a layout copy -- and reads the image gradient in that storage.
    planar=False -> (rgba (N,H,W,4), pix_to_face): what `renderer(meshes_world=, cameras=)` returns.
    Differentiable w.r.t. `verts` and `texture` (UV mode, any shape ending in (Ht,Wt,3)) or `verts_rgb`.
    """
    H, W = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
    k00, k11 = fov_scales(fov, aspect, znear)
    spec = ops.RenderSpec(image_size=(H, W), k00=k00, k11=k11, znear=znear, zfar=zfar, blur_radius=blur_radius,
                          cull_backfaces=cull_backfaces, cull_to_frustum=cull_to_frustum, ambient=tuple(ambient),
                          background=tuple(background), sigma=sigma, gamma=gamma, z_clip=z_clip,
                          layout=(ops.LAYOUT_NHWC_RGB if channels_last else ops.LAYOUT_PLANAR) if planar
                          else ops.LAYOUT_NHWC_RGBA)
    if lights is not None and lights.get("kind", "ambient") != "ambient":
        kind = lights["kind"]
        if kind not in ("point", "directional"):
            raise ValueError(f"unknown light kind {kind!r}")
        spec.light_kind = ops.LIGHT_POINT if kind == "point" else ops.LIGHT_DIRECTIONAL
        spec.light_vec = tuple(float(v) for v in lights["location" if kind == "point" else "direction"])
        spec.light_diffuse = tuple(float(v) for v in lights.get("diffuse", (0.0, 0.0, 0.0)))
        spec.light_specular = tuple(float(v) for v in lights.get("specular", (0.0, 0.0, 0.0)))
        spec.shininess = float(lights.get("shininess", 64.0))
    if background_image is not None:
        background_image = background_image.detach()
    if texture is not None:
        if face_uvs is None:
            raise ValueError("texture needs face_uvs (F,3,2) = verts_uvs[faces_uvs]")
        return _RenderFn.apply(verts, texture, faces, R, T, face_uvs, spec, ops.TEX_UV, background_image)
    if verts_rgb is None:
        raise ValueError("render_views needs either texture+face_uvs or verts_rgb")
    return _RenderFn.apply(verts, verts_rgb, faces, R, T, None, spec, ops.TEX_VERTEX, background_image)


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, mask, fill):
        ctx.save_for_backward(mask)
        return ops.composite_forward(image.detach(), mask, fill.detach())

    @staticmethod
    def backward(ctx, grad_out):
        (mask,) = ctx.saved_tensors
        return ops.composite_backward(grad_out.contiguous(), mask), None, None


def composite_background(images: torch.Tensor, masks: torch.Tensor, fill: torch.Tensor) -> torch.Tensor:
    """utils.py:19-30 (`apply_background`, 'noise' / 'style'): images * masks + fill * (1 - masks) as one kernel;
    differentiable w.r.t. `images` (the fill is a constant of the step)."""
    return _CompositeFn.apply(images, masks, fill)


# ------------------------------------------------------------------------------------------------
# operator boundary with autograd (what pytorch3d's _RasterizeFaceVerts / interpolate_face_attributes are)
# ------------------------------------------------------------------------------------------------
class _RasterizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, face_verts, first, num, image_size, blur_radius, K, persp, clip, cull, neighbor=None):
        p2f, zbuf, bary, dists = ops.rasterize_meshes(face_verts.detach(), first, num, image_size, blur_radius, K, 0, 0,
                                                      persp, clip, cull, clipped_faces_neighbor_idx=neighbor)
        ctx.save_for_backward(face_verts.detach(), p2f)
        ctx.flags = (persp, clip)
        ctx.mark_non_differentiable(p2f)
        return p2f, zbuf, bary, dists

    @staticmethod
    def backward(ctx, _, g_zbuf, g_bary, g_dists):
        face_verts, p2f = ctx.saved_tensors
        g = ops.rasterize_meshes_backward(face_verts, p2f, g_zbuf.contiguous(), g_bary.contiguous(),
                                          g_dists.contiguous(), *ctx.flags)
        return g, None, None, None, None, None, None, None, None, None


def rasterize_meshes(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, blur_radius=0.0,
                     faces_per_pixel=1, perspective_correct=True, clip_barycentric_coords=False, cull_backfaces=False,
                     z_clip_value: Optional[float] = None, cull_to_frustum: bool = False):
    """Upstream `rasterize_meshes` from packed face vertices on: optional near-plane clipping (torch ops, as
    upstream's clip.py) -> `_C.rasterize_meshes` (libst3d kernels) -> indices / barycentrics mapped back to the
    unclipped faces.  Differentiable w.r.t. face_verts."""
    args = (image_size, float(blur_radius), int(faces_per_pixel), bool(perspective_correct),
            bool(clip_barycentric_coords), bool(cull_backfaces))
    if z_clip_value is None and not cull_to_frustum:
        return _RasterizeFn.apply(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, *args)
    from . import clip as _clip
    cl = _clip.clip_faces(face_verts, mesh_to_face_first_idx, num_faces_per_mesh,
                          None if z_clip_value is None else float(z_clip_value), bool(perspective_correct),
                          bool(cull_to_frustum))
    p2f, zbuf, bary, dists = _RasterizeFn.apply(cl.face_verts, cl.mesh_to_face_first_idx, cl.num_faces_per_mesh, *args,
                                                cl.clipped_faces_neighbor_idx)
    p2f, bary = _clip.convert_clipped_rasterization_to_original_faces(p2f, bary, cl)
    return p2f, zbuf, bary, dists


class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pix_to_face, bary, face_attrs):
        ctx.save_for_backward(pix_to_face, bary.detach(), face_attrs.detach())
        return ops.interp_face_attrs_forward(pix_to_face, bary.detach(), face_attrs.detach())

    @staticmethod
    def backward(ctx, grad_out):
        p2f, bary, attrs = ctx.saved_tensors
        g_bary, g_attrs = ops.interp_face_attrs_backward(p2f, bary, attrs, grad_out.contiguous())
        return None, g_bary.reshape(bary.shape), g_attrs


def interpolate_face_attributes(pix_to_face, bary, face_attrs):
    """(N,H,W,K), (N,H,W,K,3), (F,3,D) -> (N,H,W,K,D)."""
    out = _InterpFn.apply(pix_to_face.reshape(-1), bary.reshape(-1, 3), face_attrs)
    return out.reshape(*pix_to_face.shape, face_attrs.shape[-1])


# ------------------------------------------------------------------------------------------------
# Gram / style / MSE losses
# ------------------------------------------------------------------------------------------------
class _GramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, precision):
        ctx.save_for_backward(feat.detach())
        ctx.precision = precision
        return ops.gram_forward(feat.detach(), precision)

    @staticmethod
    def backward(ctx, dgram):
        (feat,) = ctx.saved_tensors
        return ops.gram_backward(feat, dgram.contiguous(), 1.0, precision=ctx.precision), None


def gram_matrix(tensor: torch.Tensor, precision=None) -> torch.Tensor:
    """style_transfer.py:31-35: (B,C,H,W) -> (B,C,C) = F F^T (no normalisation)."""
    if tensor.dim() != 4:
        raise ValueError("gram_matrix expects (B,C,H,W)")
    return _GramFn.apply(tensor, precision)


class _StyleLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, target, weight, precision):
        B, C, H, W = feat.shape
        scale = float(weight) / (B * C * C) / (float(C) ** 2 * float(H) ** 2)
        loss = torch.zeros(1, device=feat.device, dtype=torch.float32)
        dgram, _ = ops.gram_mse_forward(feat.detach(), target.detach(), scale, loss, precision=precision)
        ctx.save_for_backward(feat.detach(), dgram)
        ctx.precision = precision
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        feat, dgram = ctx.saved_tensors
        # dgram = dL/dG for grad_out == 1; the upstream scalar is applied on the device without a host sync
        # (general in the target: an arbitrary (C,C) target gives a non-symmetric dG, so the dG + dG^T pass stays; the
        # fused VGG taps, whose targets are Gram matrices by construction, skip it -- st3d/vgg.py)
        g = ops.gram_backward(feat, dgram, 1.0, precision=ctx.precision, scale_tensor=grad_out)
        return g, None, None, None


def style_layer_loss(feat: torch.Tensor, target_gram: torch.Tensor, weight: float = 1.0, precision=None):
    """losses.py:35-39 for one layer, fused: weight * mean((gram(feat) - target)^2) / (C^2 * H^2).
    target_gram (1|B,C,C) is treated as a constant (it comes from the style image)."""
    if feat.dim() != 4:
        raise ValueError("style_layer_loss expects (B,C,H,W) features")
    return _StyleLayerFn.apply(feat, target_gram, weight, precision)


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mask):
        loss = torch.zeros(1, device=a.device, dtype=torch.float32)
        grad = ops.mse_forward(a.detach(), b.detach(), 1.0 / max(a.numel(), 1), loss, mask=mask,
                               want_grad=ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        ga = grad * grad_out if ctx.needs_input_grad[0] else None
        gb = -(grad * grad_out) if ctx.needs_input_grad[1] else None
        return ga, gb, None


def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean((a - b)^2) over all elements (losses.py:31 content loss)."""
    if a.shape != b.shape:
        raise ValueError(f"mse_loss: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    return _MseFn.apply(a, b, None)


def masked_mse_loss(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(a * mask, b * mask) of losses.py:71-75: the mean runs over ALL elements."""
    return _MseFn.apply(a, b, mask)


# ------------------------------------------------------------------------------------------------
# camera transform with autograd (MeshRasterizer.transform of the unfused, Fragments-returning path)
# ------------------------------------------------------------------------------------------------
class _TransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, R, T, k00, k11):
        ctx.save_for_backward(verts.detach(), R, T)
        ctx.k = (k00, k11)
        return ops.transform_verts(verts.detach(), R, T, k00, k11)

    @staticmethod
    def backward(ctx, grad_ndc):
        verts, R, T = ctx.saved_tensors
        return ops.transform_verts_backward(verts, R, T, ctx.k[0], ctx.k[1], grad_ndc.contiguous()), None, None, None, None


def transform_verts(verts, R, T, fov: float = 60.0, aspect: float = 1.0, znear: float = 1.0):
    """(V,3) world -> (N,V,3) [x_ndc, y_ndc, z_view] for N FoV-perspective cameras (row-vector R, T)."""
    k00, k11 = fov_scales(fov, aspect, znear)
    return _TransformFn.apply(verts, R.reshape(-1, 3, 3), T.reshape(-1, 3), k00, k11)

"""Renderer surface used by the reference (first_approach.py:106-114, utils.py:69, 149, 168, 208):
cameras, rasterization settings, rasterizer, shader, renderer, textures.

`MeshRenderer(...)(meshes_world=, cameras=)` runs the FUSED libst3d path (vertex transform -> tile bins
-> fine raster -> texture sample -> ambient shade -> soft blend in one launch sequence) for every camera
passed, whether that is one camera (the reference's per-view loop, utils.py:68-69) or a whole batch.
`MeshRasterizer(...)(meshes)` alone returns Fragments through the operator-boundary kernels; Point /
Directional lights and faces_per_pixel > 1 take the general path (Fragments + shading.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional, Sequence, Union

import torch

from st3d import cameras as _cam
from st3d import functional as _fn

from ..structures import Meshes
from .cameras import FoVPerspectiveCameras, look_at_view_transform  # noqa: F401  (re-exported)


# ------------------------------------------------------------------------------------------------
# textures
# ------------------------------------------------------------------------------------------------
def _stack(x, name):
    if isinstance(x, (list, tuple)):
        if len(x) == 1:
            return x[0][None]           # a view: a leaf passed in a one-element list stays reachable
        if any(t.shape != x[0].shape for t in x):
            raise NotImplementedError(f"{name}: the meshes of a batch must share their texture shapes")
        return torch.stack(list(x), dim=0)
    return x


class TexturesUV:
    """maps (N,H,W,C), faces_uvs (N,F,3) int64, verts_uvs (N,V,2); bilinear, border padding,
    align_corners=True, v flipped.  `maps_padded()` returns the tensor that was passed in,
    so a leaf texture keeps receiving gradients (utils.py:177-185)."""

    def __init__(self, maps, faces_uvs, verts_uvs, padding_mode: str = "border", align_corners: bool = True,
                 sampling_mode: str = "bilinear"):
        if (padding_mode, align_corners, sampling_mode) != ("border", True, "bilinear"):
            raise NotImplementedError("TexturesUV: only bilinear / border / align_corners=True sampling is implemented")
        self._maps = _stack(maps, "maps")
        self._faces_uvs = _stack(faces_uvs, "faces_uvs")
        self._verts_uvs = _stack(verts_uvs, "verts_uvs")
        if self._maps.dim() != 4 or self._faces_uvs.dim() != 3 or self._verts_uvs.dim() != 3:
            raise ValueError("TexturesUV expects maps (N,H,W,C), faces_uvs (N,F,3), verts_uvs (N,V,2)")
        if not (self._maps.shape[0] == self._faces_uvs.shape[0] == self._verts_uvs.shape[0]):
            raise ValueError("TexturesUV: batch sizes differ")
        self.device = self._maps.device

    def maps_padded(self):
        return self._maps

    def faces_uvs_padded(self):
        return self._faces_uvs

    def verts_uvs_padded(self):
        return self._verts_uvs

    def maps_list(self):
        return [m for m in self._maps]

    def faces_verts_uvs(self):
        """(F,3,2) per-face UVs of mesh 0."""
        return self._verts_uvs[0][self._faces_uvs[0].long()]

    def __len__(self):
        return self._maps.shape[0]

    def __getitem__(self, index):
        return TexturesUV(self._maps[index:index + 1], self._faces_uvs[index:index + 1], self._verts_uvs[index:index + 1])

    def clone(self):
        return TexturesUV(self._maps.clone(), self._faces_uvs.clone(), self._verts_uvs.clone())

    def detach(self):
        return TexturesUV(self._maps.detach(), self._faces_uvs.detach(), self._verts_uvs.detach())

    def to(self, device):
        return TexturesUV(self._maps.to(device), self._faces_uvs.to(device), self._verts_uvs.to(device))


class TexturesVertex:
    """verts_features (N,V,C): one colour per vertex, interpolated with the barycentrics."""

    def __init__(self, verts_features):
        self._feats = _stack(verts_features, "verts_features")
        if self._feats.dim() != 3:
            raise ValueError("TexturesVertex expects verts_features (N,V,C)")
        self.device = self._feats.device

    def verts_features_padded(self):
        return self._feats

    def verts_features_packed(self):
        return self._feats[0]

    def __len__(self):
        return self._feats.shape[0]

    def __getitem__(self, index):
        return TexturesVertex(self._feats[index:index + 1])

    def clone(self):
        return TexturesVertex(self._feats.clone())

    def detach(self):
        return TexturesVertex(self._feats.detach())

    def to(self, device):
        return TexturesVertex(self._feats.to(device))


# ------------------------------------------------------------------------------------------------
# settings, lights, materials, blending
# ------------------------------------------------------------------------------------------------
@dataclass
class RasterizationSettings:
    image_size: Union[int, Sequence[int]] = 256
    blur_radius: float = 0.0
    faces_per_pixel: int = 1
    bin_size: Optional[int] = None            # accepted, unused: libst3d bins into 16x16 tiles with exact lists
    max_faces_per_bin: Optional[int] = None   # accepted, unused: no fixed per-bin capacity, no dropped faces
    perspective_correct: Optional[bool] = None
    clip_barycentric_coords: Optional[bool] = None
    cull_backfaces: bool = False
    z_clip_value: Optional[float] = None
    cull_to_frustum: bool = False


@dataclass
class BlendParams:
    sigma: float = 1e-4
    gamma: float = 1e-4
    background_color: Sequence[float] = (1.0, 1.0, 1.0)


def _color(c):
    t = torch.as_tensor(c, dtype=torch.float32).reshape(-1)
    if t.numel() != 3:
        raise ValueError("colours are RGB triples")
    return tuple(float(v) for v in t)


class Materials:
    def __init__(self, ambient_color=((1, 1, 1),), diffuse_color=((1, 1, 1),), specular_color=((1, 1, 1),),
                 shininess=64, device="cpu"):
        self.ambient_color, self.diffuse_color = _color(ambient_color), _color(diffuse_color)
        self.specular_color, self.shininess, self.device = _color(specular_color), float(shininess), device


class AmbientLights:
    """Ambient-only lighting: diffuse = specular = 0, so pixel colour = ambient * texel (A.5)."""

    kind = "ambient"

    def __init__(self, ambient_color=((1.0, 1.0, 1.0),), device="cpu"):
        self.ambient_color, self.device = _color(ambient_color), device

    def to(self, device):
        self.device = device
        return self


class PointLights:
    kind = "point"

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), location=((0, 1, 0),), device="cpu"):
        self.ambient_color, self.diffuse_color = _color(ambient_color), _color(diffuse_color)
        self.specular_color, self.location, self.device = _color(specular_color), _color(location), device

    def to(self, device):
        self.device = device
        return self


class DirectionalLights:
    kind = "directional"

    def __init__(self, ambient_color=((0.5, 0.5, 0.5),), diffuse_color=((0.3, 0.3, 0.3),),
                 specular_color=((0.2, 0.2, 0.2),), direction=((0, 1, 0),), device="cpu"):
        self.ambient_color, self.diffuse_color = _color(ambient_color), _color(diffuse_color)
        self.specular_color, self.direction, self.device = _color(specular_color), _color(direction), device

    def to(self, device):
        self.device = device
        return self


class Fragments(NamedTuple):
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: torch.Tensor
    dists: torch.Tensor


# ------------------------------------------------------------------------------------------------
# rasterizer / shader / renderer
# ------------------------------------------------------------------------------------------------
def _image_hw(size):
    return (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))


def _one_mesh(meshes: Meshes):
    if not isinstance(meshes, Meshes):
        raise ValueError("meshes_world must be a pytorch3d.structures.Meshes")
    if len(meshes) != 1:
        raise NotImplementedError("one mesh per call: libst3d batches camera VIEWS of a single mesh")
    verts, faces = meshes.verts_packed(), meshes.faces_packed()
    if not verts.is_cuda:
        raise RuntimeError("this renderer runs on CUDA only (libst3d has no CPU path): move the mesh to a cuda device")
    return verts, faces


def _per_mesh(meshes: Meshes, cameras):
    """A batch of M > 1 meshes renders mesh i with camera i (or the one camera given), as upstream pairs them; this layer
    batches VIEWS inside one launch sequence, so meshes are walked one launch sequence each."""
    if isinstance(cameras, (list, tuple)):
        cameras = FoVPerspectiveCameras.join(cameras)
    M = len(meshes)
    if cameras is not None and len(cameras) not in (1, M):
        raise ValueError(f"{M} meshes need {M} cameras (or one), got {len(cameras)}")
    for i in range(M):
        yield meshes[i], (cameras if cameras is None or len(cameras) == 1 else cameras[i])


def _camera_params(cameras):
    if cameras is None:
        raise ValueError("cameras must be given to the rasterizer / renderer, at construction or per call")
    if isinstance(cameras, (list, tuple)):
        cameras = FoVPerspectiveCameras.join(cameras)
    if not isinstance(cameras, FoVPerspectiveCameras):
        raise NotImplementedError("only FoVPerspectiveCameras is implemented")
    return cameras, cameras.uniform_intrinsics()


class MeshRasterizer(torch.nn.Module):
    def __init__(self, cameras=None, raster_settings: Optional[RasterizationSettings] = None):
        super().__init__()
        self.cameras = cameras
        self.raster_settings = raster_settings if raster_settings is not None else RasterizationSettings()

    def to(self, device):
        if self.cameras is not None:
            self.cameras = self.cameras.to(device)
        return self

    def _settings(self, kwargs):
        return kwargs.get("raster_settings", self.raster_settings)

    def forward(self, meshes_world, **kwargs) -> Fragments:
        if isinstance(meshes_world, Meshes) and len(meshes_world) > 1:
            parts, offset, out = [], 0, []
            for mesh, cam in _per_mesh(meshes_world, kwargs.get("cameras", self.cameras)):
                fr = self.forward(mesh, **dict(kwargs, cameras=cam))
                # pix_to_face indexes the PACKED faces of the whole batch
                out.append(fr._replace(pix_to_face=torch.where(fr.pix_to_face >= 0, fr.pix_to_face + offset, fr.pix_to_face)))
                offset += mesh.faces_packed().shape[0]
            return Fragments(*(torch.cat([getattr(f, name) for f in out], dim=0) for name in Fragments._fields))
        cams, (fov, aspect, znear, zfar) = _camera_params(kwargs.get("cameras", self.cameras))
        rs = self._settings(kwargs)
        verts, faces = _one_mesh(meshes_world)
        R, T = cams.R.to(verts.device), cams.T.to(verts.device)
        N, Fn = R.shape[0], faces.shape[0]
        ndc = _fn.transform_verts(verts, R, T, fov, aspect, znear)                    # (N,V,3)
        face_verts = ndc[:, faces.long()].reshape(N * Fn, 3, 3)
        first = torch.arange(N, device=verts.device, dtype=torch.int64) * Fn
        num = torch.full((N,), Fn, device=verts.device, dtype=torch.int64)
        clip_bary = rs.clip_barycentric_coords if rs.clip_barycentric_coords is not None else rs.blur_radius > 0.0
        persp = rs.perspective_correct if rs.perspective_correct is not None else True
        # upstream: z_clip = znear / 2 for perspective-correct rasterization unless given
        z_clip = rs.z_clip_value if rs.z_clip_value is not None else (znear / 2.0 if persp else None)
        p2f, zbuf, bary, dists = _fn.rasterize_meshes(face_verts, first, num, _image_hw(rs.image_size), rs.blur_radius,
                                                      rs.faces_per_pixel, persp, clip_bary, rs.cull_backfaces,
                                                      z_clip_value=z_clip, cull_to_frustum=rs.cull_to_frustum)
        return Fragments(p2f, zbuf, bary, dists)


class SoftPhongShader(torch.nn.Module):
    def __init__(self, device="cpu", cameras=None, lights=None, materials=None, blend_params=None):
        super().__init__()
        self.lights = lights if lights is not None else AmbientLights(device=device)
        self.materials = materials if materials is not None else Materials(device=device)
        self.cameras = cameras
        self.blend_params = blend_params if blend_params is not None else BlendParams()

    def to(self, device):
        return self

    def forward(self, fragments, meshes, **kwargs):
        """General path (any lights, any faces_per_pixel) on top of Fragments; see shading.py."""
        from . import shading
        cams, (fov, aspect, znear, zfar) = _camera_params(kwargs.get("cameras", self.cameras))
        lights = kwargs.get("lights", self.lights)
        materials = kwargs.get("materials", self.materials)
        blend = kwargs.get("blend_params", self.blend_params)
        verts, faces = _one_mesh(meshes)
        tex = meshes.textures
        if isinstance(tex, TexturesUV):
            texels = shading.sample_textures_uv(fragments, tex, faces.shape[0])
        elif isinstance(tex, TexturesVertex):
            texels = shading.sample_textures_vertex(fragments, tex, faces, faces.shape[0])
        else:
            raise ValueError("meshes.textures must be TexturesUV or TexturesVertex")
        colors = shading.phong_shading(fragments, verts, faces, texels, lights, materials,
                                       cams.get_camera_center().to(verts.device))
        return shading.softmax_rgb_blend(colors, fragments, blend, znear, zfar)


class MeshRenderer(torch.nn.Module):
    """renderer(meshes_world=mesh, cameras=camera) -> (N,H,W,4) RGBA, N = number of cameras."""

    def __init__(self, rasterizer: MeshRasterizer, shader: SoftPhongShader):
        super().__init__()
        self.rasterizer, self.shader = rasterizer, shader

    def to(self, device):
        self.rasterizer.to(device)
        return self

    def _fused_args(self, meshes_world, kwargs):
        cams, (fov, aspect, znear, zfar) = _camera_params(kwargs.get("cameras", self.rasterizer.cameras))
        rs = kwargs.get("raster_settings", self.rasterizer.raster_settings)
        lights = kwargs.get("lights", self.shader.lights)
        materials = kwargs.get("materials", self.shader.materials)
        blend = kwargs.get("blend_params", self.shader.blend_params)
        if rs.faces_per_pixel != 1:
            raise NotImplementedError("the fused renderer implements faces_per_pixel=1 (what the reference uses); "
                                      "use MeshRasterizer for K > 1 fragments")
        if rs.perspective_correct is False:
            raise NotImplementedError("perspective_correct=False is not implemented in the fused path")
        verts, faces = _one_mesh(meshes_world)
        tex = meshes_world.textures
        ambient = tuple(l * m for l, m in zip(lights.ambient_color, materials.ambient_color))
        phong = None
        if not isinstance(lights, AmbientLights):       # Point / Directional: phong_shading in the same epilogue
            kind = lights.kind
            phong = dict(kind=kind, diffuse=tuple(l * m for l, m in zip(lights.diffuse_color, materials.diffuse_color)),
                         specular=tuple(l * m for l, m in zip(lights.specular_color, materials.specular_color)),
                         shininess=materials.shininess)
            phong["location" if kind == "point" else "direction"] = lights.location if kind == "point" else lights.direction
        common = dict(fov=fov, aspect=aspect, znear=znear, zfar=zfar, blur_radius=rs.blur_radius, lights=phong,
                      cull_to_frustum=rs.cull_to_frustum,
                      cull_backfaces=rs.cull_backfaces, ambient=ambient, background=_color(blend.background_color),
                      sigma=blend.sigma, gamma=blend.gamma,
                      z_clip=rs.z_clip_value if rs.z_clip_value is not None else znear / 2.0)
        R, T = cams.R.to(verts.device).float(), cams.T.to(verts.device).float()
        if isinstance(tex, TexturesUV):
            maps = tex.maps_padded()
            if maps.shape[0] != 1 or maps.shape[-1] != 3:
                raise NotImplementedError("one RGB texture map per mesh")
            tex_kw = dict(texture=maps, face_uvs=tex.faces_verts_uvs())
        elif isinstance(tex, TexturesVertex):
            tex_kw = dict(verts_rgb=tex.verts_features_packed())
        else:
            raise ValueError("meshes_world.textures must be TexturesUV or TexturesVertex")
        return verts, faces, R, T, _image_hw(rs.image_size), tex_kw, common

    def _can_fuse(self, meshes_world, kwargs):
        rs = kwargs.get("raster_settings", self.rasterizer.raster_settings)
        lights = kwargs.get("lights", self.shader.lights)
        # Point / Directional lights are shaded in the fused epilogue too; their backward reaches the texture / vertex
        # colours only, so a mesh whose vertices are being optimised under such lights takes the general path
        lit_ok = isinstance(lights, AmbientLights) or (isinstance(lights, (PointLights, DirectionalLights))
                                                       and not (torch.is_grad_enabled()
                                                                and meshes_world.verts_packed().requires_grad))
        # one face per pixel, hard or soft (blur_radius > 0: tile-bin path, barycentrics clamped as upstream's default
        # clip_barycentric_coords = blur_radius > 0 does); both clip faces at the near plane inside the kernels
        clip_bary = rs.clip_barycentric_coords
        return (lit_ok and rs.faces_per_pixel == 1 and rs.perspective_correct is not False
                and (clip_bary is None or bool(clip_bary) == (rs.blur_radius > 0.0)))

    def forward(self, meshes_world, **kwargs):
        if isinstance(meshes_world, Meshes) and len(meshes_world) > 1:
            cams = kwargs.get("cameras", self.rasterizer.cameras)
            return torch.cat([self.forward(mesh, **dict(kwargs, cameras=cam)) for mesh, cam in _per_mesh(meshes_world, cams)],
                             dim=0)
        if not self._can_fuse(meshes_world, kwargs):       # K > 1, lit geometry gradients: Fragments + general shader
            fragments = self.rasterizer(meshes_world, **kwargs)
            return self.shader(fragments, meshes_world, **kwargs)
        verts, faces, R, T, size, tex_kw, common = self._fused_args(meshes_world, kwargs)
        rgba, _ = _fn.render_views(verts, faces, R, T, size, planar=False, **tex_kw, **common)
        return rgba

    def render_planar(self, meshes_world, **kwargs):
        """Batched fast path used by this repo's `utils.render_meshes`: ((N,3,H,W) images, (N,1,H,W) masks)
        straight from the kernel epilogue -- no permute / compare / stack passes."""
        if (isinstance(meshes_world, Meshes) and len(meshes_world) > 1) or not self._can_fuse(meshes_world, kwargs):
            rgba = self.forward(meshes_world, **kwargs)
            return rgba[..., :3].permute(0, 3, 1, 2).contiguous(), (rgba[..., 3:4] > 0).to(rgba.dtype).permute(0, 3, 1, 2)
        verts, faces, R, T, size, tex_kw, common = self._fused_args(meshes_world, kwargs)
        images, masks, _ = _fn.render_views(verts, faces, R, T, size, planar=True, **tex_kw, **common)
        return images, masks


__all__ = ["FoVPerspectiveCameras", "look_at_view_transform", "TexturesUV", "TexturesVertex", "RasterizationSettings",
           "BlendParams", "Materials", "AmbientLights", "PointLights", "DirectionalLights", "Fragments",
           "MeshRasterizer", "SoftPhongShader", "MeshRenderer"]

"""General (unfused) shading path: what SoftPhongShader does on top of Fragments for any lights and any
faces_per_pixel.  Per-pixel attribute interpolation runs in the libst3d
operator-boundary kernels (`interp_face_attrs`, with autograd); lighting and blending are elementwise
torch ops on the GPU, as they are in upstream PyTorch3D.  The configuration the reference uses
(AmbientLights, faces_per_pixel = 1) never comes here: MeshRenderer fuses it into the raster epilogue."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from st3d import functional as _fn


def _local_faces(pix_to_face, num_faces):
    """Packed (view-major) face index -> index into the mesh's own face list; -1 stays -1."""
    return torch.where(pix_to_face >= 0, pix_to_face % num_faces, pix_to_face)


def sample_textures_uv(fragments, textures, num_faces):
    """(N,H,W,K,3) texels: bilinear, border padding, align_corners=True, v flipped (A.4)."""
    p2f = _local_faces(fragments.pix_to_face, num_faces)
    N, H, W, K = p2f.shape
    uv = _fn.interpolate_face_attributes(p2f, fragments.bary_coords, textures.faces_verts_uvs().contiguous())
    grid = torch.stack([2.0 * uv[..., 0] - 1.0, 1.0 - 2.0 * uv[..., 1]], dim=-1)
    grid = grid.permute(0, 3, 1, 2, 4).reshape(N * K, H, W, 2)
    maps = textures.maps_padded().permute(0, 3, 1, 2)                     # (1,C,Ht,Wt)
    maps = maps.expand(N * K, -1, -1, -1)
    texels = F.grid_sample(maps, grid, mode="bilinear", padding_mode="border", align_corners=True)
    return texels.reshape(N, K, -1, H, W).permute(0, 3, 4, 1, 2)


def sample_textures_vertex(fragments, textures, faces, num_faces):
    p2f = _local_faces(fragments.pix_to_face, num_faces)
    attrs = textures.verts_features_packed()[faces.long()].contiguous()    # (F,3,C)
    return _fn.interpolate_face_attributes(p2f, fragments.bary_coords, attrs)


def vertex_normals(verts, faces):
    """Area-weighted vertex normals: sum of (v2 - v1) x (v0 - v1) over incident faces, normalised."""
    f = faces.long()
    v0, v1, v2 = verts[f[:, 0]], verts[f[:, 1]], verts[f[:, 2]]
    fn = torch.cross(v2 - v1, v0 - v1, dim=1)
    vn = torch.zeros_like(verts)
    for i in range(3):
        vn = vn.index_add(0, f[:, i], fn)
    return F.normalize(vn, eps=1e-6, dim=1)


def phong_shading(fragments, verts, faces, texels, lights, materials, camera_center):
    """colors = (ambient + diffuse) * texels + specular  (A.5), world-space positions and normals."""
    dev, dt = texels.device, texels.dtype
    col = lambda c: torch.tensor(c, device=dev, dtype=dt)
    ambient = col(materials.ambient_color) * col(lights.ambient_color)
    kind = getattr(lights, "kind", "ambient")
    if kind == "ambient":
        return ambient * texels
    num_faces = faces.shape[0]
    p2f = _local_faces(fragments.pix_to_face, num_faces)
    f = faces.long()
    pos = _fn.interpolate_face_attributes(p2f, fragments.bary_coords, verts[f].contiguous())
    nrm = _fn.interpolate_face_attributes(p2f, fragments.bary_coords, vertex_normals(verts, faces)[f].contiguous())
    direction = (col(lights.location) - pos) if kind == "point" else col(lights.direction).expand_as(pos)
    n_hat = F.normalize(nrm, eps=1e-6, dim=-1)
    l_hat = F.normalize(direction, eps=1e-6, dim=-1)
    cos = (n_hat * l_hat).sum(-1)
    diffuse = col(materials.diffuse_color) * col(lights.diffuse_color) * torch.relu(cos)[..., None]
    N = p2f.shape[0]
    view = F.normalize(camera_center.to(dev, dt).view(N, 1, 1, 1, 3) - pos, eps=1e-6, dim=-1)
    refl = -l_hat + 2.0 * cos[..., None] * n_hat
    alpha = torch.relu((view * refl).sum(-1)) * (cos > 0).to(dt)
    specular = col(materials.specular_color) * col(lights.specular_color) * torch.pow(alpha, materials.shininess)[..., None]
    return (ambient + diffuse) * texels + specular


def softmax_rgb_blend(colors, fragments, blend_params, znear: float = 1.0, zfar: float = 100.0):
    """(N,H,W,K,3) colours + fragments -> (N,H,W,4) RGBA (A.6)."""
    dt = colors.dtype
    mask = (fragments.pix_to_face >= 0).to(dt)
    prob = torch.sigmoid(-fragments.dists / blend_params.sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    z_inv = (zfar - fragments.zbuf) / (zfar - znear) * mask
    z_max = z_inv.max(dim=-1, keepdim=True).values.clamp(min=1e-10)
    weights = prob * torch.exp((z_inv - z_max) / blend_params.gamma)
    delta = torch.exp((1e-10 - z_max) / blend_params.gamma).clamp(min=1e-10)
    bg = torch.tensor(tuple(blend_params.background_color), device=colors.device, dtype=dt)
    rgb = ((weights[..., None] * colors).sum(dim=-2) + delta * bg) / (weights.sum(dim=-1, keepdim=True) + delta)
    return torch.cat([rgb, (1.0 - alpha)[..., None]], dim=-1)

"""Near-plane face clipping for the operator-boundary (Fragments) path.

Upstream PyTorch3D clips in Python/torch *around* `_C.rasterize_meshes` (renderer/mesh/clip.py, called from
rasterize_meshes.py with `z_clip_value = znear / 2` inferred by the MeshRasterizer the reference builds at
first_approach.py:107-111), so the drop-in does the same: batched torch ops
on the device the faces live on, differentiable by autograd, feeding `st3d_rasterize_meshes_forward`.
The fused renderer (`st3d_render_forward`) clips inside its kernels instead (csrc/clip.cuh).

Cases per face, by the number of vertices with z < z_clip:
  0 -> kept as is;  3 -> removed;
  2 -> the one vertex in front (p1) and the two plane crossings form the triangle (p4, p5, p1);
  1 -> the vertex behind (p1) is cut off; the remaining quad is split into (p4, p2, p5) and (p5, p2, p3),
       which are recorded as each other's neighbour.
p2 is the vertex BEFORE p1 in the face, p3 the one after; p4 lies on p1-p2 and p5 on p1-p3.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class ClippedFaces:
    face_verts: torch.Tensor                 # (Fc,3,3)
    mesh_to_face_first_idx: torch.Tensor     # (N,) int64
    num_faces_per_mesh: torch.Tensor         # (N,) int64
    faces_clipped_to_unclipped_idx: Optional[torch.Tensor] = None   # (Fc,) int64; None = nothing was clipped
    barycentric_conversion: Optional[torch.Tensor] = None           # (Fc,3,3): column k = unclipped barycentrics of vertex k
    was_clipped: Optional[torch.Tensor] = None                       # (Fc,) bool
    clipped_faces_neighbor_idx: Optional[torch.Tensor] = None        # (Fc,) int64, -1 = none


def _crossings(tris: torch.Tensor, i1: torch.Tensor, z_clip: float, perspective_correct: bool):
    """tris (T,3,3), i1 (T,) index of the vertex that is alone on its side of the plane."""
    i2, i3 = (i1 + 2) % 3, (i1 + 1) % 3
    rows = torch.arange(tris.shape[0], device=tris.device)
    p1, p2, p3 = tris[rows, i1], tris[rows, i2], tris[rows, i3]
    w2 = ((p1[:, 2] - z_clip) / (p1[:, 2] - p2[:, 2]))[:, None]
    w3 = ((p1[:, 2] - z_clip) / (p1[:, 2] - p3[:, 2]))[:, None]
    if perspective_correct:
        def unproject(p):
            return torch.cat([p[:, :2] * p[:, 2:3], p[:, 2:3]], dim=1)

        def project(q):
            return torch.cat([q[:, :2] / q[:, 2:3], q[:, 2:3]], dim=1)
        q1, q2, q3 = unproject(p1), unproject(p2), unproject(p3)
        p4 = project(q1 * (1 - w2) + q2 * w2)
        p5 = project(q1 * (1 - w3) + q3 * w3)
    else:
        p4 = p1 * (1 - w2) + p2 * w2
        p5 = p1 * (1 - w3) + p3 * w3
    one_hot = torch.nn.functional.one_hot
    e1, e2, e3 = (one_hot(i, 3).to(tris.dtype) for i in (i1, i2, i3))
    b4 = e1 * (1 - w2) + e2 * w2
    b5 = e1 * (1 - w3) + e3 * w3
    return (p1, p2, p3, p4, p5), (e1, e2, e3, b4, b5)


def frustum_culled(face_verts: torch.Tensor) -> torch.Tensor:
    """(F,) bool: faces that upstream's `cull_to_frustum=True` removes -- all three vertices beyond ONE of the side
    planes x = -1, x = 1, y = -1, y = 1 of the NDC frustum (ClipFrustum(left=-1, right=1, top=-1, bottom=1): the z
    planes are not set by the rasterizer)."""
    x, y = face_verts[:, :, 0].detach(), face_verts[:, :, 1].detach()
    return ((x < -1.0).all(dim=1) | (x > 1.0).all(dim=1) | (y < -1.0).all(dim=1) | (y > 1.0).all(dim=1))


def clip_faces(face_verts: torch.Tensor, mesh_to_face_first_idx: torch.Tensor, num_faces_per_mesh: torch.Tensor,
               z_clip: Optional[float], perspective_correct: bool = True, cull_to_frustum: bool = False) -> ClippedFaces:
    """Returns the inputs unchanged (conversion fields None) when no vertex lies behind the plane and nothing is
    culled -- the case of every scene of the reference.  z_clip None: culling only."""
    if z_clip is None:
        behind = torch.zeros(face_verts.shape[:2], dtype=torch.bool, device=face_verts.device)
        z_clip = 0.0
    else:
        behind = face_verts[:, :, 2].detach() < z_clip
    nb = behind.sum(dim=1)
    culled = frustum_culled(face_verts) if cull_to_frustum else None
    if not bool(nb.any()) and not (culled is not None and bool(culled.any())):   # host read, as upstream's early return
        return ClippedFaces(face_verts, mesh_to_face_first_idx, num_faces_per_mesh)
    if culled is not None:
        nb = torch.where(culled, torch.full_like(nb, 3), nb)        # a culled face goes the way of one behind the plane
    dev, dt, Fu = face_verts.device, face_verts.dtype, face_verts.shape[0]
    count = torch.where(nb == 1, 2, torch.where(nb == 3, 0, 1))
    csum = torch.cat([count.new_zeros(1), count.cumsum(0)])
    start = csum[:-1]
    Fc = int(csum[-1])
    first = mesh_to_face_first_idx.to(dev)
    num = num_faces_per_mesh.to(dev)
    new_first, new_num = csum[first], csum[first + num] - csum[first]

    keep = (nb == 0).nonzero()[:, 0]
    two = (nb == 2).nonzero()[:, 0]     # two vertices behind -> one smaller triangle
    one = (nb == 1).nonzero()[:, 0]     # one vertex behind  -> two triangles
    fv = face_verts.new_zeros((Fc, 3, 3))
    conv = face_verts.new_zeros((Fc, 3, 3))
    eye = torch.eye(3, device=dev, dtype=dt)
    fv = fv.index_copy(0, start[keep], face_verts[keep])
    conv = conv.index_copy(0, start[keep], eye.expand(keep.numel(), 3, 3))
    if two.numel():
        i1 = (~behind[two]).to(torch.int64).argmax(dim=1)
        (p1, _, _, p4, p5), (e1, _, _, b4, b5) = _crossings(face_verts[two], i1, z_clip, perspective_correct)
        fv = fv.index_copy(0, start[two], torch.stack([p4, p5, p1], dim=1))
        conv = conv.index_copy(0, start[two], torch.stack([b4, b5, e1], dim=2))
    if one.numel():
        i1 = behind[one].to(torch.int64).argmax(dim=1)
        (_, p2, p3, p4, p5), (_, e2, e3, b4, b5) = _crossings(face_verts[one], i1, z_clip, perspective_correct)
        fv = fv.index_copy(0, start[one], torch.stack([p4, p2, p5], dim=1))
        conv = conv.index_copy(0, start[one], torch.stack([b4, e2, b5], dim=2))
        fv = fv.index_copy(0, start[one] + 1, torch.stack([p5, p2, p3], dim=1))
        conv = conv.index_copy(0, start[one] + 1, torch.stack([b5, e2, e3], dim=2))
    to_unclipped = torch.repeat_interleave(torch.arange(Fu, device=dev), count)
    was = (nb[to_unclipped] > 0)
    neighbor = torch.full((Fc,), -1, device=dev, dtype=torch.int64)
    neighbor[start[one]] = start[one] + 1
    neighbor[start[one] + 1] = start[one]
    return ClippedFaces(fv, new_first, new_num, to_unclipped, conv, was, neighbor)


def convert_clipped_rasterization_to_original_faces(pix_to_face: torch.Tensor, bary: torch.Tensor,
                                                    clipped: ClippedFaces):
    """pix_to_face -> indices into the unclipped face list; barycentrics of clipped faces are mapped to the
    unclipped face: b_unclipped = barycentric_conversion @ b_clipped."""
    if clipped.faces_clipped_to_unclipped_idx is None:
        return pix_to_face, bary
    mask = pix_to_face >= 0
    idx = pix_to_face.clamp(min=0)
    p2f = torch.where(mask, clipped.faces_clipped_to_unclipped_idx[idx], pix_to_face)
    conv = clipped.barycentric_conversion[idx]                               # (...,3,3)
    b = torch.matmul(conv, bary[..., None])[..., 0]
    use = (mask & clipped.was_clipped[idx])[..., None]
    return p2f, torch.where(use, b, bary)

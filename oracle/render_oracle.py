"""oracle/render_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch + the C library built from raster_oracle.c) of the render half of the
reference's hot path.  The reference itself holds no renderer code: `utils.py:65-77`
(`render_meshes`) calls PyTorch3D's `MeshRenderer(MeshRasterizer, SoftPhongShader)` built at
`first_approach.py:106-114` / `second_approach.py:100-108`.  PyTorch3D is an un-vendored,
un-pinned third-party dependency that is absent from /root/reference and from this image, so the
functions below restate its published algorithm as recorded in SURVEY.md Appendix A
(A.1 cameras, A.2 raster settings, A.3 rasterization, A.4 texture sampling, A.5 Phong,
A.6 softmax_rgb_blend).  PARITY UNPINNED at that boundary (no upstream golden vectors on disk).
What pins the restatement instead (tests/test_oracle_pinned.py): the definitions of A.3 / A.4 / A.6 recomputed in exact
rational arithmetic for a triangle with three different depths (inside test, perspective-corrected barycentrics, depth,
signed edge distance of every pixel), hand-derived clip geometry, and texture sampling through torch's own grid_sample.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
K_EPS = 1e-8


def build(force: bool = False) -> str:
    """Compile raster_oracle.c -> liboracle.so (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.oracle_transform_verts.argtypes = [fp, ctypes.c_int64, fp, fp, ctypes.c_float, ctypes.c_float, fp]
        L.oracle_transform_verts.restype = None
        L.oracle_rasterize_naive.argtypes = [fp, ip, ip] + [ctypes.c_int] * 3 + [ctypes.c_float] + \
            [ctypes.c_int] * 5 + [ip, fp, fp, fp]
        L.oracle_rasterize_naive.restype = ctypes.c_int
        L.oracle_rasterize_naive_nb.argtypes = [fp, ip, ip, ip] + [ctypes.c_int] * 3 + [ctypes.c_float] + \
            [ctypes.c_int] * 5 + [ip, fp, fp, fp]
        L.oracle_rasterize_naive_nb.restype = ctypes.c_int
        L.oracle_pix_to_ndc.argtypes = [ctypes.c_int] * 3
        L.oracle_pix_to_ndc.restype = ctypes.c_float
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int64))


# --------------------------------------------------------------------------------------------
# A.1 cameras (reference call sites: utils.py:121-170)
# --------------------------------------------------------------------------------------------
def look_at_view_transform(dist, elev, azim, at=((0.0, 0.0, 0.0),), up=((0.0, 1.0, 0.0),)):
    """SURVEY A.1 `look_at_view_transform(degrees=True)`; used by utils.py:161-166."""
    elev = torch.as_tensor(elev, dtype=torch.float32).reshape(-1)
    azim = torch.as_tensor(azim, dtype=torch.float32).reshape(-1)
    dist = torch.as_tensor(dist, dtype=torch.float32).reshape(-1)
    n = max(elev.numel(), azim.numel(), dist.numel())
    elev, azim, dist = elev.expand(n), azim.expand(n), dist.expand(n)
    at = torch.as_tensor(at, dtype=torch.float32).reshape(-1, 3).expand(n, 3)
    up = torch.as_tensor(up, dtype=torch.float32).reshape(-1, 3).expand(n, 3)
    e = elev * (math.pi / 180.0)
    a = azim * (math.pi / 180.0)
    cam = torch.stack([dist * torch.cos(e) * torch.sin(a), dist * torch.sin(e),
                       dist * torch.cos(e) * torch.cos(a)], dim=1) + at
    z_axis = F.normalize(at - cam, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    close = torch.isclose(x_axis, torch.zeros(()), atol=5e-3).all(dim=1, keepdim=True)
    if close.any():
        x_axis = torch.where(close, F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5), x_axis)
    R = torch.stack([x_axis, y_axis, z_axis], dim=1).transpose(1, 2)  # columns are the axes
    T = -torch.bmm(R.transpose(1, 2), cam[:, :, None])[:, :, 0]
    return R.contiguous(), T.contiguous()


def rotate_axis_angle_R(angle_deg: float, axis: str) -> torch.Tensor:
    """3x3 block of RotateAxisAngle(...).get_matrix() as sliced at utils.py:142."""
    a = torch.tensor(float(angle_deg), dtype=torch.float32) * (math.pi / 180.0)
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones(()), torch.zeros(())
    if axis == "X":
        m = torch.stack([o, z, z, z, c, -s, z, s, c])
    elif axis == "Y":
        m = torch.stack([c, z, s, z, o, z, -s, z, c])
    elif axis == "Z":
        m = torch.stack([c, -s, z, s, c, z, z, z, o])
    else:
        raise ValueError(axis)
    return m.reshape(3, 3).t().contiguous()  # transpose of the column-vector rotation


def fixed_cameras(n_views: int, dist: float = 3.0):
    """utils.py:121-151 with shuffle=False."""
    x_views = n_views // 2
    y_views = n_views - x_views
    angles = [(float(a), "X") for a in torch.linspace(0, 315, x_views)] + \
             [(float(a), "Y") for a in torch.linspace(45, 315, y_views)]
    R = torch.stack([rotate_axis_angle_R(a, ax) for a, ax in angles], dim=0)
    T = torch.tensor([[0.0, 0.0, dist]]).repeat(len(angles), 1)
    return R, T


def random_cameras(n_views: int, dist: float = 2.10, generator=None):
    """utils.py:154-170 (cos-elevation uniform, azimuth uniform, at=(0,0.10,0.25))."""
    cos_elevs = torch.rand(n_views, generator=generator) * 2 - 1
    elevs = torch.acos(cos_elevs) * 180 / torch.pi - 90
    azims = torch.rand(n_views, generator=generator) * 360 - 180
    return look_at_view_transform(dist, elevs, azims, at=((0, 0.10, 0.25),))


def fov_scales(fov_deg: float = 60.0, aspect: float = 1.0, znear: float = 1.0):
    """K00, K11 of the FoV projection, rounded to fp32 on the host."""
    f32 = np.float32
    tan_half = f32(math.tan(f32(fov_deg) * f32(math.pi / 180.0) / 2.0))
    max_y = f32(tan_half * f32(znear))
    max_x = f32(max_y * f32(aspect))
    k00 = f32(f32(2.0) * f32(znear) / f32(max_x - (-max_x)))
    k11 = f32(f32(2.0) * f32(znear) / f32(max_y - (-max_y)))
    return float(k00), float(k11)


def transform_verts_exact(verts, R, T, k00, k11):
    """(V,3) world -> (N,V,3) [x_ndc, y_ndc, z_view], exact op order of raster_oracle.c."""
    verts = verts.detach().to(torch.float32).contiguous().cpu()
    R = R.detach().to(torch.float32).contiguous().cpu().reshape(-1, 3, 3)
    T = T.detach().to(torch.float32).contiguous().cpu().reshape(-1, 3)
    out = torch.empty((R.shape[0],) + tuple(verts.shape), dtype=torch.float32)
    for n in range(R.shape[0]):
        lib().oracle_transform_verts(_fp(verts), verts.shape[0], _fp(R[n]), _fp(T[n]), k00, k11, _fp(out[n]))
    return out


def transform_verts_torch(verts, R, T, k00, k11):
    """Differentiable version of the same map (any float dtype)."""
    R = R.to(verts.dtype).reshape(-1, 3, 3)
    T = T.to(verts.dtype).reshape(-1, 3)
    view = torch.einsum("vi,nij->nvj", verts, R) + T[:, None, :]
    z = view[..., 2]
    return torch.stack([view[..., 0] * k00 / z, view[..., 1] * k11 / z, z], dim=-1)


# --------------------------------------------------------------------------------------------
# A.3 rasterization
# --------------------------------------------------------------------------------------------
def rasterize_naive(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size,
                    blur_radius=0.0, faces_per_pixel=1, perspective_correct=True,
                    clip_barycentric_coords=False, cull_backfaces=False, nthreads=1, clipped_faces_neighbor_idx=None):
    """Mirrors pytorch3d._C.rasterize_meshes (naive path).  Returns pix_to_face i64, zbuf, bary, dists.
    clipped_faces_neighbor_idx (F,) int64 or None: the two halves of a clipped quad de-duplicate per pixel (A.2)."""
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    fv = face_verts.detach().to(torch.float32).contiguous().cpu()
    first = mesh_to_face_first_idx.to(torch.int64).contiguous().cpu()
    num = num_faces_per_mesh.to(torch.int64).contiguous().cpu()
    N, K = first.numel(), int(faces_per_pixel)
    p2f = torch.empty((N, H, W, K), dtype=torch.int64)
    zbuf = torch.empty((N, H, W, K), dtype=torch.float32)
    bary = torch.empty((N, H, W, K, 3), dtype=torch.float32)
    dists = torch.empty((N, H, W, K), dtype=torch.float32)
    if clipped_faces_neighbor_idx is None:
        rc = lib().oracle_rasterize_naive(_fp(fv), _ip(first), _ip(num), N, H, W, float(blur_radius), K,
                                          int(perspective_correct), int(clip_barycentric_coords),
                                          int(cull_backfaces), int(nthreads), _ip(p2f), _fp(zbuf),
                                          _fp(bary), _fp(dists))
    else:
        nbi = clipped_faces_neighbor_idx.to(torch.int64).contiguous().cpu()
        rc = lib().oracle_rasterize_naive_nb(_fp(fv), _ip(first), _ip(num), _ip(nbi), N, H, W, float(blur_radius), K,
                                             int(perspective_correct), int(clip_barycentric_coords),
                                             int(cull_backfaces), int(nthreads), _ip(p2f), _fp(zbuf),
                                             _fp(bary), _fp(dists))
    if rc != 0:
        raise ValueError("oracle_rasterize_naive: bad arguments")
    return p2f, zbuf, bary, dists


def pixel_ndc_grid(H, W, dtype=torch.float32):
    """NDC coordinates of output pixel centres (A.3): both axes flipped (+X left, +Y up)."""
    ys = torch.tensor([lib().oracle_pix_to_ndc(H - 1 - yi, H, W) for yi in range(H)], dtype=dtype)
    xs = torch.tensor([lib().oracle_pix_to_ndc(W - 1 - xi, W, H) for xi in range(W)], dtype=dtype)
    return xs, ys


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist2(px, py, ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    degenerate = l2 <= K_EPS
    l2s = torch.where(degenerate, torch.ones_like(l2), l2)
    t = ((dx * (px - ax) + dy * (py - ay)) / l2s).clamp(0.0, 1.0)
    qx, qy = ax + t * dx, ay + t * dy
    d = (px - qx) ** 2 + (py - qy) ** 2
    return torch.where(degenerate, (px - bx) ** 2 + (py - by) ** 2, d)


def fragments_from_faces(face_verts, pix_to_face, perspective_correct=True, clip_barycentric_coords=False):
    """Differentiable recompute of (zbuf, bary, dists) for a FIXED pix_to_face (A.3 steps 3-8).

    This is the autograd oracle for rasterize_meshes_backward (SURVEY section 4 / A.3 "Backward").
    """
    N, H, W, K = pix_to_face.shape
    dt = face_verts.dtype
    mask = pix_to_face >= 0
    fv = face_verts[pix_to_face.clamp(min=0)]  # (N,H,W,K,3,3)
    xs, ys = pixel_ndc_grid(H, W, dt)
    px = xs.view(1, 1, W, 1).expand(N, H, W, K)
    py = ys.view(1, H, 1, 1).expand(N, H, W, K)
    x0, y0, z0 = fv[..., 0, 0], fv[..., 0, 1], fv[..., 0, 2]
    x1, y1, z1 = fv[..., 1, 0], fv[..., 1, 1], fv[..., 1, 2]
    x2, y2, z2 = fv[..., 2, 0], fv[..., 2, 1], fv[..., 2, 2]
    area = _edge(x2, y2, x0, y0, x1, y1)
    denom = torch.where(mask, area + K_EPS, torch.ones_like(area))
    b0 = _edge(px, py, x1, y1, x2, y2) / denom
    b1 = _edge(px, py, x2, y2, x0, y0) / denom
    b2 = _edge(px, py, x0, y0, x1, y1) / denom
    if perspective_correct:
        t0, t1, t2 = b0 * z1 * z2, z0 * b1 * z2, z0 * z1 * b2
        d = (t0 + t1 + t2).clamp(min=K_EPS)
        b0, b1, b2 = t0 / d, t1 / d, t2 / d
    inside = (b0 > 0) & (b1 > 0) & (b2 > 0)
    if clip_barycentric_coords:
        c0, c1, c2 = b0.clamp(0, 1), b1.clamp(0, 1), b2.clamp(0, 1)
        s = (c0 + c1 + c2).clamp(min=1e-5)
        b0, b1, b2 = c0 / s, c1 / s, c2 / s
    pz = b0 * z0 + b1 * z1 + b2 * z2
    dist = torch.minimum(_seg_dist2(px, py, x0, y0, x1, y1),
                         torch.minimum(_seg_dist2(px, py, x0, y0, x2, y2),
                                       _seg_dist2(px, py, x1, y1, x2, y2)))
    dist = torch.where(inside, -dist, dist)
    neg = -torch.ones((), dtype=dt)
    zbuf = torch.where(mask, pz, neg)
    dists = torch.where(mask, dist, neg)
    bary = torch.where(mask[..., None], torch.stack([b0, b1, b2], dim=-1), neg)
    return zbuf, bary, dists


# --------------------------------------------------------------------------------------------
# A.2 near-plane clipping (upstream renderer/mesh/clip.py: clip_faces +
# convert_clipped_rasterization_to_original_faces; reached from the reference through the
# MeshRasterizer built at first_approach.py:107-111 with z_clip_value = znear / 2)
# --------------------------------------------------------------------------------------------
def _plane_crossings(tri, i1, z_clip, perspective_correct):
    """tri (3,3) rows = vertices [x_ndc, y_ndc, z_view]; i1 = index of the vertex that is alone on its
    side of the plane.  p2 is the PREVIOUS vertex, p3 the NEXT one.  Returns the crossing points p4 (on
    p1-p2), p5 (on p1-p3) and the barycentric coordinates of p4, p5 w.r.t. the unclipped face."""
    i2, i3 = (i1 - 1) % 3, (i1 + 1) % 3
    p1, p2, p3 = tri[i1], tri[i2], tri[i3]
    w2 = (p1[2] - z_clip) / (p1[2] - p2[2])
    w3 = (p1[2] - z_clip) / (p1[2] - p3[2])
    if perspective_correct:      # interpolate in view space (x_ndc * z), then divide again
        def unproject(p):
            return torch.stack([p[0] * p[2], p[1] * p[2], p[2]])

        def project(q):
            return torch.stack([q[0] / q[2], q[1] / q[2], q[2]])
        q1, q2, q3 = unproject(p1), unproject(p2), unproject(p3)
        p4 = project(q1 * (1 - w2) + q2 * w2)
        p5 = project(q1 * (1 - w3) + q3 * w3)
    else:
        p4 = p1 * (1 - w2) + p2 * w2
        p5 = p1 * (1 - w3) + p3 * w3
    zero = torch.zeros((), dtype=tri.dtype)
    b4, b5 = [zero, zero, zero], [zero, zero, zero]
    b4[i1], b4[i2] = 1 - w2, w2
    b5[i1], b5[i3] = 1 - w3, w3
    return p4, p5, torch.stack(b4), torch.stack(b5), i2, i3


def frustum_culled(face_verts):
    """cull_to_frustum (upstream clip_faces with ClipFrustum(left=-1, right=1, top=-1, bottom=1, cull=True)):
    faces whose three vertices all lie beyond ONE of the planes x = -1, x = 1, y = -1, y = 1."""
    x, y = face_verts[:, :, 0].detach(), face_verts[:, :, 1].detach()
    return (x < -1).all(1) | (x > 1).all(1) | (y < -1).all(1) | (y > 1).all(1)


def _with_cull_of(fv, fv_exact):
    """fv (differentiable, any dtype) with the faces the EXACT fp32 coordinates cull pushed far outside the frustum, so
    that the differentiable clip removes the same faces whatever the rounding of fv (a face the exact coordinates keep
    but fv would cull trips the to_unclipped assert in render_views)."""
    culled = frustum_culled(fv_exact)
    if not bool(culled.any()):
        return fv
    shift = torch.zeros_like(fv)
    shift[culled, :, 0] = 1e6
    return fv + shift


def clip_faces(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, z_clip, perspective_correct=True,
               cull_to_frustum=False):
    """Clip every face against the plane z = z_clip.  Per face, by the number of vertices
    with z < z_clip:  0 -> kept;  3 -> removed;  2 (p1 in front) -> the triangle (p4, p5, p1);
    1 (p1 behind) -> the quad split into (p4, p2, p5) and (p5, p2, p3), linked as neighbours.
    Differentiable in the dtype of face_verts.  Returns a dict:
      face_verts (Fc,3,3), first (N), num (N), to_unclipped (Fc) i64, conversion (Fc,3,3) whose COLUMN k
      holds the unclipped-face barycentrics of clipped vertex k (identity for kept faces),
      was_clipped (Fc) bool, neighbor (Fc) i64 (-1 = none)."""
    dt = face_verts.dtype
    first = mesh_to_face_first_idx.tolist()
    num = num_faces_per_mesh.tolist()
    behind = (face_verts[:, :, 2].detach() < z_clip)
    culled = frustum_culled(face_verts) if cull_to_frustum else torch.zeros(face_verts.shape[0], dtype=torch.bool)
    out_fv, to_unc, conv, was, neigh, new_first, new_num = [], [], [], [], [], [], []
    eye = torch.eye(3, dtype=dt)
    for n in range(len(first)):
        new_first.append(len(out_fv))
        for f in range(first[n], first[n] + num[n]):
            nb = 3 if bool(culled[f]) else int(behind[f].sum())     # a culled face is removed like one behind the plane
            tri = face_verts[f]
            if nb == 0:
                out_fv.append(tri); to_unc.append(f); conv.append(eye); was.append(False); neigh.append(-1)
            elif nb == 2:
                i1 = int((~behind[f]).nonzero()[0])
                p4, p5, b4, b5, i2, i3 = _plane_crossings(tri, i1, z_clip, perspective_correct)
                out_fv.append(torch.stack([p4, p5, tri[i1]]))
                conv.append(torch.stack([b4, b5, eye[i1]], dim=1))
                to_unc.append(f); was.append(True); neigh.append(-1)
            elif nb == 1:
                i1 = int(behind[f].nonzero()[0])
                p4, p5, b4, b5, i2, i3 = _plane_crossings(tri, i1, z_clip, perspective_correct)
                k = len(out_fv)
                out_fv.append(torch.stack([p4, tri[i2], p5]))
                conv.append(torch.stack([b4, eye[i2], b5], dim=1))
                out_fv.append(torch.stack([p5, tri[i2], tri[i3]]))
                conv.append(torch.stack([b5, eye[i2], eye[i3]], dim=1))
                to_unc += [f, f]; was += [True, True]; neigh += [k + 1, k]
        new_num.append(len(out_fv) - new_first[-1])
    Fc = len(out_fv)
    return dict(face_verts=torch.stack(out_fv) if Fc else face_verts.new_zeros((0, 3, 3)),
                first=torch.tensor(new_first, dtype=torch.int64), num=torch.tensor(new_num, dtype=torch.int64),
                to_unclipped=torch.tensor(to_unc, dtype=torch.int64),
                conversion=torch.stack(conv) if Fc else face_verts.new_zeros((0, 3, 3)),
                was_clipped=torch.tensor(was, dtype=torch.bool), neighbor=torch.tensor(neigh, dtype=torch.int64))


def convert_clipped_to_unclipped(pix_to_face, bary, clipped):
    """pix_to_face -> indices of the unclipped faces; barycentrics of clipped faces -> barycentrics of
    the unclipped face: b_unclipped = conversion @ b_clipped."""
    mask = pix_to_face >= 0
    idx = pix_to_face.clamp(min=0)
    p2f = torch.where(mask, clipped["to_unclipped"][idx], pix_to_face)
    conv = clipped["conversion"].to(bary.dtype)[idx]                     # (...,3,3)
    b = (conv * bary[..., None, :]).sum(dim=-1)
    use = (mask & clipped["was_clipped"][idx])[..., None]
    return p2f, torch.where(use, b, bary)


# --------------------------------------------------------------------------------------------
# A.4 texture sampling
# --------------------------------------------------------------------------------------------
def interpolate_face_attributes(pix_to_face, bary, face_attrs):
    """out[p] = sum_i bary[p,i] * face_attrs[f(p), i]; 0 where f < 0 (A.4)."""
    mask = pix_to_face < 0
    attrs = face_attrs[pix_to_face.clamp(min=0)]  # (...,3,D)
    out = (bary[..., None] * attrs).sum(dim=-2)
    return out.masked_fill(mask[..., None], 0.0)


def sample_textures_uv(pix_to_face, bary, faces_verts_uvs, maps):
    """TexturesUV.sample_textures (A.4): bilinear, border padding, align_corners=True, v flipped."""
    N, H, W, K = pix_to_face.shape
    uv = interpolate_face_attributes(pix_to_face, bary, faces_verts_uvs)  # (N,H,W,K,2)
    grid = torch.stack([2.0 * uv[..., 0] - 1.0, 1.0 - 2.0 * uv[..., 1]], dim=-1)
    grid = grid.permute(0, 3, 1, 2, 4).reshape(N * K, H, W, 2)
    tex = maps.permute(0, 3, 1, 2)  # (M,3,Ht,Wt)
    if tex.shape[0] == 1 and N > 1:
        tex = tex.expand(N, -1, -1, -1)
    tex = tex[:, None].expand(-1, K, -1, -1, -1).reshape(N * K, *tex.shape[1:])
    texels = F.grid_sample(tex, grid, mode="bilinear", padding_mode="border", align_corners=True)
    return texels.reshape(N, K, -1, H, W).permute(0, 3, 4, 1, 2)  # (N,H,W,K,3)


def sample_textures_vertex(pix_to_face, bary, faces_verts_rgb):
    """TexturesVertex.sample_textures (A.4): barycentric mix of per-vertex colours."""
    return interpolate_face_attributes(pix_to_face, bary, faces_verts_rgb)


# --------------------------------------------------------------------------------------------
# A.5 Phong shading
# --------------------------------------------------------------------------------------------
def vertex_normals(verts, faces):
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    fn = torch.cross(v2 - v1, v0 - v1, dim=1)
    vn = torch.zeros_like(verts)
    for i in range(3):
        vn = vn.index_add(0, faces[:, i], fn)
    return F.normalize(vn, eps=1e-6, dim=1)


def phong_colors(texels, pix_to_face, bary, verts_world, faces, lights, materials, camera_center):
    """colors = (ambient + diffuse) * texels + specular (A.5).

    lights: dict(kind='ambient'|'point'|'directional', ambient=(3,), diffuse=(3,), specular=(3,),
    location=(3,) or direction=(3,)).  materials: dict(ambient, diffuse, specular, shininess).
    camera_center: (N,3).
    """
    dt = texels.dtype
    la = torch.as_tensor(lights["ambient"], dtype=dt)
    ma = torch.as_tensor(materials["ambient"], dtype=dt)
    ambient = ma * la
    if lights["kind"] == "ambient":
        return ambient * texels
    fv = verts_world[faces]  # (F,3,3)
    vn = vertex_normals(verts_world, faces)[faces]
    N, H, W, K = pix_to_face.shape
    F_ = faces.shape[0]
    local = torch.where(pix_to_face >= 0, pix_to_face % F_, pix_to_face)
    pos = interpolate_face_attributes(local, bary, fv)
    nrm = interpolate_face_attributes(local, bary, vn)
    if lights["kind"] == "point":
        direction = torch.as_tensor(lights["location"], dtype=dt) - pos
    else:
        direction = torch.as_tensor(lights["direction"], dtype=dt).expand_as(pos)
    n_hat = F.normalize(nrm, eps=1e-6, dim=-1)
    l_hat = F.normalize(direction, eps=1e-6, dim=-1)
    cos = (n_hat * l_hat).sum(-1)
    diffuse = torch.as_tensor(materials["diffuse"], dtype=dt) * torch.as_tensor(lights["diffuse"], dtype=dt) \
        * torch.relu(cos)[..., None]
    view = F.normalize(camera_center.to(dt).view(N, 1, 1, 1, 3) - pos, eps=1e-6, dim=-1)
    refl = -l_hat + 2.0 * cos[..., None] * n_hat
    alpha = torch.relu((view * refl).sum(-1)) * (cos > 0).to(dt)
    spec = torch.as_tensor(materials["specular"], dtype=dt) * torch.as_tensor(lights["specular"], dtype=dt) \
        * torch.pow(alpha, float(materials["shininess"]))[..., None]
    return (ambient + diffuse) * texels + spec


# --------------------------------------------------------------------------------------------
# A.6 softmax_rgb_blend
# --------------------------------------------------------------------------------------------
def softmax_rgb_blend(colors, pix_to_face, dists, zbuf, sigma=1e-4, gamma=1e-4,
                      background=(1.0, 1.0, 1.0), znear=1.0, zfar=100.0):
    dt = colors.dtype
    mask = (pix_to_face >= 0).to(dt)
    prob = torch.sigmoid(-dists / sigma) * mask
    alpha = torch.prod(1.0 - prob, dim=-1)
    z_inv = (zfar - zbuf) / (zfar - znear) * mask
    z_max = z_inv.max(dim=-1, keepdim=True).values.clamp(min=1e-10)
    w = prob * torch.exp((z_inv - z_max) / gamma)
    delta = torch.exp((1e-10 - z_max) / gamma).clamp(min=1e-10)
    denom = w.sum(dim=-1, keepdim=True) + delta
    bg = torch.as_tensor(background, dtype=dt)
    rgb = ((w[..., None] * colors).sum(dim=-2) + delta * bg) / denom
    return torch.cat([rgb, (1.0 - alpha)[..., None]], dim=-1)


# --------------------------------------------------------------------------------------------
# Whole renderer call = what utils.py:69 expands to, batched over views
# --------------------------------------------------------------------------------------------
AMBIENT_LIGHTS = dict(kind="ambient", ambient=(1.0, 1.0, 1.0), diffuse=(0.0, 0.0, 0.0), specular=(0.0, 0.0, 0.0))
DEFAULT_MATERIALS = dict(ambient=(1.0, 1.0, 1.0), diffuse=(1.0, 1.0, 1.0), specular=(1.0, 1.0, 1.0), shininess=64.0)


def render_views(verts, faces, R, T, image_size, texture=None, verts_uvs=None, faces_uvs=None,
                 verts_rgb=None, blur_radius=0.0, faces_per_pixel=1, fov=60.0, znear=1.0, zfar=100.0,
                 sigma=1e-4, gamma=1e-4, background=(1.0, 1.0, 1.0), lights=None, materials=None,
                 nthreads=1, return_fragments=False, z_clip=None, cull_to_frustum=False):
    """Render N views of one mesh.  Differentiable w.r.t. verts / texture / verts_rgb.
    z_clip: near-plane clip depth (None = znear / 2, what MeshRasterizer infers for perspective cameras).

    Coverage (pix_to_face) comes from the exact C rasterizer on exactly-transformed fp32 verts;
    zbuf/bary/dists are then recomputed differentiably in the dtype of `verts`.
    Returns rgba (N,H,W,4) [and the fragments].
    """
    lights = lights or AMBIENT_LIGHTS
    materials = materials or DEFAULT_MATERIALS
    H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
    k00, k11 = fov_scales(fov, 1.0, znear)
    N, Fn = R.shape[0], faces.shape[0]
    faces = faces.to(torch.int64)
    ndc_exact = transform_verts_exact(verts, R, T, k00, k11)  # (N,V,3) fp32
    fv_exact = ndc_exact[:, faces].reshape(N * Fn, 3, 3)
    first = torch.arange(N, dtype=torch.int64) * Fn
    num = torch.full((N,), Fn, dtype=torch.int64)
    z_clip = znear / 2.0 if z_clip is None else z_clip
    ndc = transform_verts_torch(verts, R, T, k00, k11)
    fv = ndc[:, faces].reshape(N * Fn, 3, 3)
    if bool((fv_exact[:, :, 2] < z_clip).any()) or (cull_to_frustum and bool(frustum_culled(fv_exact).any())):
        # A.2: some face crosses (or lies behind) the clip plane, or is culled
        cl_x = clip_faces(fv_exact, first, num, np.float32(z_clip).item(), True, cull_to_frustum)
        p2f_c, zbuf_x, bary_x, dists_x = rasterize_naive(cl_x["face_verts"], cl_x["first"], cl_x["num"], (H, W),
                                                          blur_radius, faces_per_pixel, True, blur_radius > 0, False,
                                                          nthreads, clipped_faces_neighbor_idx=cl_x["neighbor"])
        cl = clip_faces(_with_cull_of(fv, fv_exact) if cull_to_frustum else fv, first, num, z_clip, True, cull_to_frustum)
        assert torch.equal(cl["to_unclipped"], cl_x["to_unclipped"])
        zbuf, bary_c, dists = fragments_from_faces(cl["face_verts"], p2f_c, True, blur_radius > 0)
        p2f, bary = convert_clipped_to_unclipped(p2f_c, bary_c, cl)
        _, bary_x = convert_clipped_to_unclipped(p2f_c, bary_x, cl_x)
    else:
        p2f, zbuf_x, bary_x, dists_x = rasterize_naive(fv_exact, first, num, (H, W), blur_radius, faces_per_pixel,
                                                        True, blur_radius > 0, False, nthreads)
        zbuf, bary, dists = fragments_from_faces(fv, p2f, True, blur_radius > 0)
    local = torch.where(p2f >= 0, p2f % Fn, p2f)
    if texture is not None:
        fuv = verts_uvs.to(verts.dtype)[faces_uvs.to(torch.int64)]
        texels = sample_textures_uv(local, bary, fuv, texture.reshape(1, *texture.shape[-3:]))
    else:
        texels = sample_textures_vertex(local, bary, verts_rgb[faces])
    cam_center = -torch.einsum("ni,nji->nj", T.to(verts.dtype), R.to(verts.dtype))  # -T . R^T
    colors = phong_colors(texels, p2f, bary, verts, faces, lights, materials, cam_center)
    rgba = softmax_rgb_blend(colors, p2f, dists, zbuf, sigma, gamma, background, znear, zfar)
    if return_fragments:
        return rgba, dict(pix_to_face=p2f, zbuf=zbuf, bary=bary, dists=dists,
                          zbuf_exact=zbuf_x, bary_exact=bary_x, dists_exact=dists_x)
    return rgba


def images_and_masks(rgba):
    """utils.py:70-76: RGB -> (B,3,H,W); mask = (alpha > 0).float() -> (B,1,H,W)."""
    return rgba[..., :3].permute(0, 3, 1, 2), (rgba[..., 3] > 0).to(rgba.dtype)[:, None]


# --------------------------------------------------------------------------------------------
# Minimal OBJ reader for fixtures (A.8: 0-based int64, fan triangulation, no v flip)
# --------------------------------------------------------------------------------------------
def read_obj(path):
    verts, uvs, f_v, f_t = [], [], [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(t) for t in tok[1:4]])
            elif tok[0] == "vt":
                uvs.append([float(t) for t in tok[1:3]])
            elif tok[0] == "f":
                vi, ti = [], []
                for c in tok[1:]:
                    parts = c.split("/")
                    vi.append(int(parts[0]))
                    ti.append(int(parts[1]) if len(parts) > 1 and parts[1] else 0)
                vi = [i - 1 if i > 0 else len(verts) + i for i in vi]
                ti = [i - 1 if i > 0 else (len(uvs) + i if i < 0 else -1) for i in ti]
                for j in range(1, len(vi) - 1):
                    f_v.append([vi[0], vi[j], vi[j + 1]])
                    f_t.append([ti[0], ti[j], ti[j + 1]])
    return (torch.tensor(verts, dtype=torch.float32), torch.tensor(f_v, dtype=torch.int64),
            torch.tensor(uvs, dtype=torch.float32).reshape(-1, 2), torch.tensor(f_t, dtype=torch.int64))

"""Camera math used by the reference's view builders (utils.py:121-170), restated from the published
PyTorch3D conventions: row-vector transforms X_view = X_world . R + T,
+X left / +Y up NDC, look-at extrinsics, axis-angle rotations.  Tiny host-side (or any-device) torch
math; the projection itself runs inside the raster kernels (st3d_transform_verts_forward)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _vec3(v, n, device):
    t = torch.as_tensor(v, dtype=torch.float32, device=device)
    return t.reshape(-1, 3).expand(n, 3)


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees: bool = True, at=((0.0, 0.0, 0.0),),
                           up=((0.0, 1.0, 0.0),), device="cpu"):
    """(R (N,3,3), T (N,3)) of cameras on a sphere of radius `dist` around `at` looking at `at`."""
    vals = [torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1) for v in (dist, elev, azim)]
    n = max(v.numel() for v in vals)
    dist, elev, azim = (v.expand(n) for v in vals)
    if degrees:
        elev, azim = elev * (math.pi / 180.0), azim * (math.pi / 180.0)
    at, up = _vec3(at, n, device), _vec3(up, n, device)
    centre = torch.stack([dist * torch.cos(elev) * torch.sin(azim), dist * torch.sin(elev),
                          dist * torch.cos(elev) * torch.cos(azim)], dim=1) + at
    z = F.normalize(at - centre, eps=1e-5)
    x = F.normalize(torch.cross(up, z, dim=1), eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    degenerate = torch.isclose(x, torch.zeros((), device=x.device), atol=5e-3).all(dim=1, keepdim=True)
    if bool(degenerate.any()):
        x = torch.where(degenerate, F.normalize(torch.cross(y, z, dim=1), eps=1e-5), x)
    R = torch.stack([x, y, z], dim=2)                       # columns are the camera axes
    T = -torch.einsum("nji,nj->ni", R, centre)              # -R^T C
    return R.contiguous(), T.contiguous()


def rotate_axis_angle_matrix(angle, axis: str = "X", degrees: bool = True, device="cpu") -> torch.Tensor:
    """(1,4,4) matrix of RotateAxisAngle(angle, axis).get_matrix(): the 3x3 block is the transpose of the
    column-vector rotation (row-vector convention)."""
    a = torch.as_tensor(angle, dtype=torch.float32, device=device).reshape(())
    if degrees:
        a = a * (math.pi / 180.0)
    c, s = torch.cos(a), torch.sin(a)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    rows = {"X": [one, zero, zero, zero, c, -s, zero, s, c],
            "Y": [c, zero, s, zero, one, zero, -s, zero, c],
            "Z": [c, -s, zero, s, c, zero, zero, zero, one]}
    if axis not in rows:
        raise ValueError(f"axis must be one of X, Y, Z; got {axis!r}")
    m = torch.eye(4, dtype=torch.float32, device=device)
    m[:3, :3] = torch.stack(rows[axis]).reshape(3, 3).t()
    return m[None]


def random_view_cameras(n_views: int, dist: float = 2.10, at=((0.0, 0.10, 0.25),), generator=None, device="cpu"):
    """The sampling rule of utils.py:154-170: cos(elevation) and azimuth uniform."""
    cos_elev = torch.rand(n_views, generator=generator) * 2 - 1
    elev = torch.acos(cos_elev) * 180 / torch.pi - 90
    azim = torch.rand(n_views, generator=generator) * 360 - 180
    return look_at_view_transform(dist, elev, azim, at=at, device=device)

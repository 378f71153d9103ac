"""FoVPerspectiveCameras + look_at_view_transform (utils.py:7, 9, 149, 161-168; first_approach.py:106)."""
from __future__ import annotations

import torch

from st3d import cameras as _cam


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees: bool = True, eye=None, at=((0, 0, 0),),
                           up=((0, 1, 0),), device="cpu"):
    if eye is not None:
        raise NotImplementedError("look_at_view_transform(eye=...) is not implemented; pass dist / elev / azim")
    return _cam.look_at_view_transform(dist, elev, azim, degrees, at, up, device)


def _per_camera(value, n, name):
    t = torch.as_tensor(value, dtype=torch.float32).reshape(-1)
    if t.numel() == 1:
        return t.expand(n).clone()
    if t.numel() != n:
        raise ValueError(f"{name} has {t.numel()} entries for {n} cameras")
    return t.clone()


class FoVPerspectiveCameras:
    """N cameras with R (N,3,3), T (N,3) in the row-vector convention X_view = X_world R + T and a
    field-of-view projection (defaults znear=1, zfar=100, aspect=1, fov=60 degrees)."""

    def __init__(self, znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees: bool = True, R=None, T=None,
                 K=None, device="cpu"):
        if K is not None:
            raise NotImplementedError("explicit projection matrices K are not implemented")
        R = torch.eye(3)[None] if R is None else torch.as_tensor(R, dtype=torch.float32)
        T = torch.zeros(1, 3) if T is None else torch.as_tensor(T, dtype=torch.float32)
        R, T = R.reshape(-1, 3, 3), T.reshape(-1, 3)
        n = max(R.shape[0], T.shape[0])
        if R.shape[0] not in (1, n) or T.shape[0] not in (1, n):
            raise ValueError("R and T describe different numbers of cameras")
        self.device = torch.device(device)
        self.R = R.expand(n, 3, 3).to(self.device)
        self.T = T.expand(n, 3).to(self.device)
        self.znear, self.zfar = _per_camera(znear, n, "znear"), _per_camera(zfar, n, "zfar")
        self.aspect_ratio, self.fov = _per_camera(aspect_ratio, n, "aspect_ratio"), _per_camera(fov, n, "fov")
        self.degrees = degrees

    def __len__(self):
        return self.R.shape[0]

    def _select(self, idx):
        c = FoVPerspectiveCameras.__new__(FoVPerspectiveCameras)
        c.device, c.degrees = self.device, self.degrees
        c.R, c.T = self.R[idx], self.T[idx]
        for k in ("znear", "zfar", "aspect_ratio", "fov"):
            setattr(c, k, getattr(self, k)[idx])
        return c

    def __getitem__(self, index):
        if isinstance(index, int):
            if index >= len(self) or index < -len(self):
                raise IndexError(f"camera index {index} out of range for {len(self)} cameras")
            index = [index % len(self)]
        elif isinstance(index, slice):
            index = list(range(len(self)))[index]
        elif torch.is_tensor(index):
            index = index.tolist()
        return self._select(list(index))

    @staticmethod
    def join(cameras):
        """One batch from a list of camera objects (what the reference's loops hand to render_meshes)."""
        cameras = list(cameras)
        if not cameras:
            raise ValueError("empty camera list")
        c = FoVPerspectiveCameras.__new__(FoVPerspectiveCameras)
        c.device, c.degrees = cameras[0].device, cameras[0].degrees
        c.R = torch.cat([x.R for x in cameras], dim=0)
        c.T = torch.cat([x.T for x in cameras], dim=0)
        for k in ("znear", "zfar", "aspect_ratio", "fov"):
            setattr(c, k, torch.cat([getattr(x, k) for x in cameras], dim=0))
        return c

    def to(self, device):
        c = self._select(list(range(len(self))))
        c.device = torch.device(device)
        c.R, c.T = c.R.to(device), c.T.to(device)
        return c

    def clone(self):
        return self._select(list(range(len(self))))

    def is_perspective(self):
        return True

    def in_ndc(self):
        return True

    def get_camera_center(self):
        return -torch.einsum("ni,nji->nj", self.T, self.R)          # -T R^T

    def uniform_intrinsics(self):
        """(fov_degrees, aspect, znear, zfar) shared by every camera of the batch."""
        vals = []
        for k in ("fov", "aspect_ratio", "znear", "zfar"):
            t = getattr(self, k)
            if not bool((t == t[0]).all()):
                raise NotImplementedError(f"cameras of one batch must share {k}")
            vals.append(float(t[0]))
        if not self.degrees:
            vals[0] = vals[0] * 180.0 / torch.pi
        return tuple(vals)


__all__ = ["FoVPerspectiveCameras", "look_at_view_transform"]

"""Row-vector 4x4 transforms (the subset used at utils.py:142)."""
from __future__ import annotations

from st3d.cameras import rotate_axis_angle_matrix


class RotateAxisAngle:
    def __init__(self, angle, axis: str = "X", degrees: bool = True, dtype=None, device="cpu"):
        self._matrix = rotate_axis_angle_matrix(angle, axis.upper(), degrees, device)
        if dtype is not None:
            self._matrix = self._matrix.to(dtype)

    def get_matrix(self):
        return self._matrix

    def to(self, device):
        self._matrix = self._matrix.to(device)
        return self


__all__ = ["RotateAxisAngle"]

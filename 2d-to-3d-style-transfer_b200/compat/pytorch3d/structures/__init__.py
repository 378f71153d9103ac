"""Meshes container (the subset of pytorch3d.structures.Meshes the reference touches: utils.py:101-111,
175-179, 209)."""
from __future__ import annotations

from typing import Optional

import torch


def _as_list(x, last_dim, name):
    if isinstance(x, (list, tuple)):
        out = list(x)
    elif torch.is_tensor(x) and x.dim() == 3:
        out = [x[i] for i in range(x.shape[0])] if x.shape[0] != 1 else [x[0]]
    else:
        raise ValueError(f"{name} must be a list of (*, {last_dim}) tensors or a (N, *, {last_dim}) tensor")
    for t in out:
        if not torch.is_tensor(t) or t.dim() != 2 or t.shape[1] != last_dim:
            raise ValueError(f"{name}: every entry must have shape (*, {last_dim})")
    return out


class Meshes:
    def __init__(self, verts=None, faces=None, textures=None):
        self._verts_list = _as_list(verts, 3, "verts")
        self._faces_list = _as_list(faces, 3, "faces")
        if len(self._verts_list) != len(self._faces_list):
            raise ValueError("verts and faces describe different numbers of meshes")
        self.textures = textures
        self.device = self._verts_list[0].device if self._verts_list else torch.device("cpu")

    def __len__(self):
        return len(self._verts_list)

    def __getitem__(self, index):
        """Mesh `index` (an int) as a one-mesh Meshes that shares the vertex / face / texture tensors."""
        if not isinstance(index, int):
            raise IndexError("Meshes supports integer indexing")
        if not -len(self) <= index < len(self):
            raise IndexError(f"mesh index {index} out of range for {len(self)} meshes")
        index %= len(self)
        tex = self.textures[index] if self.textures is not None else None
        return Meshes([self._verts_list[index]], [self._faces_list[index]], tex)

    def verts_list(self):
        return self._verts_list

    def faces_list(self):
        return self._faces_list

    def _single(self, what):
        if len(self) != 1:
            raise NotImplementedError(f"{what}: this compatibility layer batches VIEWS of one mesh, not meshes")

    def verts_packed(self):
        # for a one-mesh batch this is the very tensor that was passed in, so a leaf stays a leaf
        return self._verts_list[0] if len(self) == 1 else torch.cat(self._verts_list, dim=0)

    def faces_packed(self):
        if len(self) == 1:
            return self._faces_list[0]
        off, out = 0, []
        for v, f in zip(self._verts_list, self._faces_list):
            out.append(f + off)
            off += v.shape[0]
        return torch.cat(out, dim=0)

    def verts_padded(self):
        self._single("verts_padded")
        return self._verts_list[0][None]

    def faces_padded(self):
        self._single("faces_padded")
        return self._faces_list[0][None]

    def num_verts_per_mesh(self):
        return torch.tensor([v.shape[0] for v in self._verts_list], device=self.device)

    def num_faces_per_mesh(self):
        return torch.tensor([f.shape[0] for f in self._faces_list], device=self.device)

    def edges_packed(self):
        """Unique undirected edges (E,2), sorted lexicographically."""
        f = self.faces_packed()
        e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], dim=0)
        e = torch.sort(e, dim=1).values
        return torch.unique(e, dim=0)

    def clone(self):
        tex = self.textures.clone() if self.textures is not None else None
        return Meshes([v.clone() for v in self._verts_list], [f.clone() for f in self._faces_list], tex)

    def detach(self):
        tex = self.textures.detach() if self.textures is not None else None
        return Meshes([v.detach() for v in self._verts_list], [f.detach() for f in self._faces_list], tex)

    def to(self, device):
        tex = self.textures.to(device) if self.textures is not None else None
        return Meshes([v.to(device) for v in self._verts_list], [f.to(device) for f in self._faces_list], tex)

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")


__all__ = ["Meshes"]

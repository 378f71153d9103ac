"""Mesh regularisers used for the `mesh` / `both` optimisation targets (losses.py:85-87): restated from
their published definitions as plain differentiable torch ops.  They are
view-independent, O(V + F) and outside the render/loss hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _single(meshes):
    if len(meshes) != 1:
        raise NotImplementedError("one mesh per batch")
    return meshes.verts_packed(), meshes.faces_packed().long()


def mesh_edge_loss(meshes, target_length: float = 0.0):
    """mean over unique edges of (|v0 - v1| - target_length)^2."""
    verts, _ = _single(meshes)
    e = meshes.edges_packed().long()
    if e.numel() == 0:
        return verts.sum() * 0.0
    length = (verts[e[:, 0]] - verts[e[:, 1]]).norm(dim=1, p=2)
    return ((length - target_length) ** 2).mean()


def mesh_laplacian_smoothing(meshes, method: str = "uniform"):
    """mean_i |(L V)_i| with the uniform graph Laplacian L = D^-1 A - I (built without gradient)."""
    if method != "uniform":
        raise NotImplementedError("only method='uniform' (the default the reference uses) is implemented")
    verts, _ = _single(meshes)
    e = meshes.edges_packed().long()
    V = verts.shape[0]
    if e.numel() == 0:
        return verts.sum() * 0.0
    with torch.no_grad():
        src = torch.cat([e[:, 0], e[:, 1]])
        dst = torch.cat([e[:, 1], e[:, 0]])
        deg = torch.zeros(V, device=verts.device, dtype=verts.dtype).index_add_(0, src, torch.ones_like(src, dtype=verts.dtype))
        inv = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg))
    neigh = torch.zeros_like(verts).index_add(0, src, verts[dst])
    lap = neigh * inv[:, None] - verts
    return lap.norm(dim=1).mean()


def mesh_normal_consistency(meshes):
    """mean over pairs of faces sharing an edge of 1 - cos(n_a, n_b)."""
    verts, faces = _single(meshes)
    Fn, V = faces.shape[0], verts.shape[0]
    if Fn == 0:
        return verts.sum() * 0.0
    # half-edges: edge opposite to corner c of face f
    e = torch.cat([faces[:, [1, 2]], faces[:, [2, 0]], faces[:, [0, 1]]], dim=0)
    opp = torch.cat([faces[:, 0], faces[:, 1], faces[:, 2]], dim=0)
    es = torch.sort(e, dim=1).values
    key = es[:, 0] * (V + 1) + es[:, 1]
    order = torch.argsort(key, stable=True)
    key, es, opp = key[order], es[order], opp[order]
    _, inverse, counts = torch.unique_consecutive(key, return_inverse=True, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(key.numel(), device=key.device) - starts[inverse]      # position inside its edge group
    pa, pb = [], []
    for d in range(1, int(counts.max()) if counts.numel() else 0):
        # pair element i of a group with element i + d of the same group
        ok = rank + d < counts[inverse]
        idx = torch.nonzero(ok).flatten()
        pa.append(idx)
        pb.append(idx + d)
    if not pa or sum(p.numel() for p in pa) == 0:
        return verts.sum() * 0.0
    pa, pb = torch.cat(pa), torch.cat(pb)
    v0, v1 = verts[es[pa, 0]], verts[es[pa, 1]]
    n0 = torch.cross(v1 - v0, verts[opp[pa]] - v0, dim=1)
    n1 = -torch.cross(v1 - v0, verts[opp[pb]] - v0, dim=1)
    return (1.0 - F.cosine_similarity(n0, n1, dim=1)).mean()


__all__ = ["mesh_edge_loss", "mesh_laplacian_smoothing", "mesh_normal_consistency"]

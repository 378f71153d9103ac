"""Mesh regularisers used for the `mesh` / `both` optimisation targets (losses.py:85-87): restated from
their published definitions as plain differentiable torch ops on (verts, faces).
They are view-independent, O(V + F) and outside the render/loss hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def unique_edges(faces):
    """Unique undirected edges (E,2) of a triangle list, sorted lexicographically."""
    f = faces.long()
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], dim=0)
    return torch.unique(torch.sort(e, dim=1).values, dim=0)


def edge_loss(verts, faces, target_length: float = 0.0, edges=None):
    """mean over unique edges of (|v0 - v1| - target_length)^2."""
    e = unique_edges(faces) if edges is None else edges
    if e.numel() == 0:
        return verts.sum() * 0.0
    length = (verts[e[:, 0]] - verts[e[:, 1]]).norm(dim=1, p=2)
    return ((length - target_length) ** 2).mean()


def laplacian_smoothing(verts, faces, edges=None):
    """mean_i |(L V)_i| with the uniform graph Laplacian L = D^-1 A - I (built without gradient)."""
    e = unique_edges(faces) if edges is None else edges
    V = verts.shape[0]
    if e.numel() == 0:
        return verts.sum() * 0.0
    with torch.no_grad():
        src = torch.cat([e[:, 0], e[:, 1]])
        dst = torch.cat([e[:, 1], e[:, 0]])
        deg = torch.zeros(V, device=verts.device, dtype=verts.dtype).index_add_(0, src, torch.ones_like(src, dtype=verts.dtype))
        inv = torch.where(deg > 0, 1.0 / deg, torch.zeros_like(deg))
    neigh = torch.zeros_like(verts).index_add(0, src, verts[dst])
    return (neigh * inv[:, None] - verts).norm(dim=1).mean()


def normal_consistency(verts, faces):
    """mean over pairs of faces sharing an edge of 1 - cos(n_a, n_b)."""
    faces = faces.long()
    Fn, V = faces.shape[0], verts.shape[0]
    if Fn == 0:
        return verts.sum() * 0.0
    e = torch.cat([faces[:, [1, 2]], faces[:, [2, 0]], faces[:, [0, 1]]], dim=0)   # edge opposite to corner c
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
        idx = torch.nonzero(rank + d < counts[inverse]).flatten()               # pair element i with i + d of its group
        pa.append(idx)
        pb.append(idx + d)
    if not pa or sum(p.numel() for p in pa) == 0:
        return verts.sum() * 0.0
    pa, pb = torch.cat(pa), torch.cat(pb)
    v0, v1 = verts[es[pa, 0]], verts[es[pa, 1]]
    n0 = torch.cross(v1 - v0, verts[opp[pa]] - v0, dim=1)
    n1 = -torch.cross(v1 - v0, verts[opp[pb]] - v0, dim=1)
    return (1.0 - F.cosine_similarity(n0, n1, dim=1)).mean()

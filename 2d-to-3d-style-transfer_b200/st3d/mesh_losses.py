"""Mesh regularisers of the `mesh` / `both` optimisation targets (losses.py:85-87, 113-115 of the reference:
pytorch3d.loss.mesh_edge_loss, mesh_laplacian_smoothing(method='uniform'), mesh_normal_consistency; SURVEY.md
Appendix A.7, section 8 row f3).

CUDA float32 vertices: ONE forward and ONE backward launch of libst3d (`st3d_mesh_regularizers_forward / _backward`,
csrc/mesh_reg.cu) over topology tables that depend on the faces only and are built once per face list -- the torch
formulation below costs ~80 launches, three sorts and a host read per call.  The library is required for CUDA tensors.

`*_torch`: the same definitions as plain differentiable torch ops.  They build the tables' ground truth in the tests
(float64 known answers, tests/test_oracle_pinned.py) and serve tensors the library does not take (CPU, float64): the
regularisers are view-independent O(V + F) glue, not the render/loss hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


def unique_edges(faces):
    """Unique undirected edges (E,2) of a triangle list, sorted lexicographically."""
    f = faces.long()
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], dim=0)
    return torch.unique(torch.sort(e, dim=1).values, dim=0)


def _face_pairs(faces, V):
    """(P,4) int64 rows (v0, v1, a, b): one per unordered pair of faces sharing the edge (v0, v1), a and b the vertices
    opposite to it.  An edge shared by k faces yields k (k - 1) / 2 rows (upstream pairs every two of them)."""
    faces = faces.long()
    if faces.shape[0] == 0:
        return faces.new_zeros((0, 4))
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
    if not pa:
        return faces.new_zeros((0, 4))
    pa, pb = torch.cat(pa), torch.cat(pb)
    return torch.stack([es[pa, 0], es[pa, 1], opp[pa], opp[pb]], dim=1)


@dataclass
class MeshTopology:
    """Tables of `st3d_mesh_reg_args` (include/st3d.h), int32 on the device of the faces."""
    num_verts: int
    edges: torch.Tensor      # (E,2)
    adj_ptr: torch.Tensor    # (V+1,)
    adj_idx: torch.Tensor    # (2E,)
    pairs: torch.Tensor      # (P,4)


def topology(faces: torch.Tensor, num_verts: int) -> MeshTopology:
    """Built with torch ops on the device of `faces` (sorts + one host read): once per face list, see `_cached`."""
    V = int(num_verts)
    e = unique_edges(faces)
    src = torch.cat([e[:, 0], e[:, 1]])
    dst = torch.cat([e[:, 1], e[:, 0]])
    order = torch.argsort(src, stable=True)
    deg = torch.bincount(src, minlength=V)[:V] if src.numel() else torch.zeros(V, dtype=torch.long, device=faces.device)
    adj_ptr = torch.cat([deg.new_zeros(1), torch.cumsum(deg, 0)])
    i32 = dict(dtype=torch.int32)
    return MeshTopology(V, e.to(**i32).contiguous(), adj_ptr.to(**i32).contiguous(), dst[order].to(**i32).contiguous(),
                        _face_pairs(faces, V).to(**i32).contiguous())


# The tables are keyed on the STORAGE of the face tensor (the compat Meshes hands out `.detach()` views of one tensor:
# new objects, same storage, shared version counter).  Each entry keeps its face tensor alive, so the address cannot be
# handed to another tensor while the entry exists, and an in-place edit shows in `_version`.
_CACHE: dict = {}
_CACHE_MAX = 8


def _cached(faces: torch.Tensor, num_verts: int) -> MeshTopology:
    key = (faces.device, faces.data_ptr(), tuple(faces.shape), tuple(faces.stride()), faces.dtype, int(num_verts))
    hit = _CACHE.get(key)
    if hit is not None and hit[1] == faces._version:
        return hit[2]
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    topo = topology(faces, num_verts)
    _CACHE[key] = (faces, faces._version, topo)
    return topo


class _MeshRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, topo, target_length, which):
        from . import ops
        losses, ctx.state = ops.mesh_regularizers_forward(verts.detach(), topo, target_length, which)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        from . import ops
        return ops.mesh_regularizers_backward(ctx.state, grad_losses.contiguous()), None, None, None


def _fused(verts) -> bool:
    return verts.is_cuda and verts.dtype == torch.float32


def regularizers(verts, faces, target_length: float = 0.0, which: int = 7, topo: MeshTopology = None):
    """(3,) tensor [edge, laplacian, normal consistency] from one launch; entries not in `which` are 0."""
    if not _fused(verts):
        raise ValueError("regularizers: CUDA float32 vertices required (libst3d has no CPU path); "
                         "use the *_torch functions for other tensors")
    return _MeshRegFn.apply(verts, topo if topo is not None else _cached(faces, verts.shape[0]), float(target_length),
                            int(which))


def edge_loss(verts, faces, target_length: float = 0.0, edges=None):
    """mean over unique edges of (|v0 - v1| - target_length)^2."""
    if _fused(verts):
        return regularizers(verts, faces, target_length, 1)[0]
    return edge_loss_torch(verts, faces, target_length, edges)


def laplacian_smoothing(verts, faces, edges=None):
    """mean_i |(L V)_i| with the uniform graph Laplacian L = D^-1 A - I (built without gradient)."""
    if _fused(verts):
        return regularizers(verts, faces, 0.0, 2)[1]
    return laplacian_smoothing_torch(verts, faces, edges)


def normal_consistency(verts, faces):
    """mean over pairs of faces sharing an edge of 1 - cos(n_a, n_b)."""
    if _fused(verts):
        return regularizers(verts, faces, 0.0, 4)[2]
    return normal_consistency_torch(verts, faces)


# ---- the definitions as torch ops ------------------------------------------------------------------------------------
def edge_loss_torch(verts, faces, target_length: float = 0.0, edges=None):
    e = unique_edges(faces) if edges is None else edges
    if e.numel() == 0:
        return verts.sum() * 0.0
    length = (verts[e[:, 0]] - verts[e[:, 1]]).norm(dim=1, p=2)
    return ((length - target_length) ** 2).mean()


def laplacian_smoothing_torch(verts, faces, edges=None):
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


def normal_consistency_torch(verts, faces):
    q = _face_pairs(faces, verts.shape[0])
    if q.shape[0] == 0:
        return verts.sum() * 0.0
    v0, v1 = verts[q[:, 0]], verts[q[:, 1]]
    n0 = torch.cross(v1 - v0, verts[q[:, 2]] - v0, dim=1)
    n1 = -torch.cross(v1 - v0, verts[q[:, 3]] - v0, dim=1)
    return (1.0 - F.cosine_similarity(n0, n1, dim=1)).mean()

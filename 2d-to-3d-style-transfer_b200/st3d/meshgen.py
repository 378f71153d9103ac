"""Synthetic dense meshes for the scaling workload (BASELINE.json configs[4]: "synthetic subdivided mesh ~1M faces").

Loop subdivision without smoothing: every triangle splits into four at its edge midpoints; applied to the UV
topology as well, so the subdivided mesh keeps a consistent UV map (midpoint-interpolated UVs).  The cow (5856
faces) subdivided four times has 5856 * 4^4 = 1 499 136 faces.  Host-side tensor code, runs once at setup."""
from __future__ import annotations

import torch


def subdivide(points: torch.Tensor, faces: torch.Tensor):
    """(P,D) points + (F,3) int64 faces -> (P + E, D) points, (4F,3) faces; new points are the edge midpoints."""
    faces = faces.long()
    edges = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], dim=0)
    uniq, inverse = torch.unique(torch.sort(edges, dim=1).values, dim=0, return_inverse=True)
    mid = 0.5 * (points[uniq[:, 0]] + points[uniq[:, 1]])
    P, Fc = points.shape[0], faces.shape[0]
    m01, m12, m20 = P + inverse[:Fc], P + inverse[Fc:2 * Fc], P + inverse[2 * Fc:]
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    new_faces = torch.cat([torch.stack([a, m01, m20], 1), torch.stack([m01, b, m12], 1),
                           torch.stack([m20, m12, c], 1), torch.stack([m01, m12, m20], 1)], dim=0)
    return torch.cat([points, mid], dim=0), new_faces


def subdivided_uv_mesh(verts, faces, verts_uvs, faces_uvs, levels: int):
    """`levels` rounds of 1-to-4 subdivision of a UV-mapped mesh; face i of the geometry stays face i of the UV
    topology (both are split in the same order)."""
    for _ in range(int(levels)):
        verts, faces = subdivide(verts, faces)
        verts_uvs, faces_uvs = subdivide(verts_uvs, faces_uvs)
    return verts, faces, verts_uvs, faces_uvs

"""`pytorch3d.loss` names used by the reference (losses.py:3, 85-87), taking a Meshes object."""
from __future__ import annotations

from st3d import mesh_losses as _ml


def _single(meshes):
    if len(meshes) != 1:
        raise NotImplementedError("one mesh per batch")
    return meshes.verts_packed(), meshes.faces_packed()


def mesh_edge_loss(meshes, target_length: float = 0.0):
    verts, faces = _single(meshes)
    return _ml.edge_loss(verts, faces, target_length)


def mesh_laplacian_smoothing(meshes, method: str = "uniform"):
    if method != "uniform":
        raise NotImplementedError("only method='uniform' (the default the reference uses) is implemented")
    verts, faces = _single(meshes)
    return _ml.laplacian_smoothing(verts, faces)


def mesh_normal_consistency(meshes):
    verts, faces = _single(meshes)
    return _ml.normal_consistency(verts, faces)


__all__ = ["mesh_edge_loss", "mesh_laplacian_smoothing", "mesh_normal_consistency"]

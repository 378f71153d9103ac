"""oracle/loss_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Port of the loss half of the reference hot path, written so it can travel to the GPU box
(where /root/reference does not exist).  It is pinned against the LIVE reference in
tests/test_oracle_pinned.py and tests/golden/make_golden.py (run in the build container, where
/root/reference/style_transfer.py and losses.py import with a stubbed `pytorch3d.loss`).

Follows: style_transfer.py:10-27 (get_features), :31-35 (gram_matrix), losses.py:12-44
(compute_perceptual_loss), :68-98 (compute_first_approach_loss), :101-126
(compute_second_approach_loss); mesh regularisers per SURVEY.md Appendix A.7 (pytorch3d.loss,
third-party, absent).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# VGG-19 `.features` taps (style_transfer.py:12-19).  ReLUs are inplace, so every tapped tensor
# ends up post-ReLU.
TAPS = {"0": "conv1_1", "5": "conv2_1", "10": "conv3_1", "19": "conv4_1", "21": "conv4_2", "28": "conv5_1"}
STYLE_LAYERS = ("conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv5_1")
CONTENT_LAYER = "conv4_2"


def get_features(image, model, layers=None):
    layers = TAPS if layers is None else layers
    out, x = {}, image
    for name, module in model._modules.items():
        x = module(x)
        if name in layers:
            out[layers[name]] = x
    return out


def gram_matrix(t):
    b, c, h, w = t.shape
    flat = t.reshape(b, c, h * w)
    return flat @ flat.transpose(1, 2)


def style_layer_loss(feature, target_gram):
    """losses.py:35-39: mean((G - Gs)^2) / (C^2 * H^2)  (H twice -- not C^2*H*W)."""
    g = gram_matrix(feature)
    return ((g - target_gram) ** 2).mean() / (feature.shape[1] ** 2 * feature.shape[2] ** 2)


def perceptual_loss(current, content, style, model, style_weight=1e6, content_weight=1.0):
    assert current.shape[0] == content.shape[0] == style.shape[0]
    content_feat = get_features(content, model)[CONTENT_LAYER]
    style_feats = get_features(style, model)
    grams = {k: gram_matrix(v) for k, v in style_feats.items() if k != CONTENT_LAYER}
    cur = get_features(current, model)
    c_loss = ((cur[CONTENT_LAYER] - content_feat) ** 2).mean()
    s_loss = 0
    for k, g in grams.items():
        s_loss = s_loss + style_layer_loss(cur[k], g)
    return content_weight * c_loss + style_weight * s_loss


# ---- Appendix A.7 regularisers (single mesh) ---------------------------------------------------
def _unique_edges(faces):
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], dim=0)
    e = torch.sort(e, dim=1).values
    return torch.unique(e, dim=0)


def mesh_edge_loss(verts, faces, target_length=0.0):
    e = _unique_edges(faces)
    length = (verts[e[:, 0]] - verts[e[:, 1]]).norm(dim=1, p=2)
    return ((length - target_length) ** 2).mean()


def mesh_laplacian_smoothing(verts, faces):
    """method='uniform': L_ij = 1/deg(i), L_ii = -1; loss = mean_i ||(L V)_i||."""
    e = _unique_edges(faces)
    V = verts.shape[0]
    with torch.no_grad():
        idx = torch.cat([e, e.flip(1)], dim=0).t()
        A = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1], dtype=verts.dtype), (V, V))
        deg = torch.sparse.sum(A, dim=1).to_dense()
        inv = torch.where(deg > 0, 1.0 / deg, deg)
    lv = torch.sparse.mm(A, verts) * inv[:, None] - verts
    return lv.norm(dim=1).mean()


def mesh_normal_consistency(verts, faces):
    """1 - cos(n_a, n_b) over all pairs of faces sharing an edge, averaged."""
    Fn = faces.shape[0]
    e = torch.cat([faces[:, [1, 2]], faces[:, [2, 0]], faces[:, [0, 1]]], dim=0)  # edge opposite v0,v1,v2
    opp = torch.cat([faces[:, 0], faces[:, 1], faces[:, 2]], dim=0)
    fid = torch.arange(Fn).repeat(3)
    es = torch.sort(e, dim=1).values
    key = es[:, 0] * (verts.shape[0] + 1) + es[:, 1]
    order = torch.argsort(key, stable=True)
    key, es, opp, fid = key[order], es[order], opp[order], fid[order]
    # group equal keys; emit all unordered pairs inside each group
    _, counts = torch.unique_consecutive(key, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    pa, pb = [], []
    for s, c in zip(starts.tolist(), counts.tolist()):
        for i in range(c):
            for j in range(i + 1, c):
                pa.append(s + i)
                pb.append(s + j)
    if not pa:
        return verts.sum() * 0.0
    pa, pb = torch.tensor(pa), torch.tensor(pb)
    v0, v1 = verts[es[pa, 0]], verts[es[pa, 1]]
    a, b = verts[opp[pa]], verts[opp[pb]]
    n0 = torch.cross(v1 - v0, a - v0, dim=1)
    n1 = -torch.cross(v1 - v0, b - v0, dim=1)
    return (1.0 - F.cosine_similarity(n0, n1, dim=1)).mean()


def first_approach_loss(rendered, masks, target, verts, target_verts, faces, weights, opt_type):
    r, t = rendered * masks, target * masks
    if opt_type == "texture":
        return F.mse_loss(r, t)
    loss = weights["main_loss_weight"] * F.mse_loss(r, t)
    return loss + _regularisers(verts, target_verts, faces, weights)


def second_approach_loss(current, content, style, model, style_weight, content_weight, verts, target_verts,
                         faces, weights, opt_type):
    p = perceptual_loss(current, content, style, model, style_weight, content_weight)
    if opt_type == "texture":
        return p
    return weights["main_loss_weight"] * p + _regularisers(verts, target_verts, faces, weights)


def _regularisers(verts, target_verts, faces, weights):
    loss = weights["mesh_verts_weight"] * F.mse_loss(verts, target_verts)
    loss = loss + weights["mesh_edge_loss_weight"] * mesh_edge_loss(verts, faces)
    loss = loss + weights["mesh_laplacian_smoothing_weight"] * mesh_laplacian_smoothing(verts, faces)
    loss = loss + weights["mesh_normal_consistency_weight"] * mesh_normal_consistency(verts, faces)
    return loss

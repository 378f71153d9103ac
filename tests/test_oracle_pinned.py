"""CPU tests that pin the oracle: loss half against the live reference's golden outputs
(tests/golden/loss_golden.npz, produced by importing /root/reference), render half against
hand-checkable cases and frozen oracle renders."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from oracle import render_oracle as ro

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402


def test_gram_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    _, _, _, feat, _ = mg.loss_inputs()
    np.testing.assert_allclose(lo.gram_matrix(feat).numpy(), g["gram"], rtol=1e-6, atol=1e-6)


def test_perceptual_loss_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    vgg = mg.seeded_vgg()
    cur, con, sty, _, _ = mg.loss_inputs()
    cur.requires_grad_(True)
    loss = lo.perceptual_loss(cur, con, sty, vgg, 1e6, 1.0)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["perceptual"], rtol=2e-5)
    np.testing.assert_allclose(cur.grad.numpy(), g["perceptual_grad"], rtol=1e-3, atol=1e-7)


def test_first_approach_loss_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    cur, con, _, _, masks = mg.loss_inputs()
    cur.requires_grad_(True)
    loss = lo.first_approach_loss(cur, masks, con, None, None, None, {}, "texture")
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["first"], rtol=1e-6)
    np.testing.assert_allclose(cur.grad.numpy(), g["first_grad"], rtol=1e-5, atol=1e-9)


@pytest.mark.skipif(not os.path.exists("/root/reference/losses.py"), reason="live reference only in the build container")
def test_loss_oracle_against_live_reference():
    ref_st, ref_losses = mg.import_reference()
    vgg = mg.seeded_vgg()
    g = torch.Generator().manual_seed(7)
    cur, con = torch.rand(1, 3, 48, 40, generator=g), torch.rand(1, 3, 48, 40, generator=g)
    sty = torch.rand(1, 3, 48, 40, generator=g)
    a = ref_losses.compute_perceptual_loss(cur, con, sty, vgg, style_weight=1e6, content_weight=1)
    b = lo.perceptual_loss(cur, con, sty, vgg, 1e6, 1.0)
    assert abs(a.item() - b.item()) <= 1e-5 * abs(a.item())
    f = torch.rand(3, 16, 6, 9, generator=g)
    assert torch.equal(ref_st.gram_matrix(f), lo.gram_matrix(f))


# ---- hand-checkable rasterization cases ------------------------------------------
def _raster(fv, S=8, **kw):
    fv = torch.tensor(fv, dtype=torch.float32).reshape(-1, 3, 3)
    first = torch.zeros(1, dtype=torch.int64)
    num = torch.tensor([fv.shape[0]])
    return ro.rasterize_naive(fv, first, num, S, **kw)


def test_pixel_axes_and_flip():
    # a small triangle in the NDC (+x, +y) quadrant must land top-LEFT in the image (+X left, +Y up)
    p2f, zbuf, bary, dists = _raster([[0.2, 0.2, 2.0], [0.9, 0.2, 2.0], [0.2, 0.9, 2.0]], S=8)
    ys, xs = torch.nonzero(p2f[0, ..., 0] >= 0, as_tuple=True)
    assert len(ys) > 0 and ys.max() < 4 and xs.max() < 4
    assert torch.allclose(zbuf[p2f >= 0], torch.tensor(2.0), atol=1e-6)
    assert (dists[p2f >= 0] < 0).all()
    assert torch.allclose(bary[p2f[..., 0] >= 0].sum(-1), torch.tensor(1.0), atol=1e-5)


def test_centre_on_shared_edge_is_a_hole():
    # two triangles sharing the vertical edge x = 0.125 (pixel-centre column for S=8: -1+(2i+1)/8)
    x = 0.125
    tris = [[x, -1, 1.0], [x, 1, 1.0], [1, 0, 1.0],
            [x, 1, 1.0], [x, -1, 1.0], [-1, 0, 1.0]]
    p2f, *_ = _raster(tris, S=8)
    xs_ndc, _ = ro.pixel_ndc_grid(8, 8)
    col = int(torch.nonzero(xs_ndc == x)[0])
    assert (p2f[0, :, col, 0] == -1).all()          # strict > 0 inside test => neither face owns the edge
    assert (p2f[0, 4, :, 0] >= 0).sum() >= 5        # but neighbours are covered


def test_backfaces_drawn_and_nearest_wins_with_index_tiebreak():
    front = [[-0.9, -0.9, 3.0], [0.0, 0.9, 3.0], [0.9, -0.9, 3.0]]   # edge(v2;v0,v1) > 0: front-facing
    back = [[-0.9, -0.9, 2.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]]    # area < 0 (back-facing), nearer
    p2f, zbuf, *_ = _raster(front + back, S=16)
    cov = p2f[0, ..., 0] >= 0
    assert cov.any() and (p2f[0, ..., 0][cov] == 1).all() and torch.allclose(zbuf[0, ..., 0][cov], torch.tensor(2.0))
    p2f_c, *_ = _raster(front + back, S=16, cull_backfaces=True)
    assert set(p2f_c[0, ..., 0][cov].tolist()) == {0}
    p2f_t, *_ = _raster(front + front, S=16)        # exact z tie -> lower face index
    assert (p2f_t[0, ..., 0][cov] == 0).all()
    p2f_k, zk, *_ = _raster(front + back, S=16, faces_per_pixel=3)
    assert (p2f_k[0, ..., 0][cov] == 1).all() and (p2f_k[0, ..., 1][cov] == 0).all() and (p2f_k[0, ..., 2] == -1).all()


def test_degenerate_and_behind_camera_faces_skipped():
    deg = [[0.0, 0.0, 1.0], [0.5, 0.5, 1.0], [1.0, 1.0, 1.0]]
    behind = [[-0.9, -0.9, -1.0], [0.9, -0.9, -1.0], [0.0, 0.9, -1.0]]
    p2f, *_ = _raster(deg + behind, S=8)
    assert (p2f == -1).all()


def test_fragment_recompute_matches_exact(cow):
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(3))
    rgba, fr = ro.render_views(cow["verts"], cow["faces"], R, T, 48, texture=cow["texture"],
                               verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], return_fragments=True, nthreads=4)
    m = fr["pix_to_face"] >= 0
    assert 0.1 < m.float().mean() < 0.6
    assert torch.allclose(fr["zbuf"], fr["zbuf_exact"], atol=1e-5)
    assert torch.allclose(fr["bary"], fr["bary_exact"], atol=1e-4)
    assert torch.allclose(fr["dists"], fr["dists_exact"], atol=1e-8)
    assert (fr["dists_exact"][m] <= 0).all()
    img, mask = ro.images_and_masks(rgba)
    assert torch.equal(mask[:, 0] > 0, m[..., 0])
    assert torch.equal(img.permute(0, 2, 3, 1)[~m[..., 0]], torch.ones_like(img.permute(0, 2, 3, 1)[~m[..., 0]]))


def test_render_oracle_frozen(golden_dir, cow):
    g = np.load(os.path.join(golden_dir, "render_golden.npz"))
    for sfx, S in (("", 64), ("2", 96)):
        R, T = torch.from_numpy(g["R" + sfx]), torch.from_numpy(g["T" + sfx])
        rgba, fr = ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=cow["texture"],
                                   verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], return_fragments=True, nthreads=4)
        assert np.array_equal(fr["pix_to_face"].numpy().astype(np.int32), g["p2f" + sfx])
        np.testing.assert_allclose(rgba.numpy(), g["rgba" + sfx].astype(np.float32), atol=2e-3)


def test_camera_restatement_basics():
    R, T = ro.look_at_view_transform(2.7, 10.0, 20.0)
    assert torch.allclose(R[0] @ R[0].t(), torch.eye(3), atol=1e-6)
    C = -T[0] @ R[0].t()                       # camera centre (A.1)
    assert abs(C.norm().item() - 2.7) < 1e-5
    Rf, Tf = ro.fixed_cameras(6)
    assert Rf.shape == (6, 3, 3) and torch.allclose(Tf[:, 2], torch.tensor(3.0))
    assert torch.allclose(Rf[0], torch.eye(3), atol=1e-7)   # angle 0 about X
    k00, k11 = ro.fov_scales(60.0)
    assert abs(k00 - 3 ** 0.5) < 1e-6 and k00 == k11


def test_regularisers_simple_mesh():
    verts = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    faces = torch.tensor([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]])
    e = lo.mesh_edge_loss(verts, faces)
    assert abs(e.item() - (3 * 1.0 + 3 * 2.0) / 6) < 1e-6
    assert lo.mesh_laplacian_smoothing(verts, faces).item() > 0
    n = lo.mesh_normal_consistency(verts, faces)
    assert 0.0 < n.item() < 2.0

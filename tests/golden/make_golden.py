"""Generates the committed fixtures under tests/golden/.  Run ONLY in the build container, where
/root/reference exists:   python tests/golden/make_golden.py

1. loss_golden.npz   -- outputs of the LIVE reference (`style_transfer.gram_matrix`,
   `losses.compute_perceptual_loss`, `losses.compute_first_approach_loss`) on seeded inputs.
   The reference's `losses.py` imports `pytorch3d.loss` (absent third-party library); a stub module
   with the three regulariser names is registered in sys.modules so the file imports unmodified.
2. cow_mesh.npz      -- geometry of the reference fixture objects/cow_mesh/cow.obj (+ its texture,
   bilinearly reduced to 256x256 uint8) so GPU-box tests/bench have the mesh BASELINE.json names.
2b. bob_mesh.npz (objects/bob_mesh/bob.obj: quads, fan-triangulated; stands in for the missing bunny.obj of
   BASELINE configs[2]), teapot_mesh.npz (objects/teapot_mesh/teapot.obj: no UVs -> per-vertex colours, configs[3]) and
   styles.npz (imgs/Style_1.jpg at 512x512, Style_3/4/5 at 256x256, uint8, resized the way utils.py:34-44 does:
   PIL bilinear to a square) so the GPU box can run configs 2-4 on the reference's own assets.
3. render_golden.npz -- oracle renders (pix_to_face, rgba) of that mesh; NOT upstream PyTorch3D output
   (parity unpinned at that boundary), they freeze the oracle so regressions in it are caught.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def seeded_vgg():
    import torchvision
    torch.manual_seed(0)
    m = torchvision.models.vgg19(weights=None).features.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def loss_inputs():
    g = torch.Generator().manual_seed(1234)
    cur = torch.rand(2, 3, 32, 32, generator=g)
    con = torch.rand(2, 3, 32, 32, generator=g)
    sty = torch.rand(1, 3, 32, 32, generator=g).repeat(2, 1, 1, 1)
    feat = torch.rand(2, 8, 5, 7, generator=g)
    masks = (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float()
    return cur, con, sty, feat, masks


def import_reference():
    """Imports the LIVE reference modules (style_transfer.py, losses.py) without disturbing same-named
    modules that may already be loaded (this repo ships drop-in modules with those names)."""
    stub = types.ModuleType("pytorch3d.loss")
    for n in ("mesh_edge_loss", "mesh_laplacian_smoothing", "mesh_normal_consistency"):
        setattr(stub, n, lambda *a, **k: 0.0)
    pkg = types.ModuleType("pytorch3d")
    pkg.loss = stub
    names = ("pytorch3d", "pytorch3d.loss", "style_transfer", "losses")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.modules["pytorch3d"], sys.modules["pytorch3d.loss"] = pkg, stub
    sys.path.insert(0, REF)
    try:
        import style_transfer as ref_st
        import losses as ref_losses
    finally:
        sys.path.remove(REF)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return ref_st, ref_losses


def make_loss_golden():
    ref_st, ref_losses = import_reference()
    vgg = seeded_vgg()
    cur, con, sty, feat, masks = loss_inputs()
    gram = ref_st.gram_matrix(feat)
    cur_g = cur.clone().requires_grad_(True)
    loss = ref_losses.compute_perceptual_loss(cur_g, con, sty, vgg, style_weight=1e6, content_weight=1)
    loss.backward()
    r = cur.clone().requires_grad_(True)
    fl = ref_losses.compute_first_approach_loss(r, masks, con, None, None, None, {}, "texture")
    fl.backward()
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"),
                        gram=gram.numpy(), perceptual=loss.detach().numpy(), perceptual_grad=cur_g.grad.numpy(),
                        first=fl.detach().numpy(), first_grad=r.grad.numpy())
    print("loss golden:", float(loss), float(fl))


def make_cow():
    from PIL import Image
    from oracle import render_oracle as ro
    v, f, uv, fuv = ro.read_obj(os.path.join(REF, "objects/cow_mesh/cow.obj"))
    img = Image.open(os.path.join(REF, "objects/cow_mesh/cow_texture.png")).convert("RGB").resize((256, 256), Image.BILINEAR)
    np.savez_compressed(os.path.join(HERE, "cow_mesh.npz"), verts=v.numpy(), faces=f.numpy().astype(np.int32),
                        verts_uvs=uv.numpy(), faces_uvs=fuv.numpy().astype(np.int32), texture=np.asarray(img))
    print("cow:", v.shape, f.shape, uv.shape)


def make_bob_teapot_styles():
    from PIL import Image
    from oracle import render_oracle as ro
    v, f, uv, fuv = ro.read_obj(os.path.join(REF, "objects/bob_mesh/bob.obj"))
    img = Image.open(os.path.join(REF, "objects/bob_mesh/bob.png")).convert("RGB").resize((256, 256), Image.BILINEAR)
    np.savez_compressed(os.path.join(HERE, "bob_mesh.npz"), verts=v.numpy(), faces=f.numpy().astype(np.int32),
                        verts_uvs=uv.numpy(), faces_uvs=fuv.numpy().astype(np.int32), texture=np.asarray(img))
    print("bob:", v.shape, f.shape, uv.shape)
    v, f, uv, fuv = ro.read_obj(os.path.join(REF, "objects/teapot_mesh/teapot.obj"))
    assert uv.numel() == 0
    np.savez_compressed(os.path.join(HERE, "teapot_mesh.npz"), verts=v.numpy(), faces=f.numpy().astype(np.int32))
    print("teapot:", v.shape, f.shape)
    styles = {}
    for name, file, size in (("style_1", "Style_1.jpg", 512), ("style_3", "Style_3.png", 256),
                             ("style_4", "Style_4.jpeg", 256), ("style_5", "Style_5.png", 256)):
        im = Image.open(os.path.join(REF, "imgs", file)).convert("RGB").resize((size, size), Image.BILINEAR)
        styles[name] = np.asarray(im)
    np.savez_compressed(os.path.join(HERE, "styles.npz"), **styles)
    print("styles:", {k: v.shape for k, v in styles.items()})


def make_render_golden():
    from oracle import render_oracle as ro
    d = np.load(os.path.join(HERE, "cow_mesh.npz"))
    v, f = torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long()
    uv, fuv = torch.from_numpy(d["verts_uvs"]), torch.from_numpy(d["faces_uvs"]).long()
    tex = torch.from_numpy(d["texture"]).float() / 255.0
    R, T = ro.fixed_cameras(4)
    rgba, fr = ro.render_views(v, f, R, T, 64, texture=tex, verts_uvs=uv, faces_uvs=fuv, return_fragments=True, nthreads=8)
    R2, T2 = ro.random_cameras(2, generator=torch.Generator().manual_seed(0))
    rgba2, fr2 = ro.render_views(v, f, R2, T2, 96, texture=tex, verts_uvs=uv, faces_uvs=fuv, return_fragments=True, nthreads=8)
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"),
                        R=R.numpy(), T=T.numpy(), p2f=fr["pix_to_face"].numpy().astype(np.int32),
                        rgba=rgba.numpy().astype(np.float16),
                        R2=R2.numpy(), T2=T2.numpy(), p2f2=fr2["pix_to_face"].numpy().astype(np.int32),
                        rgba2=rgba2.numpy().astype(np.float16))
    print("render golden coverage:", float((fr["pix_to_face"] >= 0).float().mean()), float((fr2["pix_to_face"] >= 0).float().mean()))


if __name__ == "__main__":
    make_loss_golden()
    make_cow()
    make_bob_teapot_styles()
    make_render_golden()

"""Drop-in for the reference's utils.py (same names, same arguments).  `render_meshes` renders ALL the
cameras it is given in one fused libst3d launch sequence instead of one renderer call per view
(utils.py:65-77); everything else is setup / I/O glue around the hot path."""
import os
import random

import torch
from PIL import Image
from pytorch3d.renderer import FoVPerspectiveCameras, TexturesUV
from pytorch3d.renderer.cameras import look_at_view_transform
from pytorch3d.structures import Meshes
from pytorch3d.transforms import RotateAxisAngle
from torchvision import models, transforms

from style_transfer import *  # noqa: F401,F403

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def apply_background(tensors, masks, background_type="noise", background=None):
    """utils.py:19-30: composite the render over fresh uniform noise / the style image / nothing."""
    if background_type == "white":
        return tensors                      # the shader's background is already white
    if background_type == "noise":
        fill = torch.rand(tensors.shape, device=tensors.device)     # torch's generator, as the reference draws it
    elif background_type == "style":
        fill = background
    else:
        return None
    if tensors.is_cuda and tensors.dtype == torch.float32 and tensors.dim() == 4 and masks.shape[1] == 1:
        from st3d import functional as _fn
        return _fn.composite_background(tensors, masks, fill)       # one libst3d kernel (+ one in the backward)
    return tensors * masks + fill * (1 - masks)


_decoded = {}       # (path, size, mtime, bytes) -> decoded tensor on `device`


def load_as_tensor(image_path, size=512):
    """utils.py:34-44: RGB image squashed to size x size, (3,size,size) in [0,1].
    second_approach.py:157 calls this for the style image in EVERY batch of every epoch (a JPEG decode + resize of
    10-20 ms on the host): the decoded tensor is kept per (file, size, mtime) and every call
    returns a fresh copy of it, so callers may still write into what they get."""
    st = os.stat(image_path)
    key = (os.path.abspath(image_path), int(size) if isinstance(size, int) else tuple(size), st.st_mtime_ns, st.st_size)
    cached = _decoded.get(key)
    if cached is None:
        with Image.open(image_path) as im:
            im = im.convert("RGB")
            tensor = transforms.ToTensor()(transforms.Resize((size, size))(im))
        if len(_decoded) >= 16:
            _decoded.pop(next(iter(_decoded)))
        cached = _decoded[key] = tensor[:3].to(device)
    return cached.clone()


def get_vgg():
    """utils.py:48-52: frozen VGG-19 `.features` with ImageNet weights.  On CUDA the module is returned with
    cuDNN's fused conv+bias+ReLU calls and channels_last weights (same module names, same tapped values)."""
    if os.environ.get("ST3D_VGG_RANDOM_INIT") == "1":
        # offline boxes (no ImageNet checkpoint): the same architecture with seeded random weights; the global RNG
        # stream is left as it was, so the cameras drawn after this call do not depend on the switch
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(int(os.environ.get("ST3D_VGG_SEED", "0")))
            vgg = models.vgg19(weights=None).features.eval().to(device)
    else:
        vgg = models.vgg19(weights=models.VGG19_Weights.IMAGENET1K_V1).features.to(device)
    for p in vgg.parameters():
        p.requires_grad_(False)
    if device.type == "cuda":
        from st3d.vgg import fuse_vgg_features
        vgg = fuse_vgg_features(vgg, channels_last=True)
    return vgg


# ---- image dumps (second_approach.py:183-185 writes every view of every batch as a PNG) ----------------
# A PNG of a 512^2 view costs ~20 ms of host time to encode -- 8 views per step against a 10 ms GPU step.  For CUDA tensors
# tensor_to_image therefore returns a stand-in for the PIL image: the 8-bit conversion runs on the GPU, the copy to pinned host
# memory is asynchronous, and `.save(path)` hands the encoding to a worker thread (the same path always to the same worker, so a
# later dump of a view replaces an earlier one in order; PNGs at zlib level 1).  Everything is written before the interpreter exits, or when
# flush_image_writes() is called; any other use of the object builds the real PIL image on the spot.
# ST3D_SYNC_IMAGE_WRITES=1 restores the synchronous behaviour.
_WRITERS, _PENDING, _WRITER_OF = [], [], {}


def _writer_for(path):
    from concurrent.futures import ThreadPoolExecutor
    if not _WRITERS:
        import atexit
        n = max(2, min(16, (os.cpu_count() or 4) - 1))
        _WRITERS.extend(ThreadPoolExecutor(max_workers=1, thread_name_prefix="st3d-png") for _ in range(n))
        atexit.register(flush_image_writes)
    key = os.path.abspath(path)
    if key not in _WRITER_OF:                   # paths are spread evenly; one path stays with one (FIFO) worker
        _WRITER_OF[key] = len(_WRITER_OF) % len(_WRITERS)
    return _WRITERS[_WRITER_OF[key]]


def flush_image_writes():
    """Waits for every image handed to `.save()` so far; re-raises the first error a writer met."""
    pending, _PENDING[:] = list(_PENDING), []
    errors = [f.exception() for f in pending]
    for e in errors:
        if e is not None:
            raise e


class _PendingImage:
    def __init__(self, tensor):
        t = tensor.detach().squeeze(0).clamp(0, 1)
        u8 = t.mul(255).to(torch.uint8)                                # ToPILImage: pic.mul(255).byte()
        u8 = (u8.permute(1, 2, 0) if u8.dim() == 3 else u8).contiguous()
        self._host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
        self._host.copy_(u8, non_blocking=True)
        self._ready = torch.cuda.Event()
        self._ready.record()
        self._pil = None

    def _image(self):
        if self._pil is None:
            self._ready.synchronize()
            a = self._host.numpy()
            self._pil = Image.fromarray(a[..., 0] if a.ndim == 3 and a.shape[-1] == 1 else a)
        return self._pil

    def save(self, path, *args, **kwargs):
        if isinstance(path, (str, os.PathLike)):
            if str(path).lower().endswith(".png"):
                kwargs.setdefault("compress_level", 1)                  # same pixels, a third of the encoding time
            _PENDING.append(_writer_for(path).submit(lambda: self._image().save(path, *args, **kwargs)))
            if len(_PENDING) > 256:                                     # bound the backlog (and surface errors early)
                flush_image_writes()
        else:
            self._image().save(path, *args, **kwargs)                   # a file object: the caller wants it written now

    def __getattr__(self, name):
        return getattr(self._image(), name)


def tensor_to_image(tensor):
    """utils.py:56-61.  CUDA tensors: see the note on image dumps above."""
    if tensor.is_cuda and os.environ.get("ST3D_SYNC_IMAGE_WRITES") != "1":
        return _PendingImage(tensor)
    return transforms.ToPILImage()(tensor.detach().clone().squeeze(0).clamp(0, 1).cpu())


def render_meshes(renderer, meshes, cameras):
    """-> ((B,3,H,W) images, (B,1,H,W) masks).  `cameras` is a list of camera objects or a camera batch."""
    if not isinstance(cameras, FoVPerspectiveCameras):
        cameras = FoVPerspectiveCameras.join(list(cameras))
    return renderer.render_planar(meshes, cameras=cameras)


def save_render(renderer, meshes, cameras, path):
    os.makedirs(path, exist_ok=True)
    tensors, _ = render_meshes(renderer, meshes, cameras)
    for i, t in enumerate(tensors):
        tensor_to_image(t).save(f"{path}/view_{i}.png")


def finalize_tensor(tensor):
    return tensor.clamp(0.0, 1.0).detach()


def finalize_mesh(mesh):
    """utils.py:94-113: same geometry, texture clamped to [0,1]."""
    tex = mesh.textures
    clamped = TexturesUV(verts_uvs=tex.verts_uvs_padded(), faces_uvs=tex.faces_uvs_padded(),
                         maps=finalize_tensor(tex.maps_padded()))
    return Meshes(verts=mesh.verts_padded(), faces=mesh.faces_padded(), textures=clamped)


def build_fixed_cameras(n_views, dist=3.0, shuffle=True):
    """utils.py:121-151: views on two great circles (about X, then about Y) at distance `dist`."""
    n_x = n_views // 2
    views = [(a.item(), "X") for a in torch.linspace(0, 315, n_x)]
    views += [(a.item(), "Y") for a in torch.linspace(45, 315, n_views - n_x)]
    if shuffle:
        random.shuffle(views)
    R = torch.stack([RotateAxisAngle(a, axis=ax, device=device).get_matrix()[0, :3, :3] for a, ax in views], dim=0)
    T = torch.tensor([0.0, 0.0, dist], device=device).repeat(len(views), 1)
    return FoVPerspectiveCameras(R=R, T=T, device=device)


def build_random_cameras(n_views, dist=2.10):
    """utils.py:154-170: uniform directions on the sphere (cos-elevation and azimuth uniform)."""
    elev = torch.acos(torch.rand(n_views) * 2 - 1) * 180 / torch.pi - 90
    azim = torch.rand(n_views) * 360 - 180
    R, T = look_at_view_transform(dist=dist, elev=elev, azim=azim, at=((0, 0.10, 0.25),))
    return FoVPerspectiveCameras(R=R, T=T, device=device)


def setup_optimizations(optimization_target, mesh, lr):
    """utils.py:173-204: clone the mesh, mark the chosen leaves, build Adam."""
    work = mesh.clone()
    leaves = {"texture_map": work.textures.maps_padded(), "verts": work.verts_packed()}
    chosen = {"texture": ["texture_map"], "mesh": ["verts"], "both": ["verts", "texture_map"]}[optimization_target]
    for name in chosen:
        leaves[name].requires_grad_(True)
    optimizer = torch.optim.Adam([leaves[name] for name in chosen], lr=lr)
    return {"optimizable_mesh": work, "optimizer": optimizer, "texture_map": leaves["texture_map"],
            "verts": leaves["verts"], "faces": work.faces_packed(),
            "verts_uvs": work.textures.verts_uvs_padded(), "faces_uvs": work.textures.faces_uvs_padded()}


def build_mesh(verts_uvs, faces_uvs, texture_map, verts, faces):
    """utils.py:207-210."""
    return Meshes(verts=[verts], faces=[faces],
                  textures=TexturesUV(verts_uvs=verts_uvs, faces_uvs=faces_uvs, maps=texture_map))

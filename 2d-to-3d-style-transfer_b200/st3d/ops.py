"""Thin tensor-level wrappers over the libst3d C ABI (include/st3d.h).

Every function takes CUDA fp32 tensors, passes raw device pointers plus the current torch stream
and raises on a non-zero return code.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes
import functools
from dataclasses import dataclass, field
from typing import Optional

import torch

from ._lib import RenderArgs, St3dError, check, lib

TEX_UV, TEX_VERTEX = 0, 1
LIGHT_AMBIENT, LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1, 2
LAYOUT_NHWC_RGBA, LAYOUT_PLANAR, LAYOUT_NHWC_RGB = 0, 1, 2
MAX_FACES_PER_PIXEL = 8
WS_HEADER_INTS = 16


# ------------------------------------------------------------------------------------------------
# optional per-op CUDA-event timing (bench.py): events go on the launching stream, nothing synchronises
# ------------------------------------------------------------------------------------------------
_profile: Optional[list] = None


def start_profile() -> None:
    global _profile
    _profile = []


def stop_profile():
    """Returns {(op, key): [ms, ...]} for every op call since start_profile(); synchronises once."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    out: dict = {}
    for name, key, e0, e1 in rec:
        out.setdefault((name, key), []).append(e0.elapsed_time(e1))
    return out


class _timed:
    def __init__(self, name, key=None):
        self.name, self.key = name, key

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profile.append((self.name, self.key, self.e0, e1))
        return False


def launch_count() -> int:
    """Kernels launched by libst3d in this process so far."""
    return int(lib().st3d_launch_count())


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    """The current torch stream of the CURRENT device; every op runs under `_on_device_of`, which makes the device
    of its tensors current for the duration of the call (the library launches on the current device)."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a
        if isinstance(a, RenderState):
            return a.keep[0]
    return None


def _on_device_of(fn):
    """Runs `fn` with the device of its first CUDA tensor argument current (what a CUDAGuard does in upstream's
    `_C` ops): kernels, workspace allocations and the stream all belong to that device even when the caller's current
    device is another one.  `_cuda_f32` / `_cuda_int` / `_feat3` then reject tensors that live elsewhere."""
    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        t = _first_tensor(args, kwargs)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return guarded


def _same_device(name: str, t: torch.Tensor) -> None:
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"{name}: tensor is on {t.device} but the call runs on cuda:{torch.cuda.current_device()} "
                         "(all tensor arguments of one call must share a device)")


def _cuda_f32(name: str, t: torch.Tensor, *shape_tail):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise St3dError(f"{name}: expected a CUDA tensor (st3d has no CPU path; the CPU oracle is test-only)")
    _same_device(name, t)
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {t.dtype}")
    if shape_tail and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError(f"{name}: expected trailing shape {shape_tail}, got {tuple(t.shape)}")
    return t.contiguous()


def _cuda_int(name: str, t: torch.Tensor, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise St3dError(f"{name}: expected a CUDA tensor (st3d has no CPU path)")
    _same_device(name, t)
    return t.to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------
# deferred bin-overflow checks (the library never synchronises; see st3d.h workspace header)
# ------------------------------------------------------------------------------------------------
_pending: list = []
_capacity_hint: dict = {}


_free_hosts: list = []      # recycled pinned header buffers (cudaHostAlloc costs milliseconds; never allocate per call)


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def read_header(ws: torch.Tensor):
    """Synchronous read of a raster workspace header (for callers that replay captured graphs and therefore own
    the workspace): (entries needed, overflow flag, 1 if the call clipped faces at the near plane)."""
    h = ws[: WS_HEADER_INTS * 4].view(torch.int32).cpu()
    return int(h[0]) + int(h[6]), int(h[1]), int(h[5])


_captured_headers: Optional[list] = None     # set by `collect_captured_headers` while a graph is being recorded


class collect_captured_headers:
    """Context manager for the owner of a CUDA graph: every raster call recorded inside also records a copy of its
    workspace header to a pinned host buffer; `.headers` lists them as (host int32 tensor, capacity key).  Each replay
    refreshes the buffers, `check_captured_headers` reads them (after the replay has finished)."""

    def __enter__(self):
        global _captured_headers
        self.headers = _captured_headers = []
        return self

    def __exit__(self, *exc):
        global _captured_headers
        _captured_headers = None
        return False


def check_captured_headers(headers) -> None:
    """Raise if a raster call of the LAST finished replay overflowed its work lists or met a case it does not handle
    (the caller has waited for that replay)."""
    for host, key in headers:
        needed, overflow = int(host[0]) + int(host[6]), int(host[1])
        _capacity_hint[key] = max(_capacity_hint.get(key, 0), needed)
        if overflow:
            raise St3dError(f"a replayed render needed {needed} work-list entries, more than the workspace recorded in "
                            "the graph holds (the mesh moved too far from the one that was captured); the results of that "
                            "replay are invalid. Capture again: the capacity hint has been raised.")


def _watch_header(ws: torch.Tensor, key) -> None:
    if _capturing():        # a captured call is replayed without Python: its owner checks the header
        if _captured_headers is not None:
            host = torch.empty(WS_HEADER_INTS, dtype=torch.int32, pin_memory=True)
            host.copy_(ws[: WS_HEADER_INTS * 4].view(torch.int32), non_blocking=True)
            _captured_headers.append((host, key))
        return
    host = _free_hosts.pop() if _free_hosts else torch.empty(WS_HEADER_INTS, dtype=torch.int32, pin_memory=True)
    host.copy_(ws[: WS_HEADER_INTS * 4].view(torch.int32), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _pending.append((ev, host, key))
    if len(_pending) > 64:      # bound the backlog when the caller never polls
        poll_overflow()


def poll_overflow(block: bool = False) -> None:
    """Raise if an earlier raster call overflowed its tile-bin pair buffer."""
    if _capturing():
        return
    keep = []
    for ev, host, key in _pending:
        if block:
            ev.synchronize()
        if ev.query():
            needed, overflow = int(host[0]) + int(host[6]), int(host[1])   # [6]: back-facing units
            _free_hosts.append(host)
            _capacity_hint[key] = max(_capacity_hint.get(key, 0), needed)
            if overflow:
                _pending.clear()
                raise St3dError(f"tile bins overflowed: {needed} (face,tile) pairs needed; results of that call are "
                                "invalid. Re-run: the capacity hint has been raised.")
        else:
            keep.append((ev, host, key))
    _pending[:] = keep


def _capacity(key, total_faces: int) -> int:
    hint = _capacity_hint.get(key)
    if hint is None:
        return 8 * total_faces + 4096
    return max(2 * hint, total_faces) + 4096


# ------------------------------------------------------------------------------------------------
# operator boundary (mirrors pytorch3d._C)
# ------------------------------------------------------------------------------------------------
@_on_device_of
def transform_verts(verts, R, T, k00: float, k11: float):
    verts = _cuda_f32("verts", verts, 3)
    R = _cuda_f32("R", R, 3, 3).reshape(-1, 3, 3)
    T = _cuda_f32("T", T, 3).reshape(-1, 3)
    N, V = R.shape[0], verts.shape[0]
    out = torch.empty((N, V, 3), device=verts.device, dtype=torch.float32)
    check(lib().st3d_transform_verts_forward(_p(verts), _p(R), _p(T), k00, k11, N, V, _p(out), _stream()),
          "st3d_transform_verts_forward")
    return out


@_on_device_of
def transform_verts_backward(verts, R, T, k00, k11, grad_ndc):
    verts = _cuda_f32("verts", verts, 3)
    R = _cuda_f32("R", R, 3, 3).reshape(-1, 3, 3)
    T = _cuda_f32("T", T, 3).reshape(-1, 3)
    grad_ndc = _cuda_f32("grad_ndc", grad_ndc, 3)
    g = torch.zeros_like(verts)
    check(lib().st3d_transform_verts_backward(_p(verts), _p(R), _p(T), k00, k11, R.shape[0], verts.shape[0],
                                              _p(grad_ndc), _p(g), _stream()), "st3d_transform_verts_backward")
    return g


@_on_device_of
def rasterize_meshes(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, image_size, blur_radius=0.0,
                     faces_per_pixel=1, bin_size=0, max_faces_per_bin=0, perspective_correct=False,
                     clip_barycentric_coords=False, cull_backfaces=False, clipped_faces_neighbor_idx=None):
    """Signature of pytorch3d._C.rasterize_meshes.  Returns (pix_to_face i64, zbuf, bary, dists)."""
    poll_overflow()
    face_verts = _cuda_f32("face_verts", face_verts, 3, 3)
    first = _cuda_int("mesh_to_face_first_idx", mesh_to_face_first_idx, torch.int64)
    num = _cuda_int("num_faces_per_mesh", num_faces_per_mesh, torch.int64)
    H, W = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
    N, Ft, K = first.numel(), face_verts.shape[0], int(faces_per_pixel)
    if num.numel() != N:
        raise ValueError("mesh_to_face_first_idx and num_faces_per_mesh differ in length")
    dev = face_verts.device
    p2f = torch.empty((N, H, W, K), device=dev, dtype=torch.int64)
    zbuf = torch.empty((N, H, W, K), device=dev, dtype=torch.float32)
    bary = torch.empty((N, H, W, K, 3), device=dev, dtype=torch.float32)
    dists = torch.empty((N, H, W, K), device=dev, dtype=torch.float32)
    if N == 0:
        return p2f, zbuf, bary, dists
    max_f = int(num.max().item()) if N > 1 else Ft
    key = ("raster", N, Ft, H, W)
    nbytes = lib().st3d_raster_workspace_size(N, Ft, H, W, _capacity(key, Ft))
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    nbi = None
    if clipped_faces_neighbor_idx is not None:
        nbi = _cuda_int("clipped_faces_neighbor_idx", clipped_faces_neighbor_idx, torch.int64)
        if nbi.numel() != Ft:
            raise ValueError("clipped_faces_neighbor_idx must hold one entry per face")
    check(lib().st3d_rasterize_meshes_forward(_p(face_verts), _p(first), _p(num), _p(nbi), N, Ft, max_f, H, W, float(blur_radius),
                                              K, int(bin_size or 0), int(max_faces_per_bin or 0), int(perspective_correct),
                                              int(clip_barycentric_coords), int(cull_backfaces), _p(ws), nbytes,
                                              _p(p2f), _p(zbuf), _p(bary), _p(dists), _stream()),
          "st3d_rasterize_meshes_forward")
    _watch_header(ws, key)
    return p2f, zbuf, bary, dists


@_on_device_of
def rasterize_meshes_backward(face_verts, pix_to_face, grad_zbuf, grad_bary, grad_dists, perspective_correct,
                              clip_barycentric_coords):
    face_verts = _cuda_f32("face_verts", face_verts, 3, 3)
    p2f = _cuda_int("pix_to_face", pix_to_face, torch.int64)
    N, H, W, K = p2f.shape
    gz = _cuda_f32("grad_zbuf", grad_zbuf)
    gb = _cuda_f32("grad_bary", grad_bary, 3)
    gd = _cuda_f32("grad_dists", grad_dists)
    out = torch.zeros_like(face_verts)
    check(lib().st3d_rasterize_meshes_backward(_p(face_verts), _p(p2f), _p(gz), _p(gb), _p(gd), N, H, W, K,
                                               face_verts.shape[0], int(perspective_correct),
                                               int(clip_barycentric_coords), _p(out), _stream()),
          "st3d_rasterize_meshes_backward")
    return out


@_on_device_of
def interp_face_attrs_forward(pix_to_face, bary, face_attrs):
    p2f = _cuda_int("pix_to_face", pix_to_face, torch.int64).reshape(-1)
    bary = _cuda_f32("bary", bary, 3).reshape(-1, 3)
    fa = _cuda_f32("face_attrs", face_attrs)
    if fa.dim() != 3 or fa.shape[1] != 3:
        raise ValueError("face_attrs must be (F,3,D)")
    P, F, D = p2f.numel(), fa.shape[0], fa.shape[2]
    out = torch.empty((P, D), device=fa.device, dtype=torch.float32)
    check(lib().st3d_interp_face_attrs_forward(_p(p2f), _p(bary), _p(fa), P, F, D, _p(out), _stream()),
          "st3d_interp_face_attrs_forward")
    return out


@_on_device_of
def interp_face_attrs_backward(pix_to_face, bary, face_attrs, grad_out):
    p2f = _cuda_int("pix_to_face", pix_to_face, torch.int64).reshape(-1)
    bary = _cuda_f32("bary", bary, 3).reshape(-1, 3)
    fa = _cuda_f32("face_attrs", face_attrs)
    go = _cuda_f32("grad_out", grad_out).reshape(p2f.numel(), -1)
    P, F, D = p2f.numel(), fa.shape[0], fa.shape[2]
    g_bary = torch.empty((P, 3), device=fa.device, dtype=torch.float32)
    g_attrs = torch.zeros_like(fa)
    check(lib().st3d_interp_face_attrs_backward(_p(p2f), _p(bary), _p(fa), _p(go), P, F, D, _p(g_bary), _p(g_attrs),
                                                _stream()), "st3d_interp_face_attrs_backward")
    return g_bary, g_attrs


# ------------------------------------------------------------------------------------------------
# fused multi-view renderer
# ------------------------------------------------------------------------------------------------
@dataclass
class RenderSpec:
    """Non-tensor settings of one fused render call (cameras' fov folded into k00/k11)."""
    image_size: tuple
    k00: float
    k11: float
    znear: float = 1.0
    zfar: float = 100.0
    blur_radius: float = 0.0
    cull_backfaces: bool = False
    cull_to_frustum: bool = False
    ambient: tuple = (1.0, 1.0, 1.0)
    background: tuple = (1.0, 1.0, 1.0)
    sigma: float = 1e-4
    gamma: float = 1e-4
    layout: int = LAYOUT_NHWC_RGBA
    z_clip: Optional[float] = None      # None -> znear / 2 (PyTorch3D's default for perspective cameras)
    # Point / Directional lights (phong_shading) evaluated in the same epilogue; `ambient` above is
    # light.ambient x material.ambient, light_diffuse / light_specular are light colour x material colour
    light_kind: int = LIGHT_AMBIENT
    light_vec: tuple = (0.0, 1.0, 0.0)
    light_diffuse: tuple = (0.0, 0.0, 0.0)
    light_specular: tuple = (0.0, 0.0, 0.0)
    shininess: float = 64.0


@dataclass
class RenderState:
    """Everything backward needs; tensors are kept alive here."""
    args: RenderArgs
    keep: list = field(default_factory=list)
    workspace: Optional[torch.Tensor] = None


@_on_device_of
def render_forward(spec: RenderSpec, verts, faces, R, T, *, face_uvs=None, texture=None, verts_rgb=None,
                   background_image=None):
    """One launch sequence for N views of one mesh.  Returns (image, mask|None, pix_to_face i32, state).
    background_image (1|N,3,H,W): pixels no face covers take their colour from it (apply_background of utils.py:19-30
    fused into the epilogue) instead of spec.background."""
    poll_overflow()
    verts = _cuda_f32("verts", verts, 3)
    faces = _cuda_int("faces", faces, torch.int32)
    R = _cuda_f32("R", R, 3, 3).reshape(-1, 3, 3)
    T = _cuda_f32("T", T, 3).reshape(-1, 3)
    if R.shape[0] != T.shape[0]:
        raise ValueError("R and T describe different numbers of cameras")
    if faces.dim() != 2 or faces.shape[1] != 3:
        raise ValueError("faces must be (F,3)")
    N, V, F = R.shape[0], verts.shape[0], faces.shape[0]
    H, W = spec.image_size
    dev = verts.device
    a = RenderArgs()
    a.verts, a.faces, a.V, a.F = _p(verts), _p(faces), V, F
    a.R, a.T, a.N = _p(R), _p(T), N
    a.k00, a.k11, a.znear, a.zfar = spec.k00, spec.k11, spec.znear, spec.zfar
    a.H, a.W, a.blur_radius, a.cull_backfaces = H, W, spec.blur_radius, int(spec.cull_backfaces)
    a.cull_to_frustum = int(spec.cull_to_frustum)
    keep = [verts, faces, R, T]
    if texture is not None:
        texture = _cuda_f32("texture", texture, 3)
        texture = texture.reshape(texture.shape[-3], texture.shape[-2], 3)
        face_uvs = _cuda_f32("face_uvs", face_uvs, 3, 2)
        if face_uvs.shape[0] != F:
            raise ValueError("face_uvs must be (F,3,2)")
        a.tex_mode, a.face_uvs, a.texture = TEX_UV, _p(face_uvs), _p(texture)
        a.Ht, a.Wt = texture.shape[0], texture.shape[1]
        keep += [texture, face_uvs]
    elif verts_rgb is not None:
        verts_rgb = _cuda_f32("verts_rgb", verts_rgb, 3)
        if verts_rgb.shape[0] != V:
            raise ValueError("verts_rgb must be (V,3)")
        a.tex_mode, a.verts_rgb = TEX_VERTEX, _p(verts_rgb)
        keep += [verts_rgb]
    else:
        raise ValueError("render_forward needs either texture+face_uvs or verts_rgb")
    a.ambient = (ctypes.c_float * 3)(*spec.ambient)
    a.background = (ctypes.c_float * 3)(*spec.background)
    a.sigma, a.gamma, a.out_layout = spec.sigma, spec.gamma, spec.layout
    a.light_kind = int(spec.light_kind)
    a.light_vec = (ctypes.c_float * 3)(*spec.light_vec)
    a.light_diffuse = (ctypes.c_float * 3)(*spec.light_diffuse)
    a.light_specular = (ctypes.c_float * 3)(*spec.light_specular)
    a.shininess = float(spec.shininess)
    if background_image is not None:
        background_image = _cuda_f32("background_image", background_image, 3, H, W).reshape(-1, 3, H, W)
        if background_image.shape[0] not in (1, N):
            raise ValueError(f"background_image batch {background_image.shape[0]} is neither 1 nor {N}")
        a.background_image, a.background_batch = _p(background_image), background_image.shape[0]
        keep.append(background_image)
    if spec.layout == LAYOUT_NHWC_RGBA:
        image = torch.empty((N, H, W, 4), device=dev, dtype=torch.float32)
        mask = None
    elif spec.layout == LAYOUT_NHWC_RGB:    # an (N,3,H,W) tensor in channels_last storage
        image = torch.empty((N, H, W, 3), device=dev, dtype=torch.float32).permute(0, 3, 1, 2)
        mask = torch.empty((N, 1, H, W), device=dev, dtype=torch.float32)
    else:
        image = torch.empty((N, 3, H, W), device=dev, dtype=torch.float32)
        mask = torch.empty((N, 1, H, W), device=dev, dtype=torch.float32)
    p2f = torch.empty((N, H, W), device=dev, dtype=torch.int32)
    a.out_image, a.out_mask, a.pix_to_face = _p(image), _p(mask), _p(p2f)
    key = ("render", N, F, H, W)
    cap = _capacity(key, N * F)
    nbytes = lib().st3d_render_workspace_size(N, V, F, H, W, cap)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    a.workspace, a.workspace_bytes, a.list_capacity = _p(ws), nbytes, cap
    a.z_clip = spec.znear / 2.0 if spec.z_clip is None else float(spec.z_clip)
    with _timed("render_forward", (N, H, W, F)):
        check(lib().st3d_render_forward(ctypes.byref(a), _stream()), "st3d_render_forward")
    if N > 0:
        _watch_header(ws, key)
    return image, mask, p2f, RenderState(args=a, keep=keep + [image, mask, p2f], workspace=ws)


@_on_device_of
def render_backward(state: RenderState, grad_image, need_texture=True, need_verts=False, need_verts_rgb=False):
    """Returns (grad_texture|None, grad_verts|None, grad_verts_rgb|None)."""
    a = state.args
    if a.out_layout == LAYOUT_NHWC_RGB:     # the gradient in the storage order of the image: (N,H,W,3)
        if not (isinstance(grad_image, torch.Tensor) and grad_image.is_cuda and grad_image.dtype == torch.float32):
            raise St3dError("grad_image: expected a CUDA float32 tensor (st3d has no CPU path)")
        _same_device("grad_image", grad_image)
        if tuple(grad_image.shape) != (a.N, 3, a.H, a.W):
            raise ValueError(f"grad_image: expected {(a.N, 3, a.H, a.W)}, got {tuple(grad_image.shape)}")
        grad_image = grad_image.permute(0, 2, 3, 1).contiguous()        # no copy when it already is channels_last
    else:
        grad_image = _cuda_f32("grad_image", grad_image)
    dev = grad_image.device
    g_tex = torch.zeros((a.Ht, a.Wt, 3), device=dev, dtype=torch.float32) if (need_texture and a.tex_mode == TEX_UV) else None
    if need_verts and a.light_kind != LIGHT_AMBIENT:
        raise NotImplementedError("vertex gradients under Point / Directional lights: use MeshRasterizer + SoftPhongShader "
                                  "(the operator-boundary path); the fused renderer differentiates lit renders w.r.t. "
                                  "the texture / vertex colours only")
    g_verts = torch.zeros((a.V, 3), device=dev, dtype=torch.float32) if need_verts else None
    g_rgb = torch.zeros((a.V, 3), device=dev, dtype=torch.float32) if (need_verts_rgb and a.tex_mode == TEX_VERTEX) else None
    # texel-padded zero scratch: the kernel scatters with 16-byte vector reductions, a second kernel folds it into g_tex
    scratch = torch.zeros((a.Ht, a.Wt, 4), device=dev, dtype=torch.float32) if g_tex is not None else None
    a.grad_texture_scratch = _p(scratch)
    with _timed("render_backward", (a.N, a.H, a.W, a.F)):
        check(lib().st3d_render_backward(ctypes.byref(a), _p(grad_image), _p(g_tex), _p(g_verts), _p(g_rgb), _stream()),
              "st3d_render_backward")
    a.grad_texture_scratch = None
    return g_tex, g_verts, g_rgb


# ------------------------------------------------------------------------------------------------
# Gram / MSE losses (style_transfer.py:31-35, losses.py:31-39, losses.py:71-75)
# ------------------------------------------------------------------------------------------------
GRAM_TF32, GRAM_FP32 = 0, 1
_PRECISIONS = {"tf32": GRAM_TF32, "fp32": GRAM_FP32, GRAM_TF32: GRAM_TF32, GRAM_FP32: GRAM_FP32}


def gram_tc_supported(C: int, HW: int) -> bool:
    """Shapes the tcgen05 (kind::tf32) Gram kernels accept; mirrors gram_tc_supported() in csrc."""
    return C in (64, 128, 256, 512) and HW % 4 == 0 and HW >= 32


def _precision(precision, C: int, HW: int) -> int:
    if precision is None:
        return GRAM_TF32 if gram_tc_supported(C, HW) else GRAM_FP32
    try:
        return _PRECISIONS[precision]
    except KeyError:
        raise ValueError(f"precision must be 'tf32', 'fp32' or None, got {precision!r}") from None


FEAT_NCHW, FEAT_NHWC = 0, 1


def _feat3(name, feat):
    """-> (tensor whose storage the kernel reads, layout, (B, C, HW)).  A channels_last (B,C,H,W) tensor is
    passed as is (its storage is (B, HW, C)); anything else is made (B, C, HW)-contiguous."""
    if not isinstance(feat, torch.Tensor) or not feat.is_cuda:
        raise St3dError(f"{name}: expected a CUDA tensor (st3d has no CPU path; the CPU oracle is test-only)")
    if feat.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {feat.dtype}")
    _same_device(name, feat)
    if feat.dim() == 4:
        B, C, H, W = feat.shape
        if C > 1 and H * W > 1 and feat.is_contiguous(memory_format=torch.channels_last) and not feat.is_contiguous():
            return feat, FEAT_NHWC, (B, C, H * W)
        return feat.contiguous().reshape(B, C, H * W), FEAT_NCHW, (B, C, H * W)
    if feat.dim() != 3:
        raise ValueError(f"{name}: expected (B,C,H,W) or (B,C,HW)")
    return feat.contiguous(), FEAT_NCHW, tuple(feat.shape)


def _gram_ws(B, C, HW, device):
    nbytes = lib().st3d_gram_workspace_size(B, C, HW)
    return torch.empty(nbytes, device=device, dtype=torch.uint8), nbytes


@_on_device_of
def gram_forward(feat, precision=None):
    """(B,C,H,W) -> (B,C,C) = F F^T.  NCHW-contiguous and channels_last inputs are both read in place."""
    f, layout, (B, C, HW) = _feat3("feat", feat)
    out = torch.empty((B, C, C), device=f.device, dtype=torch.float32)
    if B == 0:
        return out
    ws, nbytes = _gram_ws(B, C, HW, f.device)
    with _timed("gram_forward", (B, C, HW)):
        check(lib().st3d_gram_forward(_p(f), B, C, HW, _p(out), _p(ws), nbytes, _precision(precision, C, HW), layout,
                                      _stream()), "st3d_gram_forward")
    return out


@_on_device_of
def gram_mse_forward(feat, target, scale: float, loss_out, want_gram=False, precision=None):
    """loss_out[0] += scale * sum((F F^T - target)^2); returns (dgram, gram|None)."""
    f, layout, (B, C, HW) = _feat3("feat", feat)
    target = _cuda_f32("target", target, C, C).reshape(-1, C, C)
    if target.shape[0] not in (1, B):
        raise ValueError(f"target batch {target.shape[0]} is neither 1 nor {B}")
    if loss_out.dtype != torch.float32 or not loss_out.is_cuda:
        raise ValueError("loss_out must be a CUDA float32 tensor")
    dgram = torch.empty((B, C, C), device=f.device, dtype=torch.float32)
    gram = torch.empty((B, C, C), device=f.device, dtype=torch.float32) if want_gram else None
    ws, nbytes = _gram_ws(B, C, HW, f.device)
    with _timed("gram_mse_forward", (B, C, HW)):
        check(lib().st3d_gram_mse_forward(_p(f), _p(target), B, target.shape[0], C, HW, float(scale), _p(gram), _p(dgram),
                                          _p(loss_out), _p(ws), nbytes, _precision(precision, C, HW), layout, _stream()),
              "st3d_gram_mse_forward")
    return dgram, gram


@_on_device_of
def gram_backward(feat, dgram, grad_scale: float = 1.0, out=None, accumulate=False, precision=None, scale_tensor=None,
                  relu_mask=False, symmetric_dgram=False):
    """grad_feat = grad_scale * [scale_tensor] * (dG + dG^T) F, with the shape AND memory layout of feat
    (scale_tensor: 1-element CUDA tensor).  `out`, when given, must have feat's layout; with `accumulate` the
    result is added to it.  `relu_mask` then zeroes the elements whose feat value is <= 0 (the backward of the
    ReLU that produced feat, applied to the sum) -- both fused into the kernel's epilogue.  `symmetric_dgram`: dgram is
    symmetric (what gram_mse_forward returns), so the dG + dG^T pass is skipped."""
    f, layout, (B, C, HW) = _feat3("feat", feat)
    dgram = _cuda_f32("dgram", dgram, C, C).reshape(B, C, C)
    if out is None:
        out = torch.empty_like(f)           # preserve_format: channels_last stays channels_last
        accumulate = False
    else:
        if layout == FEAT_NCHW and out.dim() == 4 and out.is_contiguous() and out.shape == feat.shape:
            out = out.view(B, C, HW)        # same storage: the kernel reads and writes (B, C, HW)
        if out.shape != f.shape or out.stride() != f.stride() or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError("gram_backward: `out` must match the (normalised) feature tensor in shape and strides")
    ws, nbytes = _gram_ws(B, C, HW, f.device)
    if scale_tensor is not None:
        scale_tensor = _cuda_f32("scale_tensor", scale_tensor).reshape(1)
    flags = (1 if accumulate else 0) | (2 if relu_mask else 0)
    with _timed("gram_backward" + ("", "_acc", "_relu", "_acc_relu")[flags], (B, C, HW)):
        flags |= 4 if symmetric_dgram else 0
        check(lib().st3d_gram_backward(_p(f), _p(dgram), B, C, HW, float(grad_scale), _p(scale_tensor), flags,
                                       _p(out), _p(ws), nbytes, _precision(precision, C, HW), layout, _stream()),
              "st3d_gram_backward")
    return out if layout == FEAT_NHWC else out.reshape(feat.shape)


@_on_device_of
def composite_forward(image, mask, fill):
    """apply_background (utils.py:19-30) in one pass: image * mask + fill * (1 - mask); image (B,C,H,W), mask (B,1,H,W),
    fill (1|B,C,H,W)."""
    image = _cuda_f32("image", image)
    if image.dim() != 4:
        raise ValueError("composite_forward: image must be (B,C,H,W)")
    B, C, H, W = image.shape
    mask = _cuda_f32("mask", mask, 1, H, W)
    fill = _cuda_f32("fill", fill, C, H, W).reshape(-1, C, H, W)
    if mask.shape[0] != B or fill.shape[0] not in (1, B):
        raise ValueError("composite_forward: mask / fill batch does not match the image")
    out = torch.empty_like(image)
    with _timed("composite_forward", (image.numel(),)):
        check(lib().st3d_composite_forward(_p(image), _p(mask), _p(fill), image.numel(), H * W, C, fill.shape[0], _p(out),
                                           _stream()), "st3d_composite_forward")
    return out


@_on_device_of
def composite_backward(grad_out, mask):
    grad_out = _cuda_f32("grad_out", grad_out)
    B, C, H, W = grad_out.shape
    mask = _cuda_f32("mask", mask, 1, H, W)
    g = torch.empty_like(grad_out)
    check(lib().st3d_composite_backward(_p(grad_out), _p(mask), grad_out.numel(), H * W, C, _p(g), _stream()),
          "st3d_composite_backward")
    return g


@_on_device_of
def mse_forward(a, b, scale: float, loss_out, mask=None, want_grad=True):
    """loss_out[0] += scale * sum(m (a-b)^2); returns grad wrt a (or None).  mask: (B,1,H,W) for a (B,Cm,H,W)."""
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
    same_dense_layout = (mask is None and isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.dim() == 4
                         and a.is_cuda and b.is_cuda and a.dtype == b.dtype == torch.float32 and a.stride() == b.stride()
                         and a.is_contiguous(memory_format=torch.channels_last))
    if not same_dense_layout:       # an elementwise reduction does not care about the order of a dense layout
        a = _cuda_f32("a", a)
        b = _cuda_f32("b", b)
    inner, mask_ch = 1, 1
    if mask is not None:
        mask = _cuda_f32("mask", mask)
        if a.dim() != 4 or mask.shape != (a.shape[0], 1, a.shape[2], a.shape[3]):
            raise ValueError("mask must be (B,1,H,W) for a (B,C,H,W) input")
        inner, mask_ch = a.shape[2] * a.shape[3], a.shape[1]
    grad = torch.empty_like(a) if want_grad else None
    with _timed("mse_forward" if want_grad else "mse_value_forward", (a.numel(),)):
        check(lib().st3d_mse_forward(_p(a), _p(b), _p(mask), a.numel(), inner, mask_ch, float(scale), _p(loss_out),
                                     _p(grad), _stream()), "st3d_mse_forward")
    return grad


@_on_device_of
def mse_tap_backward(y, c, grad_in, scale: float, scale_tensor=None):
    """(y > 0) ? grad_in + 2 scale [* scale_tensor] (y - c) : 0 in one pass (st3d_mse_tap_backward).  y, c and grad_in
    must share one dense layout; the result has it too.  grad_in may be None."""
    if not (isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == torch.float32):
        raise St3dError("mse_tap_backward: expected CUDA float32 tensors (st3d has no CPU path)")
    _same_device("y", y)
    dense = y.is_contiguous() or (y.dim() == 4 and y.is_contiguous(memory_format=torch.channels_last))
    if not dense or c.shape != y.shape or c.stride() != y.stride() or c.dtype != torch.float32 or not c.is_cuda:
        raise ValueError("mse_tap_backward: y and c must be dense tensors of one shape and layout")
    if grad_in is not None and (grad_in.shape != y.shape or grad_in.stride() != y.stride() or grad_in.dtype != torch.float32):
        raise ValueError("mse_tap_backward: grad_in must match y in shape and layout")
    if y.numel() % 4:
        raise ValueError("mse_tap_backward: the number of elements must be a multiple of 4")
    out = torch.empty_like(y)
    if scale_tensor is not None:
        scale_tensor = _cuda_f32("scale_tensor", scale_tensor).reshape(1)
    with _timed("mse_tap_backward", (y.numel(),)):
        check(lib().st3d_mse_tap_backward(_p(y), _p(c), _p(grad_in), y.numel(), float(scale), _p(scale_tensor), _p(out),
                                          _stream()), "st3d_mse_tap_backward")
    return out


# ------------------------------------------------------------------------------------------------
# mesh regularisers (losses.py:85-87): one forward and one backward launch over cached topology tables
# ------------------------------------------------------------------------------------------------
MESH_EDGE, MESH_LAPLACIAN, MESH_NORMAL = 1, 2, 4
_mesh_reg_ws: dict = {}


def _mesh_reg_args(verts, topo, target_length, which, lap_dir):
    from ._lib import MeshRegArgs
    a = MeshRegArgs()
    a.verts, a.V = verts.data_ptr(), verts.shape[0]
    a.edges, a.E = topo.edges.data_ptr(), topo.edges.shape[0]
    a.adj_ptr, a.adj_idx = topo.adj_ptr.data_ptr(), topo.adj_idx.data_ptr()
    a.pairs, a.P = topo.pairs.data_ptr(), topo.pairs.shape[0]
    a.target_length, a.which = float(target_length), int(which)
    a.lap_dir = lap_dir.data_ptr()
    return a


def _check_topology(verts, topo):
    V, dev = verts.shape[0], verts.device
    for name in ("edges", "adj_ptr", "adj_idx", "pairs"):
        t = getattr(topo, name)
        if not (t.is_cuda and t.device == dev and t.dtype == torch.int32 and t.is_contiguous()):
            raise ValueError(f"mesh topology: {name} must be a contiguous int32 tensor on {dev}")
    if topo.num_verts != V or topo.adj_ptr.shape[0] != V + 1 or topo.adj_idx.shape[0] != 2 * topo.edges.shape[0]:
        raise ValueError("mesh topology: tables were built for another vertex count")


@_on_device_of
def mesh_regularizers_forward(verts, topo, target_length: float = 0.0, which: int = 7):
    """-> (losses (3,) [edge, laplacian, normal consistency], state for the backward).  `topo`: st3d.mesh_losses.topology."""
    verts = _cuda_f32("verts", verts, 3)
    _check_topology(verts, topo)
    dev = verts.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)     # launches on one stream are ordered; two streams must not share
    ws = _mesh_reg_ws.get(key)
    if ws is None:      # 32 bytes, zero once: every launch leaves it zero
        ws = _mesh_reg_ws[key] = torch.zeros(int(lib().st3d_mesh_regularizers_workspace_size()) // 8, device=dev,
                                             dtype=torch.float64)
    losses = torch.empty(3, device=dev, dtype=torch.float32)
    lap_dir = torch.empty_like(verts)
    a = _mesh_reg_args(verts, topo, target_length, which, lap_dir)
    with _timed("mesh_regularizers_forward", (verts.shape[0], topo.edges.shape[0], topo.pairs.shape[0])):
        check(lib().st3d_mesh_regularizers_forward(ctypes.byref(a), _p(ws), _p(losses), _stream()),
              "st3d_mesh_regularizers_forward")
    return losses, (verts, topo, float(target_length), int(which), lap_dir)


@_on_device_of
def mesh_regularizers_backward(state, grad_losses):
    verts, topo, target_length, which, lap_dir = state
    grad_losses = _cuda_f32("grad_losses", grad_losses, 3)
    grad = torch.empty_like(verts)
    a = _mesh_reg_args(verts, topo, target_length, which, lap_dir)
    with _timed("mesh_regularizers_backward", (verts.shape[0], topo.edges.shape[0], topo.pairs.shape[0])):
        check(lib().st3d_mesh_regularizers_backward(ctypes.byref(a), _p(grad_losses), _p(grad), _stream()),
              "st3d_mesh_regularizers_backward")
    return grad


# ------------------------------------------------------------------------------------------------
# 2x2 max pooling of channels_last feature maps (VGG-19 pools; the convolutions stay on cuDNN)
# ------------------------------------------------------------------------------------------------
def maxpool_supported(x: torch.Tensor) -> bool:
    """Shapes / layouts the libst3d pooling kernels accept: CUDA fp32 (B,C,H,W) in channels_last storage,
    even H and W, C a multiple of 4."""
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and x.shape[1] % 4 == 0 and x.shape[1] > 1
            and x.is_contiguous(memory_format=torch.channels_last))


@_on_device_of
def maxpool2x2_forward(x: torch.Tensor) -> torch.Tensor:
    if not maxpool_supported(x):
        raise ValueError("maxpool2x2_forward: expected a CUDA float32 channels_last (B,C,H,W) tensor with even H, W "
                         "and C % 4 == 0")
    B, C, H, W = x.shape
    y = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
    with _timed("maxpool_forward", (B, C, H, W)):
        check(lib().st3d_maxpool2x2_forward(_p(x), B, H, W, C, _p(y), _stream()), "st3d_maxpool2x2_forward")
    return y


@_on_device_of
def maxpool2x2_backward(x: torch.Tensor, grad_y: torch.Tensor, relu_mask: bool = False) -> torch.Tensor:
    """Gradient w.r.t. x; with relu_mask also through the ReLU that produced x (zero where x <= 0)."""
    B, C, H, W = x.shape
    if not maxpool_supported(x) or grad_y.shape != (B, C, H // 2, W // 2):
        raise ValueError("maxpool2x2_backward: x / grad_y do not match the forward call")
    grad_y = grad_y.contiguous(memory_format=torch.channels_last)
    gx = torch.empty_like(x)
    with _timed("maxpool_backward", (B, C, H, W)):
        check(lib().st3d_maxpool2x2_backward(_p(x), _p(grad_y), B, H, W, C, int(relu_mask), _p(gx), _stream()),
              "st3d_maxpool2x2_backward")
    return gx

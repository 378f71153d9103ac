"""ctypes loader for libst3d.so (C ABI declared in include/st3d.h)."""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_PKG, "lib", "libst3d.so")
_LIB = None

c_f = ctypes.c_float
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_p = ctypes.c_void_p


class St3dError(RuntimeError):
    pass


class RenderArgs(ctypes.Structure):
    """Mirror of `struct st3d_render_args` (include/st3d.h)."""
    _fields_ = [
        ("verts", c_p), ("faces", c_p), ("V", c_i64), ("F", c_i64),
        ("R", c_p), ("T", c_p), ("N", c_i),
        ("k00", c_f), ("k11", c_f), ("znear", c_f), ("zfar", c_f),
        ("H", c_i), ("W", c_i), ("blur_radius", c_f), ("cull_backfaces", c_i),
        ("tex_mode", c_i), ("face_uvs", c_p), ("texture", c_p), ("Ht", c_i), ("Wt", c_i), ("verts_rgb", c_p),
        ("ambient", c_f * 3), ("background", c_f * 3), ("sigma", c_f), ("gamma", c_f),
        ("out_layout", c_i), ("out_image", c_p), ("out_mask", c_p), ("pix_to_face", c_p),
        ("workspace", c_p), ("workspace_bytes", c_sz), ("list_capacity", c_i64), ("z_clip", c_f),
        ("light_kind", c_i), ("light_vec", c_f * 3), ("light_diffuse", c_f * 3), ("light_specular", c_f * 3),
        ("shininess", c_f), ("background_image", c_p), ("background_batch", c_i), ("cull_to_frustum", c_i),
        ("grad_texture_scratch", c_p),
    ]


class MeshRegArgs(ctypes.Structure):
    """Mirror of `struct st3d_mesh_reg_args` (include/st3d.h)."""
    _fields_ = [
        ("verts", c_p), ("V", c_i64), ("edges", c_p), ("E", c_i64), ("adj_ptr", c_p), ("adj_idx", c_p),
        ("pairs", c_p), ("P", c_i64), ("target_length", c_f), ("which", c_i), ("lap_dir", c_p),
    ]


# name -> (restype, argtypes): every symbol include/st3d.h declares
SIGNATURES = {
    "st3d_last_error": (ctypes.c_char_p, []),
    "st3d_version": (c_i, []),
    "st3d_launch_count": (ctypes.c_ulonglong, []),
    "st3d_transform_verts_forward": (c_i, [c_p, c_p, c_p, c_f, c_f, c_i, c_i64, c_p, c_p]),
    "st3d_transform_verts_backward": (c_i, [c_p, c_p, c_p, c_f, c_f, c_i, c_i64, c_p, c_p, c_p]),
    "st3d_raster_workspace_size": (c_sz, [c_i, c_i64, c_i, c_i, c_i64]),
    "st3d_rasterize_meshes_forward": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i64, c_i64, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_i,
                                            c_i, c_p, c_sz, c_p, c_p, c_p, c_p, c_p]),
    "st3d_rasterize_meshes_backward": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i64, c_i, c_i, c_p, c_p]),
    "st3d_interp_face_attrs_forward": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_i, c_p, c_p]),
    "st3d_interp_face_attrs_backward": (c_i, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i, c_p, c_p, c_p]),
    "st3d_render_workspace_size": (c_sz, [c_i, c_i64, c_i64, c_i, c_i, c_i64]),
    "st3d_render_forward": (c_i, [ctypes.POINTER(RenderArgs), c_p]),
    "st3d_render_backward": (c_i, [ctypes.POINTER(RenderArgs), c_p, c_p, c_p, c_p, c_p]),
    "st3d_gram_workspace_size": (c_sz, [c_i, c_i, c_i64]),
    "st3d_gram_forward": (c_i, [c_p, c_i, c_i, c_i64, c_p, c_p, c_sz, c_i, c_i, c_p]),
    "st3d_gram_mse_forward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i64, c_f, c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_p]),
    "st3d_gram_backward": (c_i, [c_p, c_p, c_i, c_i, c_i64, c_f, c_p, c_i, c_p, c_p, c_sz, c_i, c_i, c_p]),
    "st3d_composite_forward": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_i, c_i, c_p, c_p]),
    "st3d_composite_backward": (c_i, [c_p, c_p, c_i64, c_i64, c_i, c_p, c_p]),
    "st3d_mse_forward": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_i, c_f, c_p, c_p, c_p]),
    "st3d_mse_tap_backward": (c_i, [c_p, c_p, c_p, c_i64, c_f, c_p, c_p, c_p]),
    "st3d_mesh_regularizers_workspace_size": (c_i64, []),
    "st3d_mesh_regularizers_forward": (c_i, [c_p, c_p, c_p, c_p]),
    "st3d_mesh_regularizers_backward": (c_i, [c_p, c_p, c_p, c_p]),
    "st3d_maxpool2x2_forward": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "st3d_maxpool2x2_backward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
}


def library_path() -> str:
    return _SO


def lib():
    """Load libst3d.so; fail loudly (no fallback) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise St3dError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(_SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().st3d_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise St3dError(f"{what} failed (code {rc}): {msg}")

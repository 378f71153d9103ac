"""The C-ABI boundary: libst3d.so loads on a CPU-only box, exports every symbol include/st3d.h declares,
and the ctypes signatures mirror the header (argument counts).  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "st3d.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(st3d_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[name] = n
    return decls


def test_header_declares_the_expected_entry_points():
    d = _declared()
    for name in ("st3d_rasterize_meshes_forward", "st3d_rasterize_meshes_backward", "st3d_interp_face_attrs_forward",
                 "st3d_interp_face_attrs_backward", "st3d_render_forward", "st3d_render_backward",
                 "st3d_gram_forward", "st3d_gram_mse_forward", "st3d_gram_backward", "st3d_mse_forward",
                 "st3d_transform_verts_forward", "st3d_transform_verts_backward", "st3d_last_error"):
        assert name in d, name


def test_library_exports_every_declared_symbol():
    import st3d
    lib = st3d.lib()                       # raises loudly if the .so is missing
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in st3d.h but not exported by libst3d.so"


def test_ctypes_signatures_match_header():
    from st3d import _lib
    d = _declared()
    assert set(_lib.SIGNATURES) == set(d), set(_lib.SIGNATURES) ^ set(d)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == d[name], f"{name}: ctypes has {len(argtypes)} args, header {d[name]}"


def test_render_args_struct_matches_header():
    from st3d import _lib
    src = open(HEADER).read()
    body = re.search(r"typedef struct st3d_render_args \{(.*?)\} st3d_render_args;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")))
    assert names == [f[0] for f in _lib.RenderArgs._fields_]
    assert ctypes.sizeof(_lib.RenderArgs) % 8 == 0


def test_pure_queries_work_without_a_gpu():
    import st3d
    lib = st3d.lib()
    assert lib.st3d_version() >= 100
    assert lib.st3d_raster_workspace_size(2, 100, 64, 64, 0) > 0
    assert lib.st3d_render_workspace_size(8, 2930, 5856, 512, 512, 0) > lib.st3d_raster_workspace_size(8, 8 * 5856, 512, 512, 0)
    assert lib.st3d_gram_workspace_size(8, 64, 512 * 512) >= 8 * 64 * 64 * 4
    assert lib.st3d_launch_count() == 0


def test_ops_fail_loudly_without_cuda():
    import torch
    import st3d
    from st3d import ops
    with pytest.raises(st3d.St3dError):
        ops.gram_forward(torch.rand(1, 64, 8, 8))
    with pytest.raises(st3d.St3dError):
        ops.transform_verts(torch.rand(4, 3), torch.eye(3)[None], torch.zeros(1, 3), 1.0, 1.0)
    with pytest.raises(ValueError):
        ops._precision("bf16", 64, 1024)
    assert ops._precision(None, 64, 1024) == ops.GRAM_TF32 and ops._precision(None, 8, 35) == ops.GRAM_FP32

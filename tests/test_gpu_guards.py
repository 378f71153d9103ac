"""Out-of-bounds WRITE check of every libst3d entry point (compute-sanitizer is closed on the GPU pool).

Every buffer `st3d.ops` allocates for a kernel -- outputs, gradients, workspaces -- is carved out of a larger allocation
with a 4 KiB band of 0xA5 bytes on either side (the allocation calls of the `torch` name inside st3d.ops are replaced for
the duration of a test).  After the op has run the bands must be intact: a store one element, one row or one tile past
the end of a buffer, or before its start, lands in them.  Shapes are chosen odd / non-multiples of the tile sizes, where
such overruns live.  (Reads past a buffer are not caught by this.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096
FILL = 0xA5


class _GuardedTorch:
    """Stands in for the `torch` module inside st3d.ops: empty / zeros / *_like hand out guarded CUDA buffers."""

    def __init__(self):
        self.records = []
        self.checked = 0

    def __getattr__(self, name):
        return getattr(torch, name)

    def _carve(self, shape, dtype, device, strides=None, zero=False):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        padded = (nbytes + 255) // 256 * 256
        buf = torch.full((GUARD + padded + GUARD,), FILL, dtype=torch.uint8, device=device)
        payload = buf[GUARD:GUARD + nbytes]
        if zero:
            payload.zero_()
        t = payload.view(dtype)
        t = t.as_strided(shape, strides) if strides is not None else t.view(shape)
        self.records.append((buf, nbytes))
        return t

    def empty(self, *size, dtype=torch.float32, device=None, pin_memory=False, memory_format=None, **kw):
        if pin_memory or device is None or torch.device(device).type != "cuda":
            return torch.empty(*size, dtype=dtype, device=device, pin_memory=pin_memory, **kw)
        shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
        strides = None
        if memory_format == torch.channels_last:
            B, C, H, W = shape
            strides = (H * W * C, 1, W * C, C)
        return self._carve(shape, dtype, device, strides)

    def zeros(self, *size, dtype=torch.float32, device=None, **kw):
        if device is None or torch.device(device).type != "cuda":
            return torch.zeros(*size, dtype=dtype, device=device, **kw)
        shape = size[0] if len(size) == 1 and not isinstance(size[0], int) else size
        return self._carve(shape, dtype, device, zero=True)

    def _like(self, x, zero):
        if not x.is_cuda:
            return torch.zeros_like(x) if zero else torch.empty_like(x)
        dense = x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))
        return self._carve(x.shape, x.dtype, x.device, None if x.is_contiguous() or not dense else x.stride(), zero)

    def empty_like(self, x, **kw):
        return self._like(x, False)

    def zeros_like(self, x, **kw):
        return self._like(x, True)

    def verify(self, what):
        torch.cuda.synchronize()
        for buf, nbytes in self.records:
            head, tail = buf[:GUARD], buf[GUARD + nbytes:]
            assert bool((head == FILL).all()), f"{what}: write BEFORE a {nbytes}-byte buffer"
            assert bool((tail == FILL).all()), f"{what}: write PAST the end of a {nbytes}-byte buffer"
        self.checked += len(self.records)
        self.records.clear()


@pytest.fixture()
def guarded(monkeypatch):
    import st3d
    from st3d import ops
    st3d.lib()                  # raises when libst3d.so is not built
    g = _GuardedTorch()
    monkeypatch.setattr(ops, "torch", g)
    monkeypatch.setattr(ops, "_mesh_reg_ws", {})
    yield ops, g
    assert g.checked > 0, "no guarded allocation was checked"


def _cams(n, seed):
    from oracle import render_oracle as ro
    return ro.random_cameras(n, generator=torch.Generator().manual_seed(seed))


def test_render_ops_stay_inside_their_buffers(guarded, cow):
    from oracle import render_oracle as ro
    ops, g = guarded
    dev = "cuda"
    k00, k11 = ro.fov_scales(60.0)
    R, T = _cams(3, 5)
    R, T = R.to(dev), T.to(dev)
    verts, faces = cow["verts"].to(dev), cow["faces"].to(dev)
    fuv = cow["verts_uvs"][cow["faces_uvs"]].to(dev)
    gen = torch.Generator().manual_seed(0)
    tex = torch.rand(37, 53, 3, generator=gen).to(dev)
    vrgb = torch.rand(verts.shape[0], 3, generator=gen).to(dev)
    ndc = ops.transform_verts(verts, R, T, k00, k11)
    ops.transform_verts_backward(verts, R, T, k00, k11, torch.randn_like(ndc))
    g.verify("transform_verts")
    for size in ((45, 71), (129, 65), (16, 16)):
        H, W = size
        bg = torch.rand(1, 3, H, W, generator=gen).to(dev)
        for layout in (ops.LAYOUT_NHWC_RGBA, ops.LAYOUT_PLANAR, ops.LAYOUT_NHWC_RGB):
            for kw in (dict(face_uvs=fuv, texture=tex), dict(verts_rgb=vrgb)):
                spec = ops.RenderSpec(image_size=size, k00=k00, k11=k11, layout=layout)
                img, mask, p2f, state = ops.render_forward(spec, verts, faces.int(), R, T, background_image=bg, **kw)
                g.verify(f"render_forward {size} layout {layout}")
                assert int((p2f >= 0).sum()) > 0
                go = torch.randn(img.shape, generator=gen).to(dev)
                if layout == ops.LAYOUT_NHWC_RGB:
                    go = go.contiguous(memory_format=torch.channels_last)
                ops.render_backward(state, go, need_texture=True, need_verts=True, need_verts_rgb=True)
                g.verify(f"render_backward {size} layout {layout}")
    # a camera inside the mesh: faces cut by the near plane go through the clipped-unit code
    Rc, Tc = R.clone(), T.clone()
    Tc[:, 2] = 0.6
    spec = ops.RenderSpec(image_size=(67, 41), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
    img, mask, p2f, state = ops.render_forward(spec, verts, faces.int(), Rc, Tc, face_uvs=fuv, texture=tex)
    ops.render_backward(state, torch.randn_like(img), need_texture=True, need_verts=True)
    g.verify("render of a near-plane-clipped scene")
    spec = ops.RenderSpec(image_size=(67, 41), k00=k00, k11=k11, layout=ops.LAYOUT_NHWC_RGBA, blur_radius=6e-4)
    img, mask, p2f, state = ops.render_forward(spec, verts, faces.int(), Rc, Tc, face_uvs=fuv, texture=tex)
    ops.render_backward(state, torch.randn_like(img), need_texture=True, need_verts=True)
    g.verify("soft render (tile bins) of a near-plane-clipped scene")
    ops.poll_overflow(block=True)


@pytest.mark.parametrize("K,blur", [(1, 0.0), (3, 0.0), (4, 3e-4)])
def test_operator_boundary_ops_stay_inside_their_buffers(guarded, cow, K, blur):
    from oracle import render_oracle as ro
    ops, g = guarded
    dev = "cuda"
    k00, k11 = ro.fov_scales(60.0)
    R, T = _cams(2, 9)
    verts, faces = cow["verts"].to(dev), cow["faces"].to(dev)
    ndc = ops.transform_verts(verts, R.to(dev), T.to(dev), k00, k11)
    fv = torch.cat([ndc[i][faces] for i in range(2)], dim=0)
    F = faces.shape[0]
    first = torch.tensor([0, F], device=dev)
    num = torch.tensor([F, F], device=dev)
    g.records.clear()
    for size in ((39, 57), (64, 64)):
        p2f, zbuf, bary, dists = ops.rasterize_meshes(fv, first, num, size, blur_radius=blur, faces_per_pixel=K,
                                                      perspective_correct=True)
        g.verify(f"rasterize_meshes {size} K={K}")
        ops.rasterize_meshes_backward(fv, p2f, torch.randn_like(zbuf), torch.randn_like(bary), torch.randn_like(dists), True,
                                      False)
        g.verify("rasterize_meshes_backward")
        attrs = torch.rand(fv.shape[0], 3, 5, device=dev)
        out = ops.interp_face_attrs_forward(p2f, bary, attrs)
        ops.interp_face_attrs_backward(p2f, bary, attrs, torch.randn_like(out))
        g.verify("interp_face_attrs")
    ops.poll_overflow(block=True)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("B,C,H,W", [(3, 64, 20, 12), (1, 64, 6, 6), (3, 128, 24, 20), (2, 256, 36, 28), (1, 256, 16, 8),
                                       (3, 512, 8, 12), (1, 512, 32, 32)])
def test_gram_ops_stay_inside_their_buffers(guarded, precision, layout, B, C, H, W):
    ops, g = guarded
    gen = torch.Generator().manual_seed(C + H)
    f = torch.randn(B, C, H, W, generator=gen).cuda()
    if layout == "nhwc":
        f = f.contiguous(memory_format=torch.channels_last)
    gram = ops.gram_forward(f, precision=precision)
    g.verify("gram_forward")
    for target in (gram[:1].clone(), gram.clone()):
        loss = torch.zeros(1, device="cuda")
        dgram, _ = ops.gram_mse_forward(f, target * 0.9, 1e-3, loss, want_gram=True, precision=precision)
        g.verify("gram_mse_forward")
    for flags in (dict(), dict(symmetric_dgram=True)):
        grad = ops.gram_backward(f, dgram, 1.0, precision=precision, **flags)
        g.verify(f"gram_backward {flags}")
    out = torch.zeros_like(f)           # (a plain torch tensor: the caller's buffer; the split-K workspace is guarded)
    ops.gram_backward(f, dgram, 1.0, out=out, accumulate=True, relu_mask=True, precision=precision, symmetric_dgram=True)
    g.verify("gram_backward accumulate + relu mask")
    assert torch.isfinite(grad).all() and torch.isfinite(out).all()


def test_loss_and_pool_ops_stay_inside_their_buffers(guarded, golden_dir):
    from st3d import mesh_losses as ml
    ops, g = guarded
    dev = "cuda"
    gen = torch.Generator().manual_seed(3)
    for shape in ((2, 3, 37, 53), (1, 512, 6, 10), (3, 4, 2, 2)):
        a, b = torch.randn(shape, generator=gen).to(dev), torch.randn(shape, generator=gen).to(dev)
        mask = (torch.rand(shape[0], 1, shape[2], shape[3], generator=gen) > 0.5).float().to(dev)
        for kw in (dict(), dict(mask=mask)):
            loss = torch.zeros(1, device=dev)
            ops.mse_forward(a, b, 0.1, loss, want_grad=True, **kw)
            g.verify(f"mse_forward {shape} {list(kw)}")
        a_cl, b_cl = (t.contiguous(memory_format=torch.channels_last) for t in (a, b))
        if a.numel() % 4 == 0:
            ops.mse_tap_backward(a_cl, b_cl, torch.randn_like(a_cl), 0.3)
            g.verify("mse_tap_backward")
        fill = torch.rand(1, *shape[1:], generator=gen).to(dev)
        out = ops.composite_forward(a, mask, fill)
        ops.composite_backward(torch.randn_like(out), mask)
        g.verify("composite")
    for shape in ((3, 64, 10, 6), (1, 4, 2, 2), (2, 128, 34, 18)):
        x = torch.randn(shape, generator=gen).to(dev).contiguous(memory_format=torch.channels_last)
        y = ops.maxpool2x2_forward(x)
        ops.maxpool2x2_backward(x, torch.randn_like(y), relu_mask=True)
        g.verify(f"maxpool {shape}")
    d = np.load(os.path.join(golden_dir, "teapot_mesh.npz"))
    verts, faces = torch.from_numpy(d["verts"]).float().to(dev), torch.from_numpy(d["faces"]).long().to(dev)
    topo = ml.topology(faces, verts.shape[0])
    for which in (1, 2, 4, 7):
        losses, state = ops.mesh_regularizers_forward(verts, topo, 0.0, which)
        ops.mesh_regularizers_backward(state, torch.ones(3, device=dev))
        g.verify(f"mesh regularisers which={which}")


def test_the_guard_bands_do_see_an_overrun(guarded):
    """The check itself: a kernel told to write 128 floats into a 64-float guarded buffer is caught."""
    import ctypes
    from st3d import lib
    ops, g = guarded
    out = g.empty(64, dtype=torch.float32, device="cuda")
    y, c = torch.rand(128, device="cuda"), torch.rand(128, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
    rc = lib().st3d_mse_tap_backward(p(y), p(c), None, 128, ctypes.c_float(1.0), None, p(out),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    with pytest.raises(AssertionError, match="PAST the end"):
        g.verify("deliberate overrun")
    g.records.clear()
    g.checked += 1

"""Render backward (texture only) at C2 and 1024^2, CUDA events, GPU-bound loop (GPU box only)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np, torch
from st3d import ops, functional as Fn, cameras as cm
d = np.load(os.path.join(ROOT, "tests/golden/cow_mesh.npz"))
verts = torch.from_numpy(d["verts"]).cuda(); faces = torch.from_numpy(d["faces"]).int().cuda()
fuv = torch.from_numpy(d["verts_uvs"])[torch.from_numpy(d["faces_uvs"]).long()].cuda()
res = {}
for S in (512, 1024):
    N = 8
    tex = torch.rand(S, S, 3, device="cuda")
    R, T = cm.random_view_cameras(N, generator=torch.Generator().manual_seed(0)); R, T = R.cuda(), T.cuda()
    k00, k11 = Fn.fov_scales(60.0)
    spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
    out = ops.render_forward(spec, verts, faces, R, T, face_uvs=fuv, texture=tex)
    g = torch.randn(N, 3, S, S, device="cuda")
    a = out[3].args
    gt = torch.zeros(S, S, 3, device="cuda")
    import ctypes
    from st3d._lib import lib
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    scratch = torch.zeros(S, S, 4, device="cuda")
    for label, ptr in (("scalar_reds", None), ("vector_reds", ctypes.c_void_p(scratch.data_ptr()))):
        a.grad_texture_scratch = ptr
        def call():
            if ptr is not None:
                scratch.zero_()         # the caller's zero fill is part of the cost
            lib().st3d_render_backward(ctypes.byref(a), ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(gt.data_ptr()), None, None, st)
        for _ in range(5): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): call()
        e1.record(); torch.cuda.synchronize()
        res[f"bwd_tex_{S}_{label}_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 2)
print(json.dumps(res))

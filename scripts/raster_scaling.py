"""Render forward/backward time vs mesh density and resolution (GPU box only): cow subdivided 0..4 times
(5.9 k .. 1.5 M faces, BASELINE configs[4] scale), 8 views."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np, torch
from st3d import ops, functional as Fn, cameras as cm

from st3d.meshgen import subdivide

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps * 1e3

d = np.load(os.path.join(ROOT, "tests/golden/cow_mesh.npz"))
verts, faces = torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long()
uvs, fuvs = torch.from_numpy(d["verts_uvs"]), torch.from_numpy(d["faces_uvs"]).long()
N = 8
R, T = cm.random_view_cameras(N, generator=torch.Generator().manual_seed(0)); R, T = R.cuda(), T.cuda()
k00, k11 = Fn.fov_scales(60.0)
rows = []
levels = [int(t) for t in os.environ.get("LEVELS", "0,1,2,3,4").split(",")]
for level in range(5):
    if level:
        verts, faces = subdivide(verts, faces); uvs, fuvs = subdivide(uvs, fuvs)
    if level not in levels:
        continue
    v, f, fuv = verts.cuda(), faces.int().cuda(), uvs[fuvs].cuda()
    for S in (512, 1024):
        tex = torch.rand(S, S, 3, device="cuda")
        spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
        st = {}
        def fwd(): st["o"] = ops.render_forward(spec, v, f, R, T, face_uvs=fuv, texture=tex)
        fwd(); torch.cuda.synchronize()
        try: ops.poll_overflow(block=True)
        except Exception: fwd(); torch.cuda.synchronize(); ops.poll_overflow(block=True)
        t_f = timeit(fwd)
        g = torch.randn(N, 3, S, S, device="cuda")
        t_b = timeit(lambda: ops.render_backward(st["o"][3], g))
        t_bv = timeit(lambda: ops.render_backward(st["o"][3], g, need_verts=True))
        Fn_ = f.shape[0]
        fb = N * Fn_ * 36 + N * S * S * 20 + S * S * 12
        bb = N * S * S * 16 + N * Fn_ * 48 + S * S * 12
        rows.append(dict(faces=Fn_, size=S, fwd_us=round(t_f, 1), bwd_tex_us=round(t_b, 1), bwd_tex_verts_us=round(t_bv, 1),
                         fwd_GBs=round(fb / t_f / 1e3, 1), bwd_GBs=round(bb / t_b / 1e3, 1),
                         coverage=round(float((st["o"][2] >= 0).float().mean()), 3)))
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "raster_scaling.json")), "w"), indent=1)

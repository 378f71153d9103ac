"""Mesh regularisers forward + backward (losses.py:85-87), one launch each (csrc/mesh_reg.cu) against the torch
formulation, on the three meshes of the configs and the 1.5 M-face cow: wall time per call (what an eager step pays:
launches + the host read of the torch formulation) and CUDA-event time.  GPU box only."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np
import torch
from st3d import mesh_losses as ml, meshgen

res = {}
w = torch.tensor([1.0, 1.0, 1.0], device="cuda")


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / iters * 1e3, 4), round(e0.elapsed_time(e1) / iters, 4)


meshes = {}
for name in ("cow", "bob", "teapot"):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_mesh.npz"))
    meshes[name] = (torch.from_numpy(d["verts"]).float(), torch.from_numpy(d["faces"]).long())
v, f = meshes["cow"]
for _ in range(4):
    v, f = meshgen.subdivide(v, f)
meshes["cow_x256"] = (v, f)

only = [m for m in os.environ.get("MESHES", "").split(",") if m]
for name, (verts, faces) in meshes.items():
    if only and name not in only:
        continue
    verts, faces = verts.cuda().requires_grad_(True), faces.cuda()
    t0 = time.perf_counter()
    topo = ml.topology(faces, verts.shape[0])
    torch.cuda.synchronize()
    build_ms = round((time.perf_counter() - t0) * 1e3, 2)

    def fused():
        verts.grad = None
        torch.dot(ml.regularizers(verts, faces, topo=topo), w).backward()

    def torch_ops():
        verts.grad = None
        (ml.edge_loss_torch(verts, faces) + ml.laplacian_smoothing_torch(verts, faces)
         + ml.normal_consistency_torch(verts, faces)).backward()

    iters = 20 if faces.shape[0] > 100000 else 100
    fw, fe = timed(fused, iters)
    g_fused = verts.grad.clone()
    tw, te = timed(torch_ops, iters)
    rel = ((g_fused - verts.grad).abs().max() / verts.grad.abs().max()).item()
    res[name] = dict(faces=faces.shape[0], edges=topo.edges.shape[0], pairs=topo.pairs.shape[0], tables_build_ms=build_ms,
                     fused_wall_ms=fw, fused_gpu_ms=fe, torch_wall_ms=tw, torch_gpu_ms=te, grad_rel_diff=rel)
print(json.dumps(res))

"""Runs the fused C2 render forward/backward a few times (GPU box only; used under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np, torch
from st3d import ops, functional as Fn, cameras as cm
d = np.load(os.path.join(ROOT, "tests/golden/cow_mesh.npz"))
dev = "cuda"
verts, faces = torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long()
uvs, fuvs = torch.from_numpy(d["verts_uvs"]), torch.from_numpy(d["faces_uvs"]).long()
from st3d.meshgen import subdivide
for _ in range(int(os.environ.get("SUBDIV", 0))):      # SUBDIV=4: 1.5 M faces (BASELINE configs[4] scale)
    verts, faces = subdivide(verts, faces); uvs, fuvs = subdivide(uvs, fuvs)
fuv = uvs[fuvs].to(dev); verts = verts.to(dev); faces = faces.int().to(dev)
S = int(os.environ.get("SIZE", 512)); N = int(os.environ.get("VIEWS", 8))
tex = torch.rand(S, S, 3, device=dev)
R, T = cm.random_view_cameras(N, generator=torch.Generator().manual_seed(0)); R, T = R.to(dev), T.to(dev)
k00, k11 = Fn.fov_scales(60.0)
spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
g = torch.randn(N, 3, S, S, device=dev)
for i in range(int(os.environ.get("REPS", 3))):
    out = ops.render_forward(spec, verts, faces, R, T, face_uvs=fuv, texture=tex)
    ops.render_backward(out[3], g, need_verts=os.environ.get("NEED_VERTS", "1") == "1")
torch.cuda.synchronize(); ops.poll_overflow(block=True)
print("ok", float((out[2] >= 0).float().mean()))

"""Where the host time of a tiny (1 view x 256^2) iteration goes (GPU box only)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import torch, bench
from st3d import functional as Fn
dev = torch.device("cuda:0"); size = 256
w = bench.load_workload(size); R, T = bench.cameras(1)
verts, faces = w["verts"].to(dev), w["faces"].int().to(dev)
fuv = w["verts_uvs"][w["faces_uvs"]].to(dev)
tex = w["texture"].to(dev).clone().requires_grad_(True)
opt = torch.optim.Adam([tex], lr=0.01)
Rd, Td = R.to(dev), T.to(dev); tgt = torch.rand(1, 3, size, size, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    img, mask, _ = Fn.render_views(verts, faces, Rd, Td, size, texture=tex, face_uvs=fuv)
    loss = Fn.masked_mse_loss(img, tgt, mask); loss.backward(); opt.step()
for _ in range(10): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t0) / 200 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

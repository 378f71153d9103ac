"""Times the libst3d ops of the C2 workload in isolation with CUDA events (GPU box only).
Each op is run `reps` times back to back; inputs exceed L2 where the real ones do."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np, torch
from st3d import ops, functional as Fn, cameras as cm

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us

d = np.load(os.path.join(ROOT, "tests/golden/cow_mesh.npz"))
dev = "cuda"
verts = torch.from_numpy(d["verts"]).to(dev); faces = torch.from_numpy(d["faces"]).int().to(dev)
fuv = torch.from_numpy(d["verts_uvs"])[torch.from_numpy(d["faces_uvs"]).long()].to(dev)
S = int(os.environ.get("SIZE", 512)); N = int(os.environ.get("VIEWS", 8))
tex = torch.rand(S, S, 3, device=dev)
R, T = cm.random_view_cameras(N, generator=torch.Generator().manual_seed(0)); R, T = R.to(dev), T.to(dev)
k00, k11 = Fn.fov_scales(60.0)
spec = ops.RenderSpec(image_size=(S, S), k00=k00, k11=k11, layout=ops.LAYOUT_PLANAR)
res = {}
state = {}
def rf():
    state["o"] = ops.render_forward(spec, verts, faces, R, T, face_uvs=fuv, texture=tex)
res["render_forward_us"] = timeit(rf)
g = torch.randn(N, 3, S, S, device=dev)
res["render_backward_tex_us"] = timeit(lambda: ops.render_backward(state["o"][3], g))
res["render_backward_tex_verts_us"] = timeit(lambda: ops.render_backward(state["o"][3], g, need_verts=True))
for C, HW in ((64, S * S), (128, S * S // 4), (256, S * S // 16), (512, S * S // 64), (512, S * S // 256)):
    f = torch.relu(torch.randn(N, C, HW, device=dev))
    if os.environ.get("NHWC") == "1":
        side = int(HW ** 0.5)
        f = f.reshape(N, C, side, HW // side).contiguous(memory_format=torch.channels_last)
    tgt = torch.randn(1, C, C, device=dev); loss = torch.zeros(1, device=dev)
    dg = torch.randn(N, C, C, device=dev)
    t1 = timeit(lambda: ops.gram_mse_forward(f, tgt, 1e-9, loss))
    t2 = timeit(lambda: ops.gram_backward(f, dg, 1.0))
    t3 = timeit(lambda: ops.gram_forward(f[:1]))
    res[f"gram_mse_fwd_{C}x{HW}_us"] = t1; res[f"gram_bwd_{C}x{HW}_us"] = t2; res[f"gram_fwd_B1_{C}x{HW}_us"] = t3
    res[f"gram_mse_fwd_{C}x{HW}_GBs"] = N * C * HW * 4 / t1 / 1e3; res[f"gram_bwd_{C}x{HW}_GBs"] = 2 * N * C * HW * 4 / t2 / 1e3
a = torch.randn(N, 512, S // 8, S // 8, device=dev); b = torch.randn_like(a); loss = torch.zeros(1, device=dev)
res["mse_us"] = timeit(lambda: ops.mse_forward(a, b, 1e-6, loss))
ops.poll_overflow(block=True)
print(json.dumps({k: round(v, 2) for k, v in res.items()}, indent=1))

"""BASELINE configs[4] scale (GPU box only; one process, or one rank per GPU under torchrun): a ~1.5 M-face mesh (cow subdivided 4x, UVs interpolated at the
midpoints), 16 views x 1024^2 per GPU (128 views over 8 GPUs), texture + vertex optimisation (`both`), VGG-19 on cuDNN.
Prints one JSON line: iterations / s, per-op CUDA-event times, peak memory."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import numpy as np, torch, torchvision
import torch.nn.functional as F
from st3d import ops, cameras as cm
from st3d.optimize import StyleOptimizer

def subdivide(verts, faces):
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], dim=0)
    uniq, inv = torch.unique(torch.sort(e, dim=1).values, dim=0, return_inverse=True)
    mid = 0.5 * (verts[uniq[:, 0]] + verts[uniq[:, 1]])
    V, Fc = verts.shape[0], faces.shape[0]
    m01, m12, m20 = V + inv[:Fc], V + inv[Fc:2 * Fc], V + inv[2 * Fc:]
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    nf = torch.cat([torch.stack([a, m01, m20], 1), torch.stack([m01, b, m12], 1), torch.stack([m20, m12, c], 1),
                    torch.stack([m01, m12, m20], 1)], dim=0)
    return torch.cat([verts, mid], dim=0), nf

views, size, levels = int(os.environ.get("VIEWS", 16)), int(os.environ.get("SIZE", 1024)), int(os.environ.get("SUBDIV", 4))
target, steps = os.environ.get("TARGET", "both"), int(os.environ.get("STEPS", 5))
d = np.load(os.path.join(ROOT, "tests/golden/cow_mesh.npz"))
verts, faces = torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long()
uvs, fuvs = torch.from_numpy(d["verts_uvs"]), torch.from_numpy(d["faces_uvs"]).long()
for _ in range(levels):
    verts, faces = subdivide(verts, faces); uvs, fuvs = subdivide(uvs, fuvs)
tex = torch.from_numpy(d["texture"]).float() / 255.0
tex = F.interpolate(tex.permute(2, 0, 1)[None], size=(size, size), mode="bilinear", align_corners=False)[0].permute(1, 2, 0).contiguous()
style = F.interpolate(torch.rand(1, 3, size // 16, size // 16, generator=torch.Generator().manual_seed(0)), size=(size, size),
                      mode="bicubic", align_corners=False).clamp(0, 1).contiguous()
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:       # views shard over the ranks (weak scaling: `views` per GPU), gradients all-reduced by NCCL
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
vgg = torchvision.models.vgg19(weights=None).features.eval().to(dev)
for p in vgg.parameters():
    p.requires_grad_(False)
R, T = cm.random_view_cameras(views * world, generator=torch.Generator().manual_seed(0))
R, T = R[rank * views:(rank + 1) * views], T[rank * views:(rank + 1) * views]
opt = StyleOptimizer(verts.to(dev), faces.to(dev), vgg, size, verts_uvs=uvs.to(dev), faces_uvs=fuvs.to(dev), texture=tex.to(dev),
                     target=target, lr=0.01, world_size=world)
R, T, style = R.to(dev), T.to(dev), style.to(dev)
for _ in range(2):
    loss = opt.step(R, T, style)
torch.cuda.synchronize(); ops.poll_overflow(block=True)
if world > 1:
    dist.barrier()
ops.start_profile()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(steps):
    loss = opt.step(R, T, style)
e1.record(); torch.cuda.synchronize()
prof = ops.stop_profile()
ms = e0.elapsed_time(e1) / steps
if world > 1:       # device time, max over ranks
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    dist.barrier()
    dist.destroy_process_group()
if rank != 0:
    sys.exit(0)
stages = {op + "_" + "x".join(map(str, key)): round(sum(v) / steps, 3) for (op, key), v in sorted(prof.items())}
print(json.dumps({"workload": f"cow subdivided {levels}x ({faces.shape[0]} faces, {verts.shape[0]} verts), {views} views x {size}^2 "
                              f"per GPU on {world} GPU(s), target={target}", "n_gpus": world, "ms_per_step": ms, "it_per_s": 1e3 / ms,
                  "views_per_s": views * world * 1e3 / ms, "loss": float(loss), "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30,
                  "stages_ms_per_step": stages}))

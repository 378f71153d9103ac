#!/usr/bin/env python
"""bench.py -- style-optimisation iterations/sec on B200 (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libst3d kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), real N-view steps
    python bench.py --impl torch_unfused --steps K ...       # GPU stand-in for the reference's PyTorch3D-CUDA path

One step = one optimisation iteration of second_approach.py:147-190 (`texture` target) over the views a
rank holds: render content mesh -> render current mesh -> VGG-19 features (torch/cuDNN, out of scope but
inside the timed step) -> content MSE + Gram style loss -> backward -> [NCCL all-reduce] -> Adam step.
Weak scaling: every GPU holds `--views` (default 8) views at `--size`^2 (default 512); `value` is the
whole-job throughput in 8-view-iteration equivalents per second (= iterations/sec at N = 1).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "2d-to-3d-style-transfer_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "style-opt iters/sec (render+loss fwd/bwd)"
UNIT = "it/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="st3d", choices=["st3d", "reference", "torch_unfused"])
    ap.add_argument("--views", type=int, default=8, help="views per GPU")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--nchw", action="store_true", help="keep VGG activations NCHW (torch default) instead of channels_last")
    ap.add_argument("--unfused-vgg", action="store_true", help="conv, bias add and ReLU as separate torch kernels")
    ap.add_argument("--cudnn-benchmark", action="store_true",
                    help="let cuDNN time its convolution algorithms per shape (torch.backends.cudnn.benchmark)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the e2e / cached-constants legs")
    ap.add_argument("--no-standin", action="store_true", help="skip the unfused-torch GPU stand-in and the cuBLAS Gram A/B")
    ap.add_argument("--no-scaling-128v", action="store_true", help="skip the fixed-128-view strong-scaling record")
    ap.add_argument("--profile-run", action="store_true",
                    help="for ncu captures only: honour --warmup < 3 (a number from such a run is never reported)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload (shared by all arms): cow mesh, Style_1, seeded random-init VGG-19
# ------------------------------------------------------------------------------------------------
def load_workload(size):
    import numpy as np
    import torch
    import torch.nn.functional as F
    d = np.load(os.path.join(ROOT, "tests", "golden", "cow_mesh.npz"))
    verts = torch.from_numpy(d["verts"]).float()
    faces = torch.from_numpy(d["faces"]).long()
    uvs = torch.from_numpy(d["verts_uvs"]).float()
    fuvs = torch.from_numpy(d["faces_uvs"]).long()
    tex = torch.from_numpy(d["texture"]).float() / 255.0
    # first_approach.py:90-100 / second_approach.py: the texture is resized to size x size (bilinear)
    tex = F.interpolate(tex.permute(2, 0, 1)[None], size=(size, size), mode="bilinear", align_corners=False)[0]
    tex = tex.permute(1, 2, 0).contiguous()
    return dict(verts=verts, faces=faces, verts_uvs=uvs, faces_uvs=fuvs, texture=tex, style=style_image(size))


def style_image(size, name="style_1"):
    """The reference's imgs/Style_1.jpg (second_approach.py:26), from the committed 512^2 copy in
    tests/golden/styles.npz, resized to size x size as utils.py:34-44 does: (1,3,size,size) in [0,1]."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    a = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "styles.npz"))[name]).float() / 255.0
    x = a.permute(2, 0, 1)[None]
    if x.shape[-1] != size:
        x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False)
    return x.contiguous()


DATA = ("cow mesh + Style_1 image from the reference's own assets (tests/golden fixtures; Style_1 stored at 512^2) + "
        "seeded random-init VGG-19 (ImageNet weights unavailable offline)")


def seeded_vgg():
    import torch
    import torchvision
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    for p in vgg.parameters():
        p.requires_grad_(False)
    return vgg


def cameras(n_total):
    import torch
    from st3d import cameras as cm
    return cm.random_view_cameras(n_total, generator=torch.Generator().manual_seed(0))


def workload_config(args, world):
    return {
        "workload": f"cow_mesh texture optimisation, {args.views} views x {args.size}^2 per GPU (BASELINE configs[1])",
        "views_per_gpu": args.views, "global_views": args.views * world, "image_size": args.size,
        "faces": 5856, "verts": 2930, "texture": f"{args.size}x{args.size}x3", "target": "texture",
        "style_weight": 1e6, "content_weight": 1, "lr": 0.01, "parallelism": f"view-sharded dp{world}",
        "vgg": "torchvision VGG-19 .features, seeded random init (ImageNet weights unavailable offline), fp32 cuDNN "
               "(torch default allow_tf32), activations " + ("NCHW" if args.nchw else "channels_last (NHWC)") +
               (", conv+bias+ReLU as cuDNN's fused call, 2x2 max pools on libst3d's NHWC kernels"
                if not args.unfused_vgg else "") + ", inside the timed step",
        "l2": "per-step working set (VGG activations of 8 x 512^2 images, > 4 GB) exceeds the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path (oracle port; PyTorch3D itself is not installable here)
# ------------------------------------------------------------------------------------------------
def cpu_iterations(args, total_views, views_per_step, n_timed, n_warm, budget_s=None):
    """Times optimisation iterations of second_approach.py:147-190 on the host cores, `views_per_step` views each
    out of `total_views` cameras (views_per_step == total_views is the real iteration of the job; 1 is the bounded
    sample of the GPU arm's cpu_baseline).
    Every iteration does what the reference does: content render, style image repeated to the batch size, three VGG
    walks, loss, backward, Adam.  Returns (seconds per step, threads, steps actually timed, warm-up steps run).
    budget_s: if the first step shows that n_warm + n_timed steps would not fit, fewer are run (never fewer than
    1 + 3) and the caller reports the count it got."""
    import torch
    from oracle import loss_oracle as lo
    from oracle import render_oracle as ro
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = load_workload(args.size)
    vgg = seeded_vgg()
    R, T = cameras(total_views)
    tex0 = w["texture"]
    tex = tex0.clone().requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=0.01)
    kw = dict(verts_uvs=w["verts_uvs"], faces_uvs=w["faces_uvs"], nthreads=threads)
    B = views_per_step
    times, i, warm_run = [], 0, 0
    while len(times) < n_timed:
        lo_v = (i * B) % total_views
        idx = [(lo_v + j) % total_views for j in range(B)]
        Rb, Tb = R[idx], T[idx]
        t0 = time.perf_counter()
        opt.zero_grad()
        style = w["style"].repeat(B, 1, 1, 1)                                               # second_approach.py:157
        with torch.no_grad():
            content, _ = ro.images_and_masks(ro.render_views(w["verts"], w["faces"], Rb, Tb, args.size, texture=tex0, **kw))
        current, _ = ro.images_and_masks(ro.render_views(w["verts"], w["faces"], Rb, Tb, args.size, texture=tex, **kw))
        loss = lo.perceptual_loss(current, content, style, vgg, 1e6, 1.0)
        loss.backward()
        opt.step()
        float(loss.detach())
        dt = time.perf_counter() - t0
        if i == 0 and budget_s is not None and dt * (n_warm + n_timed) > budget_s:
            n_warm = 1
            n_timed = max(3, min(n_timed, int(budget_s / dt) - 1))
        if i >= n_warm:
            times.append(dt)
        else:
            warm_run += 1
        i += 1
    return sum(times) / len(times), threads, len(times), warm_run


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_dict(args, sec_per_step, views_per_step, threads, n_timed, n_warm):
    """value: 8-view-iteration equivalents per second (the unit of the GPU arm's `value`)."""
    import torch
    sec_iter = sec_per_step * args.views / views_per_step
    if views_per_step >= args.views:
        sample = f"{n_timed} timed + {n_warm} warm-up FULL iterations ({views_per_step} views x {args.size}^2 each)"
    else:
        sample = (f"{n_timed} timed + {n_warm} warm-up iterations of {views_per_step} of the {args.views} views at {args.size}^2 "
                  f"(bounded sample; every stage of the iteration is per-view, so a full iteration costs "
                  f"{args.views // views_per_step}x; --impl reference times full iterations)")
    return {"value": 1.0 / sec_iter, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": sample + "; CPU restatement of the PyTorch3D path (library unavailable) + the reference's loss "
                               "arithmetic, compute only (no PNG / log writes), all host threads",
            "sec_per_step": sec_per_step, "views_per_step": views_per_step, "cpu_model": cpu_model(),
            "os_cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}


def run_reference(args):
    """The reference's CPU path on the box's host cores: every step is a REAL iteration over all args.views views
    (ms_per_step x steps is the time this arm actually spent in its timed region)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    budget = float(os.environ.get("ST3D_REF_BUDGET_S", "480"))
    # the GPU arm's job at --gpus N is N x args.views views per iteration (weak scaling): the same job here
    world = max(1, args.gpus)
    total = args.views * world
    sec, threads, n_timed, n_warm = cpu_iterations(args, total, total, max(args.steps, 1), max(args.warmup, 0), budget)
    value = world / sec                 # args.views-view-iteration equivalents per second, like the GPU arm
    cb = cpu_baseline_dict(args, sec, total, threads, n_timed, n_warm)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n_timed,
        "warmup": n_warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": DATA,
        "config": workload_config(args, world), "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if n_timed != args.steps or n_warm != args.warmup:
        out["steps_requested"], out["warmup_requested"] = args.steps, args.warmup
        out["note"] = (f"the first iteration showed that {args.warmup}+{args.steps} full iterations would exceed the "
                       f"{budget:.0f} s budget of this arm (ST3D_REF_BUDGET_S): `steps`/`warmup` are what was run")
    return json.dumps(out)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
# filled from the round's ncu capture (profiles/): (bytes per launch, source) of the fused Gram backward at conv1_1
TRAFFIC_FUSED_BWD64 = (1567838400, "profiles/r2_ncu_full.csv (end-of-round capture, timed step): k_gram_tc_bwd<64, NHWC, RING> "
                                   "with ST3D_GRAM_ACCUMULATE | ST3D_GRAM_RELU_MASK | ST3D_GRAM_DGRAM_SYMMETRIC, dram read 1074.97 MB + "
                                   "write 492.87 MB per launch (the mask's second read of F hits L2)")


def algorithmic_bytes(op, key, tex=512):
    """SURVEY.md section 8(d) / DESIGN.md: bytes one call of `op` must move."""
    if op == "render_forward":      # face records + (pix_to_face i32, rgb, mask) per pixel + texture once
        N, H, W, F = key
        return N * F * 36 + N * H * W * (4 + 16) + tex * tex * 12
    if op == "render_backward":     # grad rgb + saved pix_to_face per pixel + face records + grad_texture
        N, H, W, F = key
        return N * H * W * (12 + 4) + N * F * 48 + tex * tex * 12
    if op.startswith("gram_backward"):   # read F, write dF; fused tail: + read the incoming gradient (the ReLU mask re-reads
        B, C, HW = key                   # F, which an ideal kernel would still hold: no extra algorithmic bytes)
        return (2 + ("_acc" in op)) * B * C * HW * 4
    if op in ("gram_forward", "gram_mse_forward"):   # read F, write G (+dG)
        B, C, HW = key
        return B * C * HW * 4 + B * C * C * 4
    if op == "mse_forward":         # read a, b, write grad
        return 3 * key[0] * 4
    if op == "mse_value_forward":   # the content tap's forward: read a, b; its gradient comes from mse_tap_backward
        return 2 * key[0] * 4
    if op == "mse_tap_backward":    # read y, c and the incoming gradient, write the layer's pre-activation gradient
        return 4 * key[0] * 4
    if op == "maxpool_forward":     # read x, write x / 4
        B, C, H, W = key
        return B * C * H * W * 5
    if op == "maxpool_backward":    # read x and grad_y (x / 4), write grad_x
        B, C, H, W = key
        return B * C * H * W * 9
    return None


# ------------------------------------------------------------------------------------------------
# GPU stand-in for the reference's PyTorch3D-CUDA path: the reference's own loop structure
# on the GPU with NOTHING fused -- one renderer call per view (utils.py:68-76), Fragments -> interpolation ->
# grid_sample -> elementwise blend as separate launches, stock torchvision VGG (NCHW, separate ReLUs), torch.bmm
# Gram + sub / pow / mean, autograd backward, stock Adam.  PyTorch3D itself cannot be installed here, so the
# Fragments come from this library's operator-boundary kernel (the `_C.rasterize_meshes` equivalent) -- which is
# generous to the comparator; everything after it is stock torch / cuBLAS / cuDNN.
# ------------------------------------------------------------------------------------------------
def _torch_features(x, vgg):
    taps = {"0": "conv1_1", "5": "conv2_1", "10": "conv3_1", "19": "conv4_1", "21": "conv4_2", "28": "conv5_1"}
    out = {}
    for name, layer in vgg._modules.items():        # all 37 modules, as style_transfer.py:21-26 does
        x = layer(x)
        if name in taps:
            out[taps[name]] = x
    return out


def _torch_gram(t):
    import torch
    b, c, h, w = t.shape
    f = t.view(b, c, h * w)
    return torch.bmm(f, f.transpose(1, 2))


def _torch_perceptual(current, content, style, vgg, style_weight=1e6, content_weight=1.0):
    """losses.py:12-44 in stock torch ops."""
    import torch
    content_f = _torch_features(content, vgg)["conv4_2"]
    style_f = _torch_features(style, vgg)
    grams = {k: _torch_gram(v) for k, v in style_f.items() if k != "conv4_2"}
    cur = _torch_features(current, vgg)
    loss = content_weight * torch.mean((cur["conv4_2"] - content_f) ** 2)
    s = 0
    for k, g in grams.items():
        f = cur[k]
        s = s + torch.mean((_torch_gram(f) - g) ** 2) / (f.shape[1] ** 2 * f.shape[2] ** 2)
    return loss + style_weight * s


def torch_unfused_arm(args, dev, n_timed, n_warm=3):
    """-> dict(ms_per_step, it_per_s, ...) of the unfused stand-in on configs[1]."""
    import torch
    compat = os.path.join(PKG, "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    from pytorch3d.renderer import (AmbientLights, FoVPerspectiveCameras, MeshRasterizer, RasterizationSettings,
                                    SoftPhongShader, TexturesUV)
    from pytorch3d.structures import Meshes
    w = load_workload(args.size)
    vgg = seeded_vgg().to(dev)                      # stock module: NCHW, Conv2d / ReLU(inplace) / MaxPool2d
    R, T = cameras(args.views)
    cams = FoVPerspectiveCameras(R=R.to(dev), T=T.to(dev), device=dev)
    cam_list = [cams[i] for i in range(args.views)]
    rasterizer = MeshRasterizer(cameras=FoVPerspectiveCameras(device=dev),
                                raster_settings=RasterizationSettings(image_size=args.size, blur_radius=0.0, faces_per_pixel=1))
    shader = SoftPhongShader(device=dev, lights=AmbientLights(device=dev))
    verts, faces = w["verts"].to(dev), w["faces"].to(dev)
    uvs, fuvs = w["verts_uvs"][None].to(dev), w["faces_uvs"][None].to(dev)
    tex0 = w["texture"][None].to(dev)
    tex = tex0.clone().requires_grad_(True)
    style1 = w["style"].to(dev)
    opt = torch.optim.Adam([tex], lr=0.01)

    def render(texture):
        mesh = Meshes(verts=[verts], faces=[faces], textures=TexturesUV(verts_uvs=uvs, faces_uvs=fuvs, maps=texture))
        imgs, masks = [], []
        for cam in cam_list:                         # utils.py:68-76: one renderer call per view
            frags = rasterizer(mesh, cameras=cam)
            rgba = shader(frags, mesh, cameras=cam)  # Fragments -> interp -> grid_sample -> blend, unfused
            imgs.append(rgba[0, ..., :3].permute(2, 0, 1))
            masks.append((rgba[0, ..., 3] > 0).float().unsqueeze(0))
        return torch.stack(imgs, dim=0), torch.stack(masks, dim=0)

    def step():
        opt.zero_grad()
        style = style1.repeat(args.views, 1, 1, 1)                      # second_approach.py:157
        with torch.no_grad():
            content, _ = render(tex0)
        current, _ = render(tex)
        loss = _torch_perceptual(current, content, style, vgg)
        loss.backward()
        opt.step()
        return loss

    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_timed):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_timed
    return {"ms_per_step": ms, "it_per_s": 1e3 / ms, "steps": n_timed, "warmup": n_warm, "final_loss": float(loss),
            "label": "STAND-IN for the reference's PyTorch3D-CUDA path (PyTorch3D is not installable here): per-view renderer "
                     "calls, Fragments from libst3d's operator-boundary rasterizer, then unfused torch ops (interpolation, "
                     "grid_sample, elementwise blend), stock torchvision VGG-19 (NCHW), torch.bmm Gram + sub/pow/mean, "
                     "autograd backward, stock Adam; device-resident, torch default precision (fp32 matmul, cuDNN TF32 allowed)"}


def run_torch_unfused(args):
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    if not torch.cuda.is_available():
        raise RuntimeError("--impl torch_unfused needs a CUDA device")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    r = torch_unfused_arm(args, dev, max(args.steps, 1), max(args.warmup, 3))
    return json.dumps({"impl": "torch_unfused", "metric": METRIC, "value": r["it_per_s"], "unit": UNIT, "n_gpus": 1,
                       "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                       "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA,
                       "config": workload_config(args, 1), "label": r["label"], "final_loss": r["final_loss"]})


def gram_vs_cublas(args, dev, reps=20):
    """Per style layer at the bench's shapes: this library's Gram + MSE forward and Gram backward (tcgen05 kind::tf32,
    channels_last features read in place) against what the reference runs -- torch.bmm + sub/pow/mean forward,
    autograd (two more cuBLAS GEMMs + elementwise) backward -- in fp32 and with allow_tf32."""
    import torch
    from st3d import ops
    B, S = args.views, args.size
    shapes = [("conv1_1", 64, S), ("conv2_1", 128, S // 2), ("conv3_1", 256, S // 4), ("conv4_1", 512, S // 8),
              ("conv5_1", 512, S // 16)]
    out = {}
    prev = torch.backends.cuda.matmul.allow_tf32
    gen = torch.Generator(device=dev).manual_seed(0)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    try:
        for name, C, H in shapes:
            feat = torch.relu(torch.randn(B, C, H, H, device=dev, generator=gen))
            target = torch.randn(1, C, C, device=dev, generator=gen)
            f_cl = feat.contiguous(memory_format=torch.channels_last)
            scale = 1.0 / (B * C * C) / (float(C) ** 2 * float(H) ** 2)
            loss = torch.zeros(1, device=dev)
            grad_out = torch.empty_like(f_cl)

            def ours():
                dgram, _ = ops.gram_mse_forward(f_cl, target, scale, loss, precision=args.precision)
                ops.gram_backward(f_cl, dgram, 1.0, out=grad_out, precision=args.precision)

            leaf = feat.clone().requires_grad_(True)

            def theirs():
                leaf.grad = None
                g = _torch_gram(leaf)
                l = torch.mean((g - target) ** 2) / (C ** 2 * H ** 2)
                l.backward()

            row = {"B": B, "C": C, "HW": H * H, "st3d_ms": round(timed(ours), 4)}
            torch.backends.cuda.matmul.allow_tf32 = False
            row["torch_fp32_ms"] = round(timed(theirs), 4)
            torch.backends.cuda.matmul.allow_tf32 = True
            row["torch_tf32_ms"] = round(timed(theirs), 4)
            row["speedup_vs_fp32"] = round(row["torch_fp32_ms"] / row["st3d_ms"], 2)
            row["speedup_vs_tf32"] = round(row["torch_tf32_ms"] / row["st3d_ms"], 2)
            out[name] = row
            del feat, f_cl, leaf, grad_out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    tot = {k: round(sum(r[k] for r in out.values()), 4) for k in ("st3d_ms", "torch_fp32_ms", "torch_tf32_ms")}
    out["all_layers"] = dict(tot, speedup_vs_fp32=round(tot["torch_fp32_ms"] / tot["st3d_ms"], 2),
                             speedup_vs_tf32=round(tot["torch_tf32_ms"] / tot["st3d_ms"], 2))
    out["what"] = ("style-loss body of one layer, forward + backward into the feature map: st3d_gram_mse_forward + "
                   "st3d_gram_backward (" + args.precision + ") vs torch.bmm + sub/pow/mean + autograd (style_transfer.py:31-35, "
                   "losses.py:35-39); CUDA events, " + str(reps) + " repetitions after 3 warm-ups")
    return out


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: ~1.5 M faces, a FIXED total of 128 views x 1024^2 split over the ranks (strong scaling)
# ------------------------------------------------------------------------------------------------
def scaling_128v(dev, world, rank, vgg, precision, total_views=128, size=1024, levels=4, micro_batch=16, steps=3, warm=2):
    """One optimisation iteration = all 128 views into ONE Adam step: every rank renders its 128 / world views in
    micro-batches of 16 (gradients accumulate), then one NCCL all-reduce of the flat texture gradient.  At world = 1
    that is 8 micro-batches on one GPU; at world = 8 one micro-batch per GPU: the same job, so the per-N times give
    strong-scaling efficiency directly."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from st3d import meshgen, ops
    from st3d.optimize import StyleOptimizer
    if total_views % world:
        return {"skipped": f"{total_views} views do not split over {world} ranks"}
    d = np.load(os.path.join(ROOT, "tests", "golden", "cow_mesh.npz"))
    verts, faces, uvs, fuvs = meshgen.subdivided_uv_mesh(torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long(),
                                                         torch.from_numpy(d["verts_uvs"]),
                                                         torch.from_numpy(d["faces_uvs"]).long(), levels)
    tex = torch.from_numpy(d["texture"]).float() / 255.0
    tex = F.interpolate(tex.permute(2, 0, 1)[None], size=(size, size), mode="bilinear", align_corners=False)[0]
    tex = tex.permute(1, 2, 0).contiguous()
    style = style_image(size).to(dev)
    per_rank = total_views // world
    R, T = cameras(total_views)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    R, T = R[sl].contiguous().to(dev), T[sl].contiguous().to(dev)
    opt = StyleOptimizer(verts.to(dev), faces.to(dev), vgg, size, verts_uvs=uvs.to(dev), faces_uvs=fuvs.to(dev),
                         texture=tex.to(dev), target="texture", lr=0.01, precision=precision, world_size=world)
    mb = min(micro_batch, per_rank)
    for _ in range(warm):
        opt.step(R, T, style, micro_batch=mb)
    torch.cuda.synchronize()
    ops.poll_overflow(block=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = opt.step(R, T, style, micro_batch=mb)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ar_ms = None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # the collective alone (the flat texture gradient, in place), CUDA events, max over ranks
        buf = torch.zeros_like(opt._flat_grad)
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1) / 10], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
    peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    out = {"workload": f"cow subdivided {levels}x ({faces.shape[0]} faces, {verts.shape[0]} verts), {total_views} views x "
                       f"{size}^2 in total, texture {size}^2, Style_1, texture target, one Adam step per iteration",
           "scaling": "strong", "n_gpus": world, "views_total": total_views, "views_per_gpu": per_rank, "micro_batch": mb,
           "micro_batches_per_gpu": -(-per_rank // mb), "steps": steps, "warmup": warm,
           "ms_per_iteration": ms, "it_per_s": 1e3 / ms, "views_per_s": total_views * 1e3 / ms,
           "allreduce_ms": ar_ms, "allreduce_bytes": int(opt._flat_grad.numel() * 4), "peak_mem_GB": round(peak, 2),
           "final_loss": float(loss),
           "note": "strong-scaling efficiency at N = ms_per_iteration(N=1) / (N * ms_per_iteration(N)); the only collective "
                   "is the in-place all-reduce of the flat texture gradient (allreduce_ms: that collective alone)"}
    del opt
    return out


def nst_2d_loop(dev, vgg, size=512, batch=4, steps=50):
    """One step of the 2D neural-style-transfer loop (`style_transfer()`, style_transfer.py:59-83; first stage of
    first_approach.py:171-179) exactly as the drop-in compat/style_transfer.py runs it -- fused loss walk, Adam -- launched
    from Python every step against replayed from one CUDA graph (st3d.optimize.CapturedIteration); CUDA events."""
    import torch
    from st3d import losses
    from st3d.optimize import CapturedIteration
    from st3d.vgg import fuse_vgg_features
    model = fuse_vgg_features(vgg, channels_last=True)
    g = torch.Generator().manual_seed(5)
    content = torch.rand(batch, 3, size, size, generator=g).to(dev)
    style = style_image(size).to(dev).repeat(batch, 1, 1, 1)
    with torch.no_grad():
        content_feat = losses.get_features(content, model, {"21": losses.CONTENT_LAYER})[losses.CONTENT_LAYER]
    grams = losses.style_targets(style, model)
    out = {}
    for label in ("eager", "graphed"):
        images = content.clone().requires_grad_(True)
        optimizer = torch.optim.Adam([images], lr=0.01, capturable=True, fused=True)

        def iteration():
            loss = losses.perceptual_loss_of_images(images, model, content_feat, grams, 1e6, 1.0)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()

        if label == "graphed":
            step = CapturedIteration(iteration, dev, warmup=3).replay
        else:
            for _ in range(3):
                iteration()
            step = iteration
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[label + "_ms_per_step"] = e0.elapsed_time(e1) / steps
    out["workload"] = f"2D NST step on {batch} image(s) x {size}^2 (VGG-19 forward + backward on cuDNN inside the step), {steps} steps"
    out["graph_speedup"] = out["eager_ms_per_step"] / out["graphed_ms_per_step"]
    return out


def c3_both_target(dev, vgg, precision, views=4, size=512, steps=20):
    """BASELINE configs[2], one rank's share: bob (10 688 faces, the stand-in for the missing bunny.obj), target `both`
    (texture + vertices, second_approach.py:140-190 with losses.py:101-126: perceptual loss x main_loss_weight + vertex MSE
    + edge / Laplacian / normal-consistency regularisers), `views` views x `size`^2.  One iteration launched from Python
    against the same iteration with its gradient computation replayed from one CUDA graph; CUDA events.  The regularisers
    are one libst3d forward and one backward launch (csrc/mesh_reg.cu); `regularisers_torch_ms` times the ~80-launch
    torch formulation of the same three terms beside them."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from st3d import mesh_losses as ml
    from st3d.optimize import StyleOptimizer
    d = np.load(os.path.join(ROOT, "tests", "golden", "bob_mesh.npz"))
    verts, faces = torch.from_numpy(d["verts"]).float().to(dev), torch.from_numpy(d["faces"]).long().to(dev)
    uvs, fuvs = torch.from_numpy(d["verts_uvs"]).float().to(dev), torch.from_numpy(d["faces_uvs"]).long().to(dev)
    tex = torch.from_numpy(d["texture"]).float() / 255.0
    tex = F.interpolate(tex.permute(2, 0, 1)[None], size=(size, size), mode="bilinear", align_corners=False)[0]
    tex = tex.permute(1, 2, 0).contiguous().to(dev)
    R, T = cameras(views)
    R, T, style = R.to(dev), T.to(dev), style_image(size).to(dev)
    out = {"workload": f"bob_mesh `both` target, {views} views x {size}^2 (BASELINE configs[2], one rank of 8), {steps} steps"}

    def fresh():
        return StyleOptimizer(verts, faces, vgg, size, verts_uvs=uvs, faces_uvs=fuvs, texture=tex, target="both",
                              precision=precision)
    for label in ("eager", "captured"):
        opt = fresh()
        if label == "captured":
            opt.capture(R, T, style, warmup=3)
            step = opt.step_captured
        else:
            for _ in range(3):
                opt.step(R, T, style)
            step = lambda: opt.step(R, T, style)    # noqa: E731
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        out[label + "_ms_per_step"] = e0.elapsed_time(e1) / steps
        out[label + "_it_per_s"] = 1e3 * steps / e0.elapsed_time(e1)
        out[label + "_final_loss"] = float(loss)
        del opt
    v = verts.clone().requires_grad_(True)
    w3 = torch.ones(3, device=dev)

    def fused():
        v.grad = None
        torch.dot(ml.regularizers(v, faces), w3).backward()

    def torch_ops():
        v.grad = None
        (ml.edge_loss_torch(v, faces) + ml.laplacian_smoothing_torch(v, faces) + ml.normal_consistency_torch(v, faces)).backward()
    for label, fn in (("regularisers_fused_ms", fused), ("regularisers_torch_ms", torch_ops)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[label] = e0.elapsed_time(e1) / 20
    return out


def c4_teapot_multi_style(dev, vgg, precision, views=8, size=1024, steps=10):
    """BASELINE configs[3], one rank's share: teapot (2 464 faces, no UVs -> per-vertex colours initialised at 0.5, the
    colours are what is optimised), 8 of the 64 views x 1024^2, style target Gs = sum_j 1/4 Gram(style_j) over the four
    style images the reference ships (Style_1/3/4/5; stored at 512^2 / 256^2 in tests/golden/styles.npz and resized as
    utils.py:34-44 does).  One iteration, eager and with the captured gradient computation; CUDA events."""
    import numpy as np
    import torch
    from st3d.optimize import StyleOptimizer
    d = np.load(os.path.join(ROOT, "tests", "golden", "teapot_mesh.npz"))
    verts, faces = torch.from_numpy(d["verts"]).float().to(dev), torch.from_numpy(d["faces"]).long().to(dev)
    rgb = torch.full((verts.shape[0], 3), 0.5, device=dev)
    styles = torch.cat([style_image(size, n) for n in ("style_1", "style_3", "style_4", "style_5")], dim=0).to(dev)
    R, T = cameras(views)
    R, T = R.to(dev), T.to(dev)
    out = {"workload": f"teapot_mesh per-vertex colours, {views} views x {size}^2, four blended style Grams "
                       f"(BASELINE configs[3], one rank of 8), {steps} steps"}
    for label in ("eager", "captured"):
        opt = StyleOptimizer(verts, faces, vgg, size, verts_rgb=rgb, target="texture", precision=precision,
                             style_weights=[0.25] * 4)
        if label == "captured":
            opt.capture(R, T, styles, warmup=3)
            step = opt.step_captured
        else:
            for _ in range(3):
                opt.step(R, T, styles)
            step = lambda: opt.step(R, T, styles)    # noqa: E731
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        out[label + "_ms_per_step"] = e0.elapsed_time(e1) / steps
        out[label + "_views_per_s"] = 1e3 * steps * views / e0.elapsed_time(e1)
        out[label + "_final_loss"] = float(loss)
        del opt
        gc.collect()
        torch.cuda.empty_cache()
    return out


def c1_first_approach(dev, steps=200, size=256, run_cpu=True):
    """BASELINE configs[0]: cow texture fit, 1 view 256x256, 50 Adam steps of the masked-MSE loop of
    first_approach.py:191-213 (render -> masked MSE -> backward -> Adam; no VGG on this path)."""
    import torch
    from st3d import functional as Fn
    w = load_workload(size)
    R, T = cameras(1)
    target = torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(3))
    verts, faces = w["verts"].to(dev), w["faces"].int().to(dev)
    fuv = w["verts_uvs"][w["faces_uvs"]].to(dev)
    tex = w["texture"].to(dev).clone().requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=0.01)
    Rd, Td, tgt = R.to(dev), T.to(dev), target.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        img, mask, _ = Fn.render_views(verts, faces, Rd, Td, size, texture=tex, face_uvs=fuv)
        loss = Fn.masked_mse_loss(img, tgt, mask)
        loss.backward()
        opt.step()
        return loss

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / steps
    res = {"workload": "cow texture fit, 1 view x 256^2, masked-MSE loop (first_approach.py:191-213), 200 timed Adam steps",
           "gpu_ms_per_step": gpu_ms, "gpu_it_per_s": 1e3 / gpu_ms, "final_loss": float(loss.detach())}
    # the same loop with the iteration captured in one CUDA graph (st3d.optimize.GraphedTextureFit): at this size the
    # eager loop is bound by the host's launch rate, the replayed graph by the kernels
    from st3d.optimize import GraphedTextureFit
    fit = GraphedTextureFit(verts, faces, w["verts_uvs"].to(dev), w["faces_uvs"].to(dev), w["texture"].to(dev), Rd, Td, tgt,
                            size, lr=0.01, warmup=20)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        gl = fit.step()
    g1.record()
    torch.cuda.synchronize()
    graph_ms = g0.elapsed_time(g1) / steps
    res.update({"graph_ms_per_step": graph_ms, "graph_it_per_s": 1e3 / graph_ms, "graph_final_loss": float(gl)})
    if run_cpu:
        from oracle import loss_oracle as lo
        from oracle import render_oracle as ro
        torch.set_num_threads(os.cpu_count() or 1)
        texc = w["texture"].clone().requires_grad_(True)
        optc = torch.optim.Adam([texc], lr=0.01)
        threads = os.cpu_count() or 1
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            optc.zero_grad()
            img, mask = ro.images_and_masks(ro.render_views(w["verts"], w["faces"], R, T, size, texture=texc,
                                                            verts_uvs=w["verts_uvs"], faces_uvs=w["faces_uvs"],
                                                            nthreads=threads))
            l = lo.first_approach_loss(img, mask, target, None, None, None, {}, "texture")
            l.backward()
            optc.step()
            if i:
                ts.append(time.perf_counter() - t0)
        cpu_ms = 1e3 * sum(ts) / len(ts)
        res.update({"cpu_ms_per_step": cpu_ms, "cpu_it_per_s": 1e3 / cpu_ms, "cpu_threads": threads,
                    "speedup": cpu_ms / gpu_ms, "graph_speedup": cpu_ms / graph_ms})
    return res


def bind_to_gpu_numa_node(index):
    """One process per GPU: keep the rank's host threads (Python launch loop, pinned-buffer copies) on the CPUs NVML
    names as local to its GPU, so that 8 ranks do not share one NUMA node's cores and memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:       # no NVML / not permitted: run unbound
        return None


def run_st3d(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if args.cudnn_benchmark:
        torch.backends.cudnn.benchmark = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import st3d
    from st3d import ops
    from st3d.optimize import TextureStyleOptimizer

    w = load_workload(args.size)
    vgg = seeded_vgg().to(dev)
    R_all, T_all = cameras(args.views * world)
    sl = slice(rank * args.views, (rank + 1) * args.views)
    R, T = R_all[sl].contiguous().to(dev), T_all[sl].contiguous().to(dev)
    style = w["style"].to(dev)

    def make(cache):
        return TextureStyleOptimizer(w["verts"].to(dev), w["faces"].to(dev), w["verts_uvs"].to(dev),
                                     w["faces_uvs"].to(dev), w["texture"].to(dev), vgg, args.size, lr=0.01,
                                     precision=args.precision, cache_constants=cache, world_size=world,
                                     channels_last=not args.nchw, fuse_conv_relu=not args.unfused_vgg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timed region --------------------------------------------------------------
    opt = make(False)
    warm = args.warmup if args.profile_run else max(args.warmup, 3)
    for _ in range(warm):
        opt.step(R, T, style)
    ops.poll_overflow(block=True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launch_count()
    ops.start_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = opt.step(R, T, style)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof = ops.stop_profile()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ops.poll_overflow(block=True)
    ms_step = ms_total / args.steps
    value = world / (ms_step * 1e-3)          # 8-view-iteration equivalents per second over all ranks
    final_loss = float(loss)

    # per-op breakdown (rank 0)
    stages, mine_ms, pool_ms = {}, 0.0, 0.0
    for (op, key), v in sorted(prof.items()):
        per_step = sum(v) / args.steps
        if op.startswith("maxpool"):    # VGG-side kernels of this library: reported apart from the render + loss path
            pool_ms += per_step
        else:
            mine_ms += per_step
        name = op + ("" if key is None else "_" + "x".join(str(k) for k in key))
        stages[name] = round(per_step, 4)
    # dominant KERNEL -> roofline.  gram_* / mse ops are one hot kernel each (plus a <5 us symmetrise/finalize);
    # render_forward is a sequence of 7 launches (bins, z-buffer, resolve) and is reported per op below.
    single_kernel_ops = ("gram_backward", "gram_backward_acc", "gram_backward_relu", "gram_backward_acc_relu",
                         "gram_mse_forward", "gram_forward", "mse_forward")
    dom = max(((op, key, sum(v) / len(v)) for (op, key), v in prof.items() if op in single_kernel_ops),
              key=lambda t: t[2], default=None)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    if dom:
        op, key, avg_ms = dom
        nbytes = algorithmic_bytes(op, key)
        ach = nbytes / (avg_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed `ncu --set full` captures
        traffic = {
            ("gram_backward", (8, 64, 262144)): (1024600000, "profiles/r1b_ncu_full.csv: k_gram_tc_bwd<64, NHWC> read 537.0 MB + "
                                                             "write 487.6 MB"),
            ("gram_backward_acc_relu", (8, 64, 262144)): TRAFFIC_FUSED_BWD64,
        }.get((op, key), (None, None))
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic[0], "traffic_source": traffic[1],
                    "kernel": f"{op} {key}: " + ("k_gram_tc_bwd<C, NHWC> (tcgen05 kind::tf32, TMA, TMEM; epilogue fused with the "
                                                 "gradient accumulation and the ReLU backward of the tapped layer)"
                                                 if op == "gram_backward_acc_relu"
                                                 else "k_gram_tc_bwd<C> (tcgen05 kind::tf32, TMA, TMEM)" if op.startswith("gram_backward")
                                                 else "k_gram_tc_fwd<C> (tcgen05 kind::tf32, TMA, TMEM)" if op.startswith("gram")
                                                 else "k_mse"),
                    "algorithmic_bytes": nbytes, "avg_ms": avg_ms,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}

    op_table = {}
    for (op, key), v in sorted(prof.items()):
        nb = algorithmic_bytes(op, key)
        if nb:
            avg = sum(v) / len(v)
            op_table[op + "_" + "x".join(str(k) for k in key)] = {
                "avg_ms": round(avg, 4), "calls_per_step": len(v) // args.steps, "algorithmic_MB": round(nb / 1e6, 2),
                "GBps": round(nb / (avg * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(nb / (avg * 1e-3) / 1e9 / hbm_peak, 3)}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (Gram products tf32 on tcgen05, fp32 accumulate)" if args.precision == "tf32" else "f32",
        "data": DATA,
        "config": workload_config(args, world), "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "final_loss": final_loss,
        "stages_ms_per_step": stages, "ops_vs_hbm_roofline": op_table,
        "render_loss_only": {"ms_per_step": mine_ms, "it_per_s": (world / (mine_ms * 1e-3)) if mine_ms > 0 else None,
                             "note": "sum of libst3d render / Gram / MSE op times (CUDA events) per step; VGG-19 / Adam / "
                                     "autograd glue excluded"},
        "vgg_pool_ms_per_step": pool_ms,
    }

    if not args.no_extras:
        # ---- end-to-end: host buffers in, loss + rendered views out, every step --------------------------
        opt2 = make(False)
        h_style = w["style"].pin_memory()
        h_R, h_T = R_all[sl].contiguous().pin_memory(), T_all[sl].contiguous().pin_memory()
        h_img = torch.empty((args.views, 3, args.size, args.size), dtype=torch.float32).pin_memory()
        d_style, d_R, d_T = torch.empty_like(style), torch.empty_like(R), torch.empty_like(T)

        def e2e_step():
            d_style.copy_(h_style, non_blocking=True)
            d_R.copy_(h_R, non_blocking=True)
            d_T.copy_(h_T, non_blocking=True)
            l = opt2.step(d_R, d_T, d_style, images_out=h_img)   # the reference dumps every view every step; the
            loss = float(l)                                      # device->host copy runs beside the VGG passes
            opt2.images_ready.synchronize()                      # loss.item(): second_approach.py:190
            return loss

        def time_e2e(step_fn):
            for _ in range(3):
                step_fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_fn()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps

        eager_ms = time_e2e(e2e_step)
        # the same iteration with this rank's share (renders, VGG walks, losses, backward, export of the rendered views)
        # replayed from ONE CUDA graph: StyleOptimizer.capture() / step_captured(); all-reduce and Adam stay eager
        opt2 = None
        opt4 = make(False)                          # (capture() wants an optimiser that has not stepped eagerly)
        opt4.capture(d_R, d_T, d_style, images_out=h_img)

        def e2e_step_captured():
            d_style.copy_(h_style, non_blocking=True)
            d_R.copy_(h_R, non_blocking=True)
            d_T.copy_(h_T, non_blocking=True)
            return float(opt4.step_captured())      # the host read of the loss also waits for the exported views

        e2e_ms = time_e2e(e2e_step_captured)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for _ in range(args.steps):
            opt4.step_captured()
        g1.record()
        barrier()
        graph_ms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
        h2d = h_style.numel() * 4 + h_R.numel() * 4 + h_T.numel() * 4
        d2h = h_img.numel() * 4 + 4
        out["e2e"] = {"value": world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                      "api": "st3d.optimize.TextureStyleOptimizer.capture(R, T, style, images_out=pinned) once, then per step: "
                             "pinned host style image + cameras copied into the static device tensors, step_captured() (one "
                             "CUDA graph replay + all-reduce + Adam), the loss read on the host; the rendered views land in "
                             "the pinned buffer inside the same step",
                      "eager": {"value": world / (eager_ms * 1e-3), "ms_per_step": eager_ms,
                                "api": "the same with TextureStyleOptimizer.step(R, T, style, images_out=pinned): every kernel "
                                       "launched from Python"},
                      "device_resident_captured": {"value": world / (graph_ms * 1e-3), "ms_per_step": graph_ms}}
        # ---- variant: constant content/style features cached (loop hygiene the reference lacks) ---------
        opt3 = make(True)
        for _ in range(3):
            opt3.step(R, T, style)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(args.steps):
            opt3.step(R, T, style)
        c1.record()
        barrier()
        cms = max_over_ranks(c0.elapsed_time(c1)) / args.steps
        out["cached_constants"] = {"value": world / (cms * 1e-3), "unit": UNIT, "ms_per_step": cms,
                                   "note": "content render + content/style VGG features computed once (they are "
                                           "constants of the loop); not the headline"}

    import gc
    opt = opt2 = opt3 = opt4 = None       # release the 8 x 512^2 iteration state before the other workloads
    gc.collect()
    torch.cuda.empty_cache()
    if affinity is not None:
        out["host_cpus_per_rank"] = affinity
    if not args.no_scaling_128v:
        # BASELINE configs[4] / north_star "scaling efficiency from 1 to 8 GPUs at 128 views": every rank takes part
        rec = scaling_128v(dev, world, rank, vgg, args.precision)
        out["scaling_128v"] = rec
        gc.collect()
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_standin:
        out["gpu_standin"] = torch_unfused_arm(args, dev, min(max(args.steps, 1), 10))
        out["gpu_standin"]["st3d_speedup"] = out["gpu_standin"]["ms_per_step"] / ms_step
        gc.collect()
        torch.cuda.empty_cache()
        out["gram_vs_cublas"] = gram_vs_cublas(args, dev)
        gc.collect()
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras:
        out["c1_first_approach"] = c1_first_approach(dev, run_cpu=not args.no_cpu_baseline)
        # BASELINE configs[0] is first_approach.py at 1 view x 256^2: its 2D style-transfer stage, and a 4 x 512^2 batch
        out["nst_2d_loop"] = {"1x256": nst_2d_loop(dev, vgg, size=256, batch=1), "4x512": nst_2d_loop(dev, vgg, size=512, batch=4)}
        gc.collect()
        torch.cuda.empty_cache()
        for key, fn in (("c3_both_target", c3_both_target), ("c4_teapot_multi_style", c4_teapot_multi_style)):
            try:
                out[key] = fn(dev, vgg, args.precision)
            except Exception as e:      # a side record must not take the headline line down with it
                out[key] = {"error": repr(e)[:400]}
            gc.collect()
            torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # >= 3 warm-up + >= 10 timed iterations; one view each keeps it to ~30 s of CPU work
        sec, threads, n_timed, n_warm = cpu_iterations(args, args.views, 1, 10, 3)
        out["cpu_baseline"] = cpu_baseline_dict(args, sec, 1, threads, n_timed, n_warm)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        out["lib"] = st3d.library_path()
        return json.dumps(out)
    return None


class _QuietStdout:
    """Routes fd 1 to stderr while the benchmark runs (NCCL / cuDNN may print to stdout) and restores it for
    the single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    args = parse_args()
    with _QuietStdout():
        line = {"reference": run_reference, "torch_unfused": run_torch_unfused, "st3d": run_st3d}[args.impl](args)
    if line is not None:
        print(line, flush=True)


if __name__ == "__main__":
    main()

"""Wall time per epoch of the reference's UNCHANGED second_approach.py (tests/golden/reference_scripts/) run through
`python -m st3d.run` on the GPU box: 8 views x 512^2 in one batch (BASELINE configs[1]), everything the script does per
epoch included (style image load, two render calls, loss, PNG dump of every view, backward, Adam, log write).  Startup
(imports, mesh load, VGG build, CUDA context) is removed by differencing a short and a long run.  Variants: image dumps
encoded synchronously (ST3D_SYNC_IMAGE_WRITES=1, what the reference does) and by worker threads (default)."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-to-3d-style-transfer_b200")
sys.path[:0] = [ROOT, PKG, os.path.join(PKG, "compat")]
import numpy as np
import torch
from PIL import Image
from pytorch3d.io import save_obj

d = np.load(os.path.join(ROOT, "tests", "golden", "cow_mesh.npz"))
work = tempfile.mkdtemp(prefix="st3d_script_")
save_obj(os.path.join(work, "cow.obj"), torch.from_numpy(d["verts"]), torch.from_numpy(d["faces"]).long(),
         torch.from_numpy(d["verts_uvs"]), torch.from_numpy(d["faces_uvs"]).long(), torch.from_numpy(d["texture"]).float() / 255.0)
Image.fromarray(np.load(os.path.join(ROOT, "tests", "golden", "styles.npz"))["style_1"]).save(os.path.join(work, "Style_1.png"))
script = os.path.join(ROOT, "tests", "golden", "reference_scripts", "second_approach.py")


def run(epochs, sync_writes, first=10):
    """-> (ms per epoch between the log lines of epoch `first` and the last epoch, seconds from the last log line to exit)."""
    out = tempfile.mkdtemp(prefix="out_", dir=work)
    env = dict(os.environ, ST3D_SEED="7", ST3D_VGG_RANDOM_INIT="1", PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""),
               ST3D_SYNC_IMAGE_WRITES="1" if sync_writes else "0")
    cmd = [sys.executable, "-m", "st3d.run", script, "--n_views", "8", "--batch_size", "8", "--size", "512", "--epochs", str(epochs),
           "--obj_path", os.path.join(work, "cow.obj"), "--style_path", os.path.join(work, "Style_1.png"), "--output_path", out]
    p = subprocess.Popen(cmd, env=env, cwd=out, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    log, stamps = os.path.join(out, "log.txt"), {}
    while p.poll() is None:                         # the script appends one line per epoch (second_approach.py:193-194)
        try:
            with open(log) as fh:
                n = sum(1 for _ in fh)
        except OSError:
            n = 0
        stamps.setdefault(n, time.perf_counter())
        time.sleep(0.002)
    t_exit = time.perf_counter()
    assert p.returncode == 0, p.stderr.read()[-3000:]
    assert all(os.path.exists(os.path.join(out, "current_images", f"view_{i}.png")) for i in range(8))
    last = max(k for k in stamps if k <= epochs)
    return (stamps[last] - stamps[first]) / (last - first) * 1e3, t_exit - stamps[last]


res = {"workload": "second_approach.py unchanged, cow, 8 views x 512^2 in one batch, per epoch: everything the script does "
                   "(style image load, two render calls, loss, PNG dump of the 8 views, backward, Adam, log write); "
                   "epochs 10..60 by the time stamps of the script's own log lines"}
for label, sync in (("png_synchronous", True), ("png_worker_threads", False)):
    ms, lag = run(60, sync)
    res[label] = {"ms_per_epoch": round(ms, 2), "last_log_line_to_exit_s": round(lag, 2)}
res["speedup"] = round(res["png_synchronous"]["ms_per_epoch"] / res["png_worker_threads"]["ms_per_epoch"], 2)


def run_first(steps=400, first=100):
    """first_approach.py unchanged at BASELINE configs[0] (1 view x 256^2): ms per texture-fit step (first_approach.py:191-217:
    build_mesh, render_meshes, compute_first_approach_loss, backward, Adam, loss.item(), one log line per step)."""
    out = tempfile.mkdtemp(prefix="out1_", dir=work)
    env = dict(os.environ, ST3D_SEED="7", ST3D_VGG_RANDOM_INIT="1", PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "st3d.run", os.path.join(os.path.dirname(script), "first_approach.py"), "--n_views", "1",
           "--batch_size", "1", "--size", "256", "--n_style_transfer_steps", "50", "--n_mse_steps", str(steps),
           "--obj_path", os.path.join(work, "cow.obj"), "--style_path", os.path.join(work, "Style_1.png"), "--output_path", out]
    p = subprocess.Popen(cmd, env=env, cwd=out, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    log, stamps = os.path.join(out, "log.txt"), {}
    while p.poll() is None:
        try:
            n = os.path.getsize(log)
        except OSError:
            n = 0
        if n not in stamps:
            try:
                with open(log) as fh:
                    stamps.setdefault(sum(1 for line in fh if line.startswith("Batch 0, Step")), time.perf_counter())
            except OSError:
                pass
            stamps[n] = stamps.get(n, 0.0)
        time.sleep(0.001)
    assert p.returncode == 0, p.stderr.read()[-3000:]
    lines = {k: v for k, v in stamps.items() if k <= steps and v > 0.0}
    last = max(lines)
    lo = min(k for k in lines if k >= first)
    return (lines[last] - lines[lo]) / (last - lo) * 1e3


res["first_approach_fit_step"] = {"workload": "first_approach.py unchanged, cow, 1 view x 256^2 (BASELINE configs[0]): one texture-fit step "
                                              "incl. build_mesh, loss.item() and the log write",
                                  "ms_per_step": round(run_first(), 3)}
print(json.dumps(res))

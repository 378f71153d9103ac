"""Times gram_backward at C = 512 (C2 shapes) and checks it against the fp32 FFMA kernels (GPU box only).
ST3D_GRAM_BWD512_SINGLE=1 selects the one-CTA kernel instead of the CTA-pair (cta_group::2) kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import torch
from st3d import ops
torch.manual_seed(0)
for nhwc in (True, False):
    for side in (64, 32, 20):
        f = torch.relu(torch.randn(8, 512, side, side, device="cuda"))
        if nhwc:
            f = f.contiguous(memory_format=torch.channels_last)
        dg = torch.randn(8, 512, 512, device="cuda") * 1e-3
        ref = ops.gram_backward(f, dg, 1.0, precision="fp32")
        got = ops.gram_backward(f, dg, 1.0, precision="tf32")
        torch.cuda.synchronize()
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        for _ in range(3): ops.gram_backward(f, dg, 1.0, precision="tf32")
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(20): ops.gram_backward(f, dg, 1.0, precision="tf32")
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2 * 8 * 512 * 512 * side * side
        print(f"nhwc={nhwc} HW={side*side:5d} rel_err={err:.2e} {us:7.1f} us  {fl/us/1e6:7.1f} TF/s (incl. symmetrize)", flush=True)

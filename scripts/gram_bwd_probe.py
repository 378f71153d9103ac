"""Times gram_backward on the C2 layer shapes (plain, and with the fused accumulate + ReLU-mask epilogue) and checks
both against the fp32 FFMA kernels + torch elementwise ops (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import torch
from st3d import ops
torch.manual_seed(0)

def timeit(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for C, side in ((64, 512), (128, 256), (256, 128), (512, 64), (512, 32)):
    f = torch.relu(torch.randn(8, C, side, side, device="cuda")).contiguous(memory_format=torch.channels_last)
    dg = torch.randn(8, C, C, device="cuda") * 1e-3
    chain = torch.randn_like(f)
    ref = ops.gram_backward(f, dg, 1.0, precision="fp32")
    got = ops.gram_backward(f, dg, 1.0, precision="tf32")
    want_fused = torch.ops.aten.threshold_backward(chain + got, f, 0.0)
    got_fused = ops.gram_backward(f, dg, 1.0, out=chain.clone(memory_format=torch.preserve_format), accumulate=True,
                                  precision="tf32", relu_mask=True)
    torch.cuda.synchronize()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    buf = chain.clone(memory_format=torch.preserve_format)
    t_plain = timeit(lambda: ops.gram_backward(f, dg, 1.0, precision="tf32"))
    t_fused = timeit(lambda: ops.gram_backward(f, dg, 1.0, out=buf, accumulate=True, precision="tf32", relu_mask=True))
    nb = 8 * C * side * side * 4
    print(f"C={C:3d} HW={side*side:6d} rel_err={err:.1e} fused_exact={torch.equal(got_fused, want_fused)} "
          f"plain {t_plain:7.1f} us ({2*nb/t_plain/1e3:6.0f} GB/s)  fused {t_fused:7.1f} us ({3*nb/t_fused/1e3:6.0f} GB/s of 3 passes)",
          flush=True)

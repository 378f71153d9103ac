"""Runs the five Gram layer shapes of C2 forward+backward a few times (GPU box only; used under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200")]
import torch
from st3d import ops
N, S = 8, 512
for rep in range(2):
    for C, side in ((64, S), (128, S // 2), (256, S // 4), (512, S // 8), (512, S // 16)):
        f = torch.relu(torch.randn(N, C, side, side, device="cuda"))
        if os.environ.get("NHWC", "1") == "1":
            f = f.contiguous(memory_format=torch.channels_last)
        tgt = torch.randn(1, C, C, device="cuda"); loss = torch.zeros(1, device="cuda")
        dg, _ = ops.gram_mse_forward(f, tgt, 1e-9, loss)
        ops.gram_backward(f, dg, 1.0)
torch.cuda.synchronize(); print("ok")

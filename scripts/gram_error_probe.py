"""Prints the error of the tcgen05 Gram path against fp64 for a few layer shapes (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2d-to-3d-style-transfer_b200"))
import torch
from st3d import ops

for B, C, S in ((2, 64, 64), (2, 128, 32), (1, 256, 32), (1, 512, 16)):
    g = torch.Generator().manual_seed(C)
    f = (torch.relu(torch.randn(B, C, S, S, generator=g)) * 1.7).cuda()
    t = (torch.relu(torch.randn(1, C, S, S, generator=g)) * 1.7).cuda()
    G64 = f.double().reshape(B, C, -1) @ f.double().reshape(B, C, -1).transpose(1, 2)
    T64 = t.double().reshape(1, C, -1) @ t.double().reshape(1, C, -1).transpose(1, 2)
    for prec in ("tf32", "fp32"):
        G = ops.gram_forward(f, precision=prec).double()
        T = ops.gram_forward(t, precision=prec).double()
        e_g = ((G - G64).abs().max() / G64.abs().max()).item()
        bias = ((G - G64).sum() / G64.sum()).item()
        d, d64 = G - T, G64 - T64
        e_d = ((d - d64).abs().max() / d64.abs().max()).item()
        e_l = abs(((d ** 2).sum() - (d64 ** 2).sum()) / (d64 ** 2).sum()).item()
        dF = ops.gram_backward(f, (G - T64).float(), 1.0, precision=prec).double().reshape(B, C, -1)
        dF64 = 2 * d64 @ f.double().reshape(B, C, -1)
        e_b = ((dF - dF64).abs().max() / dF64.abs().max()).item()
        print(f"C={C:3d} HW={S*S:5d} {prec}: gram_err={e_g:.2e} bias={bias:+.2e} (G-T)_err={e_d:.2e} loss_err={e_l:.2e} dF_err={e_b:.2e}")

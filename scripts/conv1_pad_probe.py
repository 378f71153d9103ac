"""cuDNN time of VGG conv1_1 (3 -> 64, 3x3) forward (fused bias + ReLU) and data gradient at 8 x 512^2, channels_last, with the
input as it is (C = 3) and zero-padded to C = 4 / 8 (zero weights for the added channels: same sums).  GPU box only."""
import json
import torch
import torch.nn.functional as F

dev = "cuda"
torch.manual_seed(0)
w3 = torch.randn(64, 3, 3, 3, device=dev) * 0.1
b = torch.randn(64, device=dev) * 0.1
x3 = torch.rand(8, 3, 512, 512, device=dev)
res = {}


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 1)


ref = None
for C in (3, 4, 8):
    w = torch.zeros(64, C, 3, 3, device=dev)
    w[:, :3] = w3
    x = torch.zeros(8, C, 512, 512, device=dev)
    x[:, :3] = x3
    w = w.contiguous(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
    y = torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
    torch.manual_seed(1)
    g = torch.randn_like(y)
    fwd = timed(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1))
    bwd = timed(lambda: torch.ops.aten.convolution_backward(g, x, w, [64], [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                            [True, False, False]))
    gx = torch.ops.aten.convolution_backward(g, x, w, [64], [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
    if ref is None:
        ref = (y, gx[:, :3])
    res[f"C{C}"] = dict(fwd_us=fwd, dgrad_us=bwd, y_maxdiff=float((y - ref[0]).abs().max()),
                        gx_reldiff=float((gx[:, :3] - ref[1]).abs().max() / ref[1].abs().max()))
# the whole detour for a 3-channel channels_last input: pad -> conv ; dgrad -> first three channels as a dense tensor
for C in (4, 8):
    w = torch.zeros(64, C, 3, 3, device=dev)
    w[:, :3] = w3
    w = w.contiguous(memory_format=torch.channels_last)
    x = x3.contiguous(memory_format=torch.channels_last)
    g = torch.randn(8, 64, 512, 512, device=dev).contiguous(memory_format=torch.channels_last)

    def fwd():
        xp = F.pad(x, (0, 0, 0, 0, 0, C - 3))
        return torch.cudnn_convolution_relu(xp, w, b, (1, 1), (1, 1), (1, 1), 1), xp

    y, xp = fwd()
    assert xp.is_contiguous(memory_format=torch.channels_last)

    def bwd():
        gx = torch.ops.aten.convolution_backward(g, xp, w, [64], [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        return gx[:, :3].contiguous(memory_format=torch.channels_last)

    res[f"pad_to_C{C}_incl_copies"] = dict(fwd_us=timed(fwd), dgrad_us=timed(bwd))
print(json.dumps(res))

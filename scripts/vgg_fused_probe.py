"""conv+bias+relu per VGG-19 layer: eager (conv, add bias, relu_) vs torch.cudnn_convolution_relu, channels_last."""
import torch, torchvision
torch.manual_seed(0)
vgg = torchvision.models.vgg19(weights=None).features[:29].eval().cuda().to(memory_format=torch.channels_last)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
x = torch.rand(8, 3, 512, 512, device="cuda").contiguous(memory_format=torch.channels_last)
tot_a = tot_b = 0.0
with torch.no_grad():
    for name, m in vgg._modules.items():
        if isinstance(m, torch.nn.Conv2d):
            a = t(lambda: torch.relu_(m(x)))
            b = t(lambda: torch.cudnn_convolution_relu(x, m.weight, m.bias, m.stride, m.padding, m.dilation, m.groups))
            ya = torch.relu_(m(x)); yb = torch.cudnn_convolution_relu(x, m.weight, m.bias, m.stride, m.padding, m.dilation, m.groups)
            print(f"conv{name:>3} {tuple(x.shape)} -> {m.out_channels}: eager {a*1e3:7.1f} us  fused {b*1e3:7.1f} us  maxdiff {(ya-yb).abs().max().item():.2e} cl={yb.is_contiguous(memory_format=torch.channels_last)}")
            tot_a += a; tot_b += b
            x = ya
        elif isinstance(m, torch.nn.MaxPool2d):
            x = m(x)
print(f"total conv+bias+relu: eager {tot_a:.2f} ms  fused {tot_b:.2f} ms")

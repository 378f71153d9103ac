"""VGG-19 fwd+bwd time for 8x3x512x512 under different torch settings (GPU box only)."""
import torch, torchvision, time
torch.manual_seed(0)
vgg = torchvision.models.vgg19(weights=None).features[:29].eval().cuda()
for p in vgg.parameters(): p.requires_grad_(False)
def run(x, model, reps=5):
    for _ in range(2):
        y = model(x); y.sum().backward()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps):
        y = model(x); y.sum().backward()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
x = torch.rand(8, 3, 512, 512, device="cuda", requires_grad=True)
print("default nchw fwd+bwd ms:", run(x, vgg))
with torch.no_grad():
    for _ in range(2): vgg(x)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(5): vgg(x)
    e1.record(); torch.cuda.synchronize(); print("default nchw fwd only ms:", e0.elapsed_time(e1) / 5)
torch.backends.cudnn.benchmark = True
print("cudnn.benchmark nchw ms:", run(x, vgg))
vcl = vgg.to(memory_format=torch.channels_last)
xcl = x.detach().to(memory_format=torch.channels_last).requires_grad_(True)
print("channels_last ms:", run(xcl, vcl))
y = vcl(xcl); print("out strides", y.shape, y.stride(), y.is_contiguous(memory_format=torch.channels_last))
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
print("channels_last, no tf32 ms:", run(xcl, vcl))

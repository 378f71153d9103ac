"""2D NST step (style_transfer.py:59-83 through compat/) eager vs replayed from a CUDA graph, CUDA events, several sizes
(GPU box only)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-to-3d-style-transfer_b200"), os.path.join(ROOT, "2d-to-3d-style-transfer_b200", "compat")]
import torch, torchvision
from st3d import losses
from st3d.optimize import CapturedIteration
from st3d.vgg import fuse_vgg_features
dev = torch.device("cuda:0")
torch.manual_seed(0)
vgg = torchvision.models.vgg19(weights=None).features.eval().to(dev)
for p in vgg.parameters():
    p.requires_grad_(False)
model = fuse_vgg_features(vgg, channels_last=True)
res = {}
for B, S in ((1, 256), (4, 512), (8, 512)):
    g = torch.Generator().manual_seed(1)
    content = torch.rand(B, 3, S, S, generator=g).to(dev)
    style = torch.rand(B, 3, S, S, generator=g).to(dev)
    with torch.no_grad():
        cf = losses.get_features(content, model, {"21": "conv4_2"})["conv4_2"]
    grams = losses.style_targets(style, model)
    for mode in ("eager", "graph"):
        images = content.clone().requires_grad_(True)
        opt = torch.optim.Adam([images], lr=0.01, capturable=True, fused=True)
        def it():
            loss = losses.perceptual_loss_of_images(images, model, cf, grams, 1e6, 1.0)
            opt.zero_grad()
            loss.backward()
            opt.step()
        if mode == "graph":
            cap = CapturedIteration(it, dev, warmup=3)
            step = cap.replay
        else:
            for _ in range(3): it()
            step = it
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); torch.cuda.synchronize()
        res[f"{B}x{S}_{mode}_ms"] = round(e0.elapsed_time(e1) / 50, 4)
print(json.dumps(res, indent=1))

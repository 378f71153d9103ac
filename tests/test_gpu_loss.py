"""GPU parity tests of the loss half (Gram / style loss / content + masked MSE) through st3d.ops.

Bar (BASELINE.json north_star): fp32 path within 1e-4 relative; the tensor-core Gram (tcgen05
kind::tf32, i.e. at least bf16-grade operands) within 2e-3 relative.
"""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo

pytestmark = pytest.mark.gpu
TOL_FP32, TOL_TC = 1e-4, 2e-3


def _ops():
    import st3d
    return st3d.ops


def _relerr(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp(min=1e-30)).item()


def _features(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, C, H, W, generator=g)) * 1.7  # post-ReLU like the VGG taps


def test_gram_matches_reference_golden(golden_dir):
    import sys
    sys.path.insert(0, golden_dir)
    import make_golden as mg
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    _, _, _, feat, _ = mg.loss_inputs()          # (2,8,5,7): odd shape -> fp32 FFMA path
    got = ops.gram_forward(feat.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), g["gram"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,H,W", [(2, 8, 5, 7), (1, 64, 16, 16), (3, 96, 9, 13), (2, 128, 24, 20)])
def test_gram_fp32(B, C, H, W):
    ops = _ops()
    f = _features(B, C, H, W, 1)
    want = lo.gram_matrix(f.double())
    got = ops.gram_forward(f.cuda(), precision="fp32")
    assert _relerr(got, want) <= TOL_FP32


@pytest.mark.parametrize("B,C,H,W", [(1, 64, 16, 16), (2, 64, 64, 64), (2, 128, 32, 32), (3, 128, 24, 20),
                                       (2, 256, 16, 16), (1, 256, 36, 28), (2, 512, 8, 8), (1, 512, 32, 32),
                                       (8, 64, 20, 12), (1, 64, 6, 6)])
def test_gram_tcgen05(B, C, H, W):
    ops = _ops()
    f = _features(B, C, H, W, 2)
    want = lo.gram_matrix(f.double())
    got = ops.gram_forward(f.cuda(), precision="tf32")
    torch.cuda.synchronize()
    err = _relerr(got, want)
    assert err <= TOL_TC, f"tcgen05 Gram error {err:.3e}"
    # determinism: fixed-order split-K reduction
    assert torch.equal(got, ops.gram_forward(f.cuda(), precision="tf32"))
    # against the exact-fp32 device path too
    assert _relerr(got, ops.gram_forward(f.cuda(), precision="fp32")) <= TOL_TC


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("tf32", TOL_TC)])
@pytest.mark.parametrize("B,Bt,C,H,W", [(2, 1, 64, 32, 32), (2, 2, 128, 16, 16), (1, 1, 256, 16, 16), (2, 1, 512, 8, 8)])
def test_gram_mse_and_backward(precision, tol, B, Bt, C, H, W):
    ops = _ops()
    f = _features(B, C, H, W, 3)
    style = _features(Bt, C, H, W, 4)
    target = lo.gram_matrix(style.double())
    f64 = f.double().requires_grad_(True)
    scale = 1e6 / (B * C * C) / (C * C * H * H)   # style_weight * mean / (C^2 H^2), losses.py:35-39
    want = 1e6 * lo.style_layer_loss(f64, target)
    want.backward()
    loss = torch.zeros(1, device="cuda")
    dgram, gram = ops.gram_mse_forward(f.cuda(), target.float().cuda(), scale, loss, want_gram=True, precision=precision)
    grad = ops.gram_backward(f.cuda(), dgram, 1.0, precision=precision)
    torch.cuda.synchronize()
    assert _relerr(gram, lo.gram_matrix(f.double())) <= tol
    assert abs(loss.item() - want.item()) <= tol * abs(want.item()), (loss.item(), want.item())
    assert grad.shape == f.shape
    assert _relerr(grad, f64.grad) <= tol, f"grad error {_relerr(grad, f64.grad):.3e}"


@pytest.mark.parametrize("C,H,W", [(64, 40, 40), (128, 12, 20), (256, 16, 8), (512, 8, 12)])
def test_gram_backward_general_dgram(C, H, W):
    """dG need not be symmetric: grad = (dG + dG^T) F; accumulate adds into an existing buffer."""
    ops = _ops()
    B = 2
    f = _features(B, C, H, W, 5)
    g = torch.Generator().manual_seed(6)
    dG = torch.randn(B, C, C, generator=g)
    f64 = f.double().requires_grad_(True)
    (lo.gram_matrix(f64) * dG.double()).sum().backward()
    for precision, tol in (("fp32", TOL_FP32), ("tf32", TOL_TC)):
        got = ops.gram_backward(f.cuda(), dG.cuda(), 0.5, precision=precision)
        assert _relerr(got, 0.5 * f64.grad) <= tol, precision
        base = torch.ones_like(f).cuda()
        got2 = ops.gram_backward(f.cuda(), dG.cuda(), 0.5, out=base.reshape(B, C, -1), accumulate=True, precision=precision)
        assert _relerr(got2, 0.5 * f64.grad + 1.0) <= tol, precision


def test_gram_unsupported_shape_fails_loudly():
    ops = _ops()
    with pytest.raises(Exception):
        ops.gram_forward(torch.rand(1, 8, 5, 7, device="cuda"), precision="tf32")
    with pytest.raises(Exception):
        ops.gram_forward(torch.rand(1, 64, 8, 8), precision="fp32")   # CPU tensor: no CPU path


def test_mse_and_masked_mse(golden_dir):
    import sys
    sys.path.insert(0, golden_dir)
    import make_golden as mg
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "loss_golden.npz"))
    cur, con, _, _, masks = mg.loss_inputs()
    loss = torch.zeros(1, device="cuda")
    grad = ops.mse_forward(cur.cuda(), con.cuda(), 1.0 / cur.numel(), loss, mask=masks.cuda())
    np.testing.assert_allclose(loss.item(), g["first"], rtol=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), g["first_grad"], rtol=1e-4, atol=1e-9)
    # unmasked, odd length (scalar path) and large (vector path)
    for n in (1001, 1 << 20):
        gen = torch.Generator().manual_seed(n)
        a, b = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
        loss.zero_()
        gr = ops.mse_forward(a.cuda(), b.cuda(), 1.0 / n, loss)
        want = ((a.double() - b.double()) ** 2).mean()
        assert abs(loss.item() - want.item()) <= 1e-5 * want.item()
        assert _relerr(gr, 2 * (a.double() - b.double()) / n) <= 1e-6


def test_gram_full_size_properties():
    """BASELINE configs[1] layer sizes (8 images at 512^2): properties instead of a CPU oracle run."""
    ops = _ops()
    for C, S in ((64, 512), (128, 256), (256, 128), (512, 64), (512, 32)):
        B = 8
        g = torch.Generator(device="cuda").manual_seed(C + S)
        f = torch.relu(torch.randn(B, C, S, S, device="cuda", generator=g))
        G = ops.gram_forward(f, precision="tf32")
        torch.cuda.synchronize()
        assert torch.allclose(G, G.transpose(1, 2), rtol=1e-5, atol=1e-3 * G.abs().max().item())
        # diagonal = squared row norms; trace check against an fp64 reduction
        tr = (f.double() ** 2).sum(dim=(1, 2, 3))
        assert ((torch.diagonal(G, dim1=1, dim2=2).double().sum(1) - tr).abs() / tr).max() <= TOL_TC
        # checksum: 1^T G 1 = || sum_c F_c ||^2
        cs = (f.double().sum(1) ** 2).sum(dim=(1, 2))
        assert ((G.double().sum(dim=(1, 2)) - cs).abs() / cs).max() <= TOL_TC
        # homogeneity: G(2F) = 4 G(F) exactly (power-of-two scaling is exact in tf32/fp32)
        assert torch.equal(ops.gram_forward(2 * f, precision="tf32"), 4 * G)
        # backward: <dF, X> = <dG + dG^T, F X^T> for a random probe X (adjoint identity), via fp32 path
        dG = torch.randn(B, C, C, device="cuda", generator=g)
        dF = ops.gram_backward(f, dG, 1.0, precision="tf32")
        dF32 = ops.gram_backward(f, dG, 1.0, precision="fp32")
        assert _relerr(dF, dF32) <= TOL_TC


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("tf32", TOL_TC)])
@pytest.mark.parametrize("B,C,H,W", [(2, 64, 32, 32), (3, 64, 20, 12), (2, 128, 16, 24), (1, 256, 16, 16), (2, 256, 12, 20),
                                       (2, 512, 8, 8), (1, 512, 16, 12), (8, 64, 24, 24)])
def test_gram_channels_last_features(precision, tol, B, C, H, W):
    """channels_last (NHWC) feature maps -- what a channels_last VGG hands over -- are read in place:
    MN-major tcgen05 operands forward, K-major backward, gradient returned channels_last."""
    ops = _ops()
    f = _features(B, C, H, W, 11)
    fcl = f.cuda().contiguous(memory_format=torch.channels_last)
    assert not fcl.is_contiguous()
    style = lo.gram_matrix(_features(1, C, H, W, 12).double())
    f64 = f.double().requires_grad_(True)
    scale = 1e6 / (B * C * C) / (C * C * H * H)
    want = 1e6 * lo.style_layer_loss(f64, style)
    want.backward()
    G = ops.gram_forward(fcl, precision=precision)
    assert _relerr(G, lo.gram_matrix(f.double())) <= tol
    assert torch.equal(G, ops.gram_forward(fcl, precision=precision))            # deterministic
    loss = torch.zeros(1, device="cuda")
    dgram, _ = ops.gram_mse_forward(fcl, style.float().cuda(), scale, loss, precision=precision)
    assert abs(loss.item() - want.item()) <= tol * abs(want.item())
    grad = ops.gram_backward(fcl, dgram, 1.0, precision=precision)
    assert grad.shape == f.shape and grad.is_contiguous(memory_format=torch.channels_last)
    assert _relerr(grad, f64.grad) <= tol, _relerr(grad, f64.grad)
    # same numbers as the NCHW path up to the arithmetic mode
    grad_nchw = ops.gram_backward(f.cuda(), dgram, 1.0, precision=precision)
    assert _relerr(grad, grad_nchw.double()) <= tol
    # accumulate into an existing channels_last buffer
    base = torch.ones_like(fcl)
    got = ops.gram_backward(fcl, dgram, 0.5, out=base, accumulate=True, precision=precision)
    assert _relerr(got, 0.5 * f64.grad + 1.0) <= tol


def test_channels_last_autograd_and_mse():
    import st3d.functional as Fn
    B, C, H, W = 2, 64, 16, 16
    f = _features(B, C, H, W, 13)
    fcl = f.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    tgt = lo.gram_matrix(_features(1, C, H, W, 14).double())
    loss = 1e6 * Fn.style_layer_loss(fcl, tgt.float().cuda())
    other = _features(B, C, H, W, 15).cuda().contiguous(memory_format=torch.channels_last)
    loss = loss + 3.0 * Fn.mse_loss(fcl, other)
    loss.backward()
    f64 = f.double().requires_grad_(True)
    want = 1e6 * lo.style_layer_loss(f64, tgt) + 3.0 * ((f64 - other.cpu().double()) ** 2).mean()
    want.backward()
    assert abs(loss.item() - want.item()) <= TOL_TC * abs(want.item())
    assert _relerr(fcl.grad, f64.grad) <= TOL_TC, _relerr(fcl.grad, f64.grad)


def test_blended_multi_style_targets():
    """configs[3]: Gs = sum_j w_j Gram(style_j); the fused loss against the blended target equals the oracle's."""
    import torchvision
    from st3d import losses
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    for p in vgg.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(9)
    styles = torch.rand(4, 3, 64, 64, generator=g)
    cur = torch.rand(2, 3, 64, 64, generator=g)
    wts = [0.25, 0.25, 0.25, 0.25]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        targets = losses.blended_style_targets(styles.cuda(), wts, vgg.cuda())
        feats = losses.get_features(cur.cuda(), vgg)
        got = sum(losses.Fn.style_layer_loss(feats[k], t) for k, t in targets.items())
        vgg_cpu = vgg.cpu()
        sf = lo.get_features(styles, vgg_cpu)
        cf = lo.get_features(cur, vgg_cpu)
        want = 0.0
        for k in lo.STYLE_LAYERS:
            tgt = (lo.gram_matrix(sf[k]) * torch.tensor(wts).reshape(-1, 1, 1)).sum(0, keepdim=True)
            want = want + lo.style_layer_loss(cf[k], tgt)
        assert abs(got.item() - want.item()) <= TOL_TC * abs(want.item()), (got.item(), want.item())
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_fused_vgg_forward_backward_equal_unfused():
    """cuDNN's fused conv+bias+ReLU entry point: same taps, same input gradient as the stock module walk."""
    import torchvision
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval().cuda()
    for p in vgg.parameters():
        p.requires_grad_(False)
    fused = fuse_vgg_features(vgg, channels_last=True)
    plain = vgg.to(memory_format=torch.channels_last)
    x = torch.rand(2, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    fa, fb = losses.get_features(xa, fused), lo.get_features(xb, plain)
    assert list(fa) == list(fb)
    la = sum((f ** 2).mean() for f in fa.values())
    lb = sum((f ** 2).mean() for f in fb.values())
    la.backward()
    lb.backward()
    for k in fb:
        assert torch.allclose(fa[k], fb[k], rtol=1e-4, atol=1e-6), k
    assert _relerr(xa.grad, xb.grad) <= 1e-3


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 32, 48), (1, 128, 16, 16), (3, 4, 2, 2), (2, 512, 8, 4)])
def test_maxpool2x2_kernels_bit_identical_to_torch(B, C, H, W):
    """libst3d's NHWC 2x2 max pooling (csrc/pool.cu; the VGG-19 MaxPool2d modules of style_transfer.py:21-26):
    forward, backward, and backward with the fused ReLU mask, against torch's pooling + threshold_backward.
    Ties (the zeros a ReLU leaves) and NaNs must route exactly as ATen routes them."""
    import torch.nn.functional as F
    ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    x = torch.relu(torch.randn(B, C, H, W, device="cuda", generator=gen))            # plenty of tied zeros
    x[0, 0, 0, 1] = float("nan")
    x = x.contiguous(memory_format=torch.channels_last)
    g = torch.randn(B, C, H // 2, W // 2, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    y = ops.maxpool2x2_forward(x)
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(torch.nan_to_num(y, nan=-7.0), torch.nan_to_num(yr.detach(), nan=-7.0))
    (gr,) = torch.autograd.grad(yr, xr, g)
    assert torch.equal(ops.maxpool2x2_backward(x, g, relu_mask=False), gr)
    want = torch.ops.aten.threshold_backward(gr, x, 0.0)                                # the ReLU in front of the pool
    got = ops.maxpool2x2_backward(x, g, relu_mask=True)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))


def test_maxpool2x2_rejects_unsupported_inputs():
    ops = _ops()
    with pytest.raises(ValueError):
        ops.maxpool2x2_forward(torch.rand(1, 8, 5, 6, device="cuda").contiguous(memory_format=torch.channels_last))
    with pytest.raises(ValueError):
        ops.maxpool2x2_forward(torch.rand(1, 8, 4, 4, device="cuda"))                  # NCHW storage
    from st3d.vgg import FusedMaxPool
    pool = FusedMaxPool(torch.nn.MaxPool2d(2, 2), after_relu=True)
    x = torch.rand(1, 8, 5, 6, device="cuda")
    assert torch.equal(pool(x), torch.nn.functional.max_pool2d(x, 2, 2))               # torch's pooling takes over


def test_fused_vgg_pools_give_bit_identical_gradients():
    """Fused pools (+ the ReLU mask they carry) against the same VGG with torch's pooling and ReLU backward."""
    import torchvision
    from st3d import losses
    from st3d.vgg import FusedMaxPool, fuse_vgg_features
    torch.manual_seed(1)
    vgg = torchvision.models.vgg19(weights=None).features.eval().cuda()
    for p in vgg.parameters():
        p.requires_grad_(False)
    a = fuse_vgg_features(vgg, channels_last=True, fuse_pool=True)
    b = fuse_vgg_features(vgg, channels_last=True, fuse_pool=False)
    assert sum(isinstance(m, FusedMaxPool) for m in a) == 5 and not any(isinstance(m, FusedMaxPool) for m in b)
    assert [n for n, m in a._modules.items() if getattr(m, "feeds_masking_pool", False)] == ["2", "7", "16", "25", "34"]
    x = torch.rand(2, 3, 64, 96, device="cuda").contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    fa, fb = losses.get_features(xa, a), losses.get_features(xb, b)
    for k in fb:
        assert torch.equal(fa[k], fb[k]), k
    sum((f ** 2).mean() for f in fa.values()).backward()
    sum((f ** 2).mean() for f in fb.values()).backward()
    assert torch.equal(xa.grad, xb.grad)
    # a pre-pool activation that is ALSO tapped keeps its own ReLU backward (second consumer): still identical
    taps = {"2": "conv1_2", "7": "conv2_2"}
    xa2, xb2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ga, gb = losses.get_features(xa2, a, taps), losses.get_features(xb2, b, taps)
    sum((f ** 3).mean() for f in ga.values()).backward()
    sum((f ** 3).mean() for f in gb.values()).backward()
    assert torch.equal(xa2.grad, xb2.grad)


@pytest.mark.parametrize("precision", [None, "fp32"])     # None: tcgen05 where the shape allows it (conv5_1 of 64^2 does not)
def test_style_taps_inside_the_conv_layers_equal_the_feature_walk(precision):
    """perceptual_loss_of_images evaluates every style tap inside its conv + ReLU layer (Gram backward + gradient
    accumulation + ReLU mask in one kernel epilogue, ST3D_GRAM_ACCUMULATE | ST3D_GRAM_RELU_MASK); loss and image
    gradient must equal get_features + perceptual_loss_from_features (losses.py:26-42): bit for bit with fp32 Gram
    products; with tf32 the taps read the symmetric dG directly (ST3D_GRAM_DGRAM_SYMMETRIC) while the feature walk rounds
    s (dG + dG^T) to tf32 -- two roundings of the same operand, 1e-4 apart."""
    import torchvision
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    torch.manual_seed(2)
    vgg = torchvision.models.vgg19(weights=None).features.eval().cuda()
    for p in vgg.parameters():
        p.requires_grad_(False)
    model = fuse_vgg_features(vgg, channels_last=True)
    x = torch.rand(2, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    style = torch.rand(1, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        content = losses.get_features(torch.rand_like(x), model, {"21": losses.CONTENT_LAYER})[losses.CONTENT_LAYER]
    grams = losses.style_targets(style, model, precision)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = losses.perceptual_loss_of_images(xa, model, content, grams, 1e6, 1.0, precision)
    lb = losses.perceptual_loss_from_features(losses.get_features(xb, model), content, grams, 1e6, 1.0, precision)
    # one accumulator fed by float atomics, the weights applied inside the kernels, against per-layer sums weighted
    # afterwards: equal up to the order and place of a handful of fp32 roundings
    assert _relerr(la, lb) <= 5e-6
    la.backward()
    lb.backward()
    if precision == "fp32":
        assert torch.equal(xa.grad, xb.grad)
    else:
        assert _relerr(xa.grad, xb.grad) <= 5e-4, _relerr(xa.grad, xb.grad)
    # an unfused model takes the feature-walk route and still agrees to rounding
    xc = x.clone().requires_grad_(True)
    lc = losses.perceptual_loss_of_images(xc, vgg.to(memory_format=torch.channels_last), content, grams, 1e6, 1.0, precision)
    lc.backward()
    assert _relerr(lc, lb) <= 1e-5 and _relerr(xc.grad, xb.grad) <= 1e-3


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("C,H,W", [(64, 24, 20), (256, 16, 8), (512, 16, 16)])
def test_gram_backward_accumulate_and_relu_mask_flags(layout, precision, C, H, W):
    ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(C + H)
    f = torch.relu(torch.randn(2, C, H, W, device="cuda", generator=gen))
    chain = torch.randn(2, C, H, W, device="cuda", generator=gen)
    if layout == "nhwc":
        f, chain = f.contiguous(memory_format=torch.channels_last), chain.contiguous(memory_format=torch.channels_last)
    dg = torch.randn(2, C, C, device="cuda", generator=gen) * 1e-2
    plain = ops.gram_backward(f, dg, 0.5, precision=precision)
    want = torch.ops.aten.threshold_backward(chain + plain, f, 0.0)
    got = ops.gram_backward(f, dg, 0.5, out=chain.clone(memory_format=torch.preserve_format), accumulate=True,
                            precision=precision, relu_mask=True)
    assert torch.equal(got, want)
    assert torch.equal(ops.gram_backward(f, dg, 0.5, precision=precision, relu_mask=True),
                       torch.ops.aten.threshold_backward(plain, f, 0.0))


def test_full_size_pool_and_fused_tail_bit_identical_to_torch():
    """BASELINE configs[1] sizes (8 views x 512^2): the pooling kernels and the fused accumulate + ReLU tail of the
    Gram backward at conv1_1 / conv2_1 size against torch's own elementwise kernels on the same GPU (selection /
    one addition per element: bit-identical)."""
    import torch.nn.functional as F
    ops = _ops()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for C, side in ((64, 512), (128, 256)):
        f = torch.relu(torch.randn(8, C, side, side, device="cuda", generator=gen)).contiguous(memory_format=torch.channels_last)
        chain = torch.randn(8, C, side, side, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
        dg = torch.randn(8, C, C, device="cuda", generator=gen) * 1e-3
        plain = ops.gram_backward(f, dg, 1.0)
        want = torch.ops.aten.threshold_backward(chain + plain, f, 0.0)
        got = ops.gram_backward(f, dg, 1.0, out=chain, accumulate=True, relu_mask=True)      # in place on `chain`
        assert got.data_ptr() == chain.data_ptr() and torch.equal(got, want)
        del plain, want, got, chain
        xr = f.clone().requires_grad_(True)
        yr = F.max_pool2d(xr, 2, 2)
        y = ops.maxpool2x2_forward(f)
        assert torch.equal(y, yr.detach())
        g = torch.randn_like(y)
        (gr,) = torch.autograd.grad(yr, xr, g)
        assert torch.equal(ops.maxpool2x2_backward(f, g, relu_mask=False), gr)
        assert torch.equal(ops.maxpool2x2_backward(f, g, relu_mask=True), torch.ops.aten.threshold_backward(gr, f, 0.0))
        # conservation: every upstream value lands on exactly one input element
        assert abs(gr.double().sum().item() - g.double().sum().item()) <= 1e-6 * g.double().abs().sum().item()
        del xr, yr, y, g, gr, f
        torch.cuda.empty_cache()


@pytest.mark.parametrize("C,H,W", [(64, 24, 20), (256, 16, 16), (512, 16, 8)])
def test_gram_backward_symmetric_dgram_skips_the_symmetrise_pass(C, H, W):
    """ST3D_GRAM_DGRAM_SYMMETRIC: for a symmetric dG (what gram_mse_forward returns for Gram-matrix targets) the GEMM reads
    dG directly and applies 2 s to its accumulator -- same gradient as the general dG + dG^T path, one launch fewer."""
    ops = _ops()
    B = 3
    f = _features(B, C, H, W, 31).cuda().contiguous(memory_format=torch.channels_last)
    target = lo.gram_matrix(_features(1, C, H, W, 32).double()).float().cuda()
    loss = torch.zeros(1, device="cuda")
    dgram, _ = ops.gram_mse_forward(f, target, 1e-3, loss, precision="tf32")
    assert (dgram - dgram.transpose(1, 2)).abs().max().item() <= 1e-6 * dgram.abs().max().item()
    scale_t = torch.tensor([0.37], device="cuda")
    chain = torch.randn_like(f)
    before = ops.launch_count()
    fast = ops.gram_backward(f, dgram, 1.5, out=chain.clone(memory_format=torch.preserve_format), accumulate=True,
                             precision="tf32", scale_tensor=scale_t, relu_mask=True, symmetric_dgram=True)
    n_fast = ops.launch_count() - before
    before = ops.launch_count()
    general = ops.gram_backward(f, dgram, 1.5, out=chain.clone(memory_format=torch.preserve_format), accumulate=True,
                                precision="tf32", scale_tensor=scale_t, relu_mask=True)
    n_general = ops.launch_count() - before
    assert n_fast == n_general - 1
    f64 = f.double().cpu()
    want = 1.5 * 0.37 * 2.0 * torch.einsum("bij,bjhw->bihw", dgram.double().cpu(), f64) + chain.double().cpu()
    want = torch.where(f64 > 0, want, torch.zeros_like(want))
    assert _relerr(fast, want) <= TOL_TC, _relerr(fast, want)
    assert _relerr(fast, general.double()) <= TOL_TC


def test_content_tap_inside_its_conv_layer_equals_the_separate_passes():
    """perceptual_loss_of_images evaluates the content tap (conv4_2, losses.py:31) inside its conv + ReLU layer: the MSE
    backward, its sum with the gradient from deeper layers and the ReLU mask are one kernel (st3d_mse_tap_backward).
    The op against torch, then the whole walk against get_features + perceptual_loss_from_features."""
    import torchvision
    ops = _ops()
    from st3d import losses
    from st3d.vgg import fuse_vgg_features
    gen = torch.Generator(device="cuda").manual_seed(3)
    y = torch.relu(torch.randn(2, 64, 12, 10, device="cuda", generator=gen)).contiguous(memory_format=torch.channels_last)
    c = torch.randn(2, 64, 12, 10, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    g_in = torch.randn(2, 64, 12, 10, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    s = torch.tensor([0.7], device="cuda")
    got = ops.mse_tap_backward(y, c, g_in, 0.25, scale_tensor=s)
    want = torch.ops.aten.threshold_backward(g_in + (2 * 0.25 * 0.7) * (y - c), y, 0.0)
    assert got.stride() == y.stride() and torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    got0 = ops.mse_tap_backward(y, c, None, 0.25)
    assert torch.allclose(got0, torch.ops.aten.threshold_backward(0.5 * (y - c), y, 0.0), rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        ops.mse_tap_backward(y, c.contiguous(), g_in, 0.25)
    # the walk
    torch.manual_seed(2)
    vgg = torchvision.models.vgg19(weights=None).features.eval().cuda()
    for p in vgg.parameters():
        p.requires_grad_(False)
    model = fuse_vgg_features(vgg, channels_last=True)
    x = torch.rand(2, 3, 64, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    style = torch.rand(1, 3, 64, 64, device="cuda")
    with torch.no_grad():
        content = losses.get_features(torch.rand_like(x), model, {"21": losses.CONTENT_LAYER})[losses.CONTENT_LAYER]
    grams = losses.style_targets(style, model, "fp32")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = losses.perceptual_loss_of_images(xa, model, content, grams, 1e3, 50.0, "fp32")       # content term dominant
    lb = losses.perceptual_loss_from_features(losses.get_features(xb, model), content, grams, 1e3, 50.0, "fp32")
    assert _relerr(la, lb) <= 5e-6          # (one accumulator, weights inside the kernels: a few fp32 roundings apart)
    la.backward()
    lb.backward()
    assert _relerr(xa.grad, xb.grad) <= 1e-5, _relerr(xa.grad, xb.grad)


def _mesh_npz(golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"{name}_mesh.npz"))
    return torch.from_numpy(d["verts"]).float(), torch.from_numpy(d["faces"]).long()


@pytest.mark.parametrize("name", ["cow", "bob", "teapot"])
def test_fused_mesh_regularisers_match_the_oracle(golden_dir, name):
    """losses.py:85-87 (pytorch3d.loss.mesh_edge_loss / mesh_laplacian_smoothing / mesh_normal_consistency): values and
    vertex gradients of the one-launch CUDA op against the float64 oracle on the three meshes of the configs, a few
    steps into an optimisation (perturbed vertices).  fp32 sums of ~10^4 terms: 2e-5 on values, 2e-4 of the largest
    gradient entry on gradients."""
    from st3d import mesh_losses as ml
    verts, faces = _mesh_npz(golden_dir, name)
    g = torch.Generator().manual_seed(5)
    verts = verts + 2e-3 * torch.randn(verts.shape, generator=g)
    w = torch.tensor([1.0, 0.7, 0.3], dtype=torch.float64)
    v64 = verts.double().requires_grad_(True)
    want = torch.stack([lo.mesh_edge_loss(v64, faces), lo.mesh_laplacian_smoothing(v64, faces),
                        lo.mesh_normal_consistency(v64, faces)])
    (want * w).sum().backward()
    vc = verts.cuda().requires_grad_(True)
    fc = faces.cuda()
    got = ml.regularizers(vc, fc)
    (got * w.float().cuda()).sum().backward()
    assert torch.allclose(got.detach().cpu().double(), want.detach(), rtol=2e-5, atol=1e-9), (got, want)
    err = (vc.grad.cpu().double() - v64.grad).abs().max().item()
    assert err <= 2e-4 * v64.grad.abs().max().item(), err
    # the three pytorch3d.loss entry points select one term each, and the target length reaches the kernel
    for k, fn in enumerate((ml.edge_loss, ml.laplacian_smoothing, ml.normal_consistency)):
        assert torch.allclose(fn(vc, fc).detach(), got[k].detach(), rtol=1e-6, atol=1e-12)
    t = 0.01
    vt = verts.cuda().requires_grad_(True)
    ml.edge_loss(vt, fc, t).backward()
    v64b = verts.double().requires_grad_(True)
    want_t = lo.mesh_edge_loss(v64b, faces, t)
    want_t.backward()
    assert abs(ml.edge_loss(vc, fc, t).item() - want_t.item()) <= 2e-5 * abs(want_t.item())
    assert (vt.grad.cpu().double() - v64b.grad).abs().max().item() <= 2e-4 * v64b.grad.abs().max().item()
    # the workspace is left zero: a second call gives the same bits
    assert torch.equal(ml.regularizers(vc, fc).detach(), got.detach())


def test_fused_mesh_regularisers_edge_cases():
    """Empty face list, isolated vertices, an edge shared by three faces (three pairs), a zero-length edge and a
    zero-area face (clamped cosine, zero sub-gradient of the norm): same values and gradients as the oracle."""
    from st3d import mesh_losses as ml
    dev = torch.device("cuda")
    v = torch.rand(5, 3, device=dev, requires_grad=True)
    out = ml.regularizers(v, torch.zeros((0, 3), dtype=torch.long, device=dev))
    assert torch.equal(out.detach().cpu(), torch.zeros(3))
    out.sum().backward()
    assert torch.equal(v.grad.cpu(), torch.zeros(5, 3))
    g = torch.Generator().manual_seed(2)
    faces = torch.tensor([[0, 1, 2], [0, 1, 3], [1, 0, 4], [0, 5, 1]])   # (0,1) shared by 4 faces; vertex 6 isolated
    topo = ml.topology(faces.cuda(), 7)
    assert topo.pairs.shape[0] == 6 and topo.edges.shape[0] == 9 and int(topo.adj_ptr[-1]) == 18
    assert int(topo.adj_ptr[7] - topo.adj_ptr[6]) == 0
    w = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    for degenerate in (False, True):
        verts = torch.rand(7, 3, generator=g)
        if degenerate:
            verts[5] = verts[0]                                # edge (0,5) has length 0; face (0,5,1) has area 0
        v64 = verts.double().requires_grad_(True)
        want = torch.stack([lo.mesh_edge_loss(v64, faces), lo.mesh_laplacian_smoothing(v64, faces),
                            lo.mesh_normal_consistency(v64, faces)])
        (want * w).sum().backward()
        vc = verts.cuda().requires_grad_(True)
        got = ml.regularizers(vc, faces.cuda())
        (got * w.float().cuda()).sum().backward()
        assert torch.allclose(got.detach().cpu().double(), want.detach(), rtol=1e-5, atol=1e-7), (got, want)
        assert torch.isfinite(vc.grad).all()
        if not degenerate:      # (a zero normal clamped to eps = 1e-8 makes the reference's own gradient ~1e8: not compared)
            assert torch.allclose(vc.grad.cpu().double(), v64.grad, rtol=1e-3, atol=1e-5), \
                (vc.grad.cpu().double() - v64.grad).abs().max()
    with pytest.raises(ValueError):
        ml.regularizers(verts, faces)                           # CPU tensors: no library path

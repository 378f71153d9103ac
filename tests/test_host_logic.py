"""Host-side logic that needs no GPU: camera math, projection constants, the VGG feature walk and the
view-sharded gradient all-reduce (gloo, world_size 2)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import loss_oracle as lo
from oracle import render_oracle as ro


def test_cameras_match_oracle():
    from st3d import cameras as cm
    for seed in (0, 3):
        R, T = cm.random_view_cameras(6, generator=torch.Generator().manual_seed(seed))
        R2, T2 = ro.random_cameras(6, generator=torch.Generator().manual_seed(seed))
        assert torch.equal(R, R2) and torch.equal(T, T2)
    for axis in "XYZ":
        assert torch.equal(cm.rotate_axis_angle_matrix(37.0, axis)[0, :3, :3], ro.rotate_axis_angle_R(37.0, axis))
    R, T = cm.look_at_view_transform(2.7, 90.0, 10.0)       # looking straight down: degenerate x axis branch
    R2, T2 = ro.look_at_view_transform(2.7, 90.0, 10.0)
    assert torch.allclose(R, R2) and torch.allclose(T, T2)
    assert torch.allclose(R[0] @ R[0].t(), torch.eye(3), atol=1e-6)


def test_fov_scales_bit_identical():
    from st3d.functional import fov_scales
    for fov in (60.0, 45.0, 90.0):
        assert fov_scales(fov) == ro.fov_scales(fov)


def test_get_features_walk():
    import torchvision
    from st3d import losses
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    x = torch.rand(1, 3, 32, 32)
    with torch.no_grad():
        a = losses.get_features(x, vgg)
        b = lo.get_features(x.clone(), vgg)
    assert list(a) == list(b) == ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv4_2", "conv5_1"]
    for k in a:
        assert torch.equal(a[k], b[k])
        assert (a[k] >= 0).all()                 # in-place ReLU: every tap is post-ReLU


def test_constants_walk_taps_post_relu_on_a_plain_vgg(monkeypatch):
    """content_and_style_constants on a torchvision VGG whose ReLUs are separate in-place modules: the content feature
    and the style Grams must be those of the POST-ReLU activations, as get_features (and the reference,
    style_transfer.py:21-26) give them.  The Gram product is stood in by the oracle's on the CPU."""
    import torchvision
    from st3d import functional as Fn
    from st3d import losses
    monkeypatch.setattr(Fn, "gram_matrix", lambda t, precision=None: lo.gram_matrix(t))
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    content, style = torch.rand(2, 3, 32, 32), torch.rand(1, 3, 32, 32)
    with torch.no_grad():
        want_c = lo.get_features(content.clone(), vgg)["conv4_2"]
        want_s = {k: lo.gram_matrix(v) for k, v in lo.get_features(style.clone(), vgg).items() if k != "conv4_2"}
    got_c, got_g = losses.content_and_style_constants(content, style, vgg)
    assert (got_c >= 0).all() and torch.allclose(got_c, want_c, atol=1e-6)
    assert set(got_g) == set(want_s)
    for k in want_s:
        assert torch.allclose(got_g[k], want_s[k], rtol=1e-5, atol=1e-6 * float(want_s[k].abs().max()))
    # blended targets (BASELINE configs[3]) and the different-size branch go through the same taps
    styles = torch.rand(2, 3, 32, 32)
    _, blended = losses.content_and_style_constants(content, styles, vgg, style_weights=[0.25, 0.75])
    with torch.no_grad():
        each = [{k: lo.gram_matrix(v) for k, v in lo.get_features(styles[j:j + 1].clone(), vgg).items()} for j in range(2)]
    for k in want_s:
        want = 0.25 * each[0][k] + 0.75 * each[1][k]
        assert torch.allclose(blended[k], want, rtol=1e-5, atol=1e-6 * float(want.abs().max()))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from st3d.optimize import allreduce_gradients
    tex = torch.zeros(4, 4, 3, requires_grad=True)
    verts = torch.zeros(5, 3, requires_grad=True)
    unused = torch.zeros(2, requires_grad=True)
    # per-rank "view shard" gradients: the sharded result must equal the sum over shards
    g = torch.Generator().manual_seed(100 + rank)
    tex.grad = torch.rand(4, 4, 3, generator=g)
    verts.grad = torch.rand(5, 3, generator=g)
    allreduce_gradients([tex, verts, unused])
    # StyleOptimizer's form: the gradients are views of ONE flat buffer, reduced in place, asynchronously
    flat = torch.zeros(64 + 15)
    a, b = torch.zeros(4, 4, 3, requires_grad=True), torch.zeros(5, 3, requires_grad=True)
    a.grad, b.grad = flat[:48].view(4, 4, 3), flat[64:79].view(5, 3)
    a.grad.copy_(torch.rand(4, 4, 3, generator=g))
    b.grad.copy_(torch.rand(5, 3, generator=g))
    works = allreduce_gradients([a, b], flat=flat, async_op=True)
    assert len(works) == 1
    for w in works:
        w.wait()
    assert a.grad.data_ptr() == flat.data_ptr()         # still views: reduced where they lie
    # by value (numpy), not as shared-memory tensors: those are rebuilt in the parent through a socket of THIS process,
    # which may be gone by the time the parent reads the queue
    q.put((rank, tex.grad.numpy().copy(), verts.grad.numpy().copy(), unused.grad is None, a.grad.numpy().copy(),
           b.grad.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run_gloo_world(world):
    """One attempt: spawn `world` ranks on a fresh rendezvous port; None when the PLUMBING failed (a port taken between the
    probe and the bind, a rank that did not come up) -- the caller retries; numerical assertions are never retried."""
    import queue
    import socket
    with socket.socket() as sock:           # a free rendezvous port (a fixed one can linger in TIME_WAIT between runs)
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = []
    try:
        for _ in range(world):
            out.append(q.get(timeout=90))
    except queue.Empty:
        out = None
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.terminate()
            out = None
        elif p.exitcode != 0:
            out = None
    return out


def test_view_sharded_gradient_allreduce_gloo():
    world = 2
    out = None
    for _ in range(3):
        out = _run_gloo_world(world)
        if out is not None:
            break
    assert out is not None, "the gloo ranks did not come up in three attempts"
    out = sorted(((r, torch.from_numpy(t), torch.from_numpy(v), u, torch.from_numpy(a), torch.from_numpy(b))
                  for r, t, v, u, a, b in out), key=lambda t: t[0])
    want_tex = sum(torch.rand(4, 4, 3, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))

    def third_and_fourth_draws(r):
        gen = torch.Generator().manual_seed(100 + r)
        torch.rand(4, 4, 3, generator=gen), torch.rand(5, 3, generator=gen)
        return torch.rand(4, 4, 3, generator=gen), torch.rand(5, 3, generator=gen)
    want_a = sum(third_and_fourth_draws(r)[0] for r in range(world))
    want_b = sum(third_and_fourth_draws(r)[1] for r in range(world))
    for rank, tex, verts, unused_none, a, b in out:
        assert torch.allclose(tex, want_tex) and unused_none
        assert torch.allclose(a, want_a) and torch.allclose(b, want_b)
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


def test_allreduce_is_noop_single_process():
    from st3d.optimize import allreduce_gradients
    p = torch.ones(3, requires_grad=True)
    p.grad = torch.full((3,), 2.0)
    allreduce_gradients([p])
    assert torch.equal(p.grad, torch.full((3,), 2.0))


def _cpu_stand_ins(monkeypatch):
    """The fused VGG modules are CUDA-only and raise on CPU tensors.  The wiring tests below walk the fused model
    on the CPU, so they stand in for the two device calls (cuDNN's fused conv+bias+ReLU, libst3d's pooling) here,
    in the test, with torch's CPU ops."""
    from st3d import vgg as V
    monkeypatch.setattr(V.FusedConvReLU, "forward", lambda self, x: torch.relu_(self.conv(x)))
    monkeypatch.setattr(V.FusedMaxPool, "forward", lambda self, x: self.pool(x))


def test_fused_vgg_raises_on_cpu_tensors():
    import torchvision
    from st3d.vgg import fuse_vgg_features
    fused = fuse_vgg_features(torchvision.models.vgg19(weights=None).features.eval(), channels_last=True)
    with pytest.raises(RuntimeError, match="CUDA only"):
        fused(torch.rand(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA only"):
        fused._modules["4"](torch.rand(1, 64, 16, 16))


def test_fused_vgg_keeps_names_and_values_on_cpu(monkeypatch):
    """fuse_vgg_features keeps the module names get_features taps, and the values of the plain walk."""
    import torchvision
    from st3d import losses
    from st3d.vgg import FusedConvReLU, fuse_vgg_features
    _cpu_stand_ins(monkeypatch)
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    fused = fuse_vgg_features(vgg, channels_last=False)
    assert list(fused._modules) == list(vgg._modules)
    assert isinstance(fused._modules["0"], FusedConvReLU) and isinstance(fused._modules["1"], torch.nn.Identity)
    assert fused._modules["0"].conv.weight is vgg._modules["0"].weight
    x = torch.rand(1, 3, 32, 32)
    with torch.no_grad():
        a, b = losses.get_features(x, fused), lo.get_features(x.clone(), vgg)
    for k in b:
        assert torch.equal(a[k], b[k])


def test_fused_vgg_pool_wiring_and_cpu_walk(monkeypatch):
    """Structure of the fused VGG with libst3d pools: every MaxPool2d(2, 2) becomes a FusedMaxPool that carries the
    ReLU mask of the conv layer in front of it, that layer is told so, and the flags that guard the skipped ReLU
    backward behave (conservative default, cleared only inside get_features for untapped layers)."""
    import torchvision
    from st3d import losses
    from st3d.vgg import FusedConvReLU, FusedMaxPool, fuse_vgg_features
    _cpu_stand_ins(monkeypatch)
    torch.manual_seed(0)
    vgg = torchvision.models.vgg19(weights=None).features.eval()
    fused = fuse_vgg_features(vgg, channels_last=True, fuse_pool=True)
    pools = [n for n, m in fused._modules.items() if isinstance(m, FusedMaxPool)]
    assert pools == ["4", "9", "18", "27", "36"] and all(fused._modules[n].after_relu for n in pools)
    feeding = [n for n, m in fused._modules.items() if isinstance(m, FusedConvReLU) and m.feeds_masking_pool]
    assert feeding == ["2", "7", "16", "25", "34"]
    assert all(m.tapped for m in fused if isinstance(m, FusedConvReLU))          # conservative default
    seen = {}
    for name in ("0", "2"):
        mod = fused._modules[name]
        orig = mod.forward
        mod.forward = (lambda x, _o=orig, _m=mod, _n=name: (seen.__setitem__(_n, _m.tapped), _o(x))[1])
    x = torch.rand(1, 3, 32, 32)
    with torch.no_grad():
        a = losses.get_features(x, fused)
        b = lo.get_features(x.clone(), vgg)
    assert seen == {"0": True, "2": False}      # conv1_1 is a tap, conv1_2 is not: only the latter may skip its ReLU backward
    assert all(m.tapped for m in fused if isinstance(m, FusedConvReLU))          # and the default is restored
    for k in b:
        assert torch.equal(a[k], b[k])
    no_pool = fuse_vgg_features(vgg, channels_last=True, fuse_pool=False)
    assert not any(isinstance(m, FusedMaxPool) for m in no_pool)
    assert not any(getattr(m, "feeds_masking_pool", False) for m in no_pool)
    # an odd-sized or NCHW input is not for the libst3d pooling kernels
    assert not FusedMaxPool.accepts(torch.nn.MaxPool2d(3, 2)) and FusedMaxPool.accepts(torch.nn.MaxPool2d(2, 2))


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny job: ONE JSON line on stdout
    with the contract's keys, real full iterations (ms_per_step x steps is what the run spent), no GPU needed."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--views", "2", "--size", "64", "--gpus", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["higher_is_better"] is True and out["unit"] == "it/s"
    assert out["steps"] == 2 and out["warmup"] == 1 and out["n_gpus"] == 1
    assert abs(out["value"] - 1e3 / out["ms_per_step"]) <= 1e-9 * out["value"]          # a step IS a full iteration
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["views_per_step"] == 2 and "FULL iterations" in cb["sample"]
    assert out["e2e"] == {"value": out["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert out["config"]["global_views"] == 2 and out["config"]["workload"].startswith("cow_mesh")


def test_mesh_topology_tables():
    """The tables `st3d_mesh_regularizers_*` walk (include/st3d.h), built by st3d.mesh_losses.topology on the host:
    unique edges, CSR neighbour lists, one row per pair of faces sharing an edge (k faces -> k (k-1) / 2 rows)."""
    import torch
    from st3d import mesh_losses as ml
    tet = torch.tensor([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]])
    t = ml.topology(tet, 4)
    assert t.edges.dtype == torch.int32 and t.edges.tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert t.adj_ptr.tolist() == [0, 3, 6, 9, 12]
    assert [sorted(t.adj_idx[3 * i:3 * i + 3].tolist()) for i in range(4)] == [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]]
    assert t.pairs.shape == (6, 4)
    for v0, v1, a, b in t.pairs.tolist():                       # the four vertices of a tetrahedron, edge first
        assert v0 < v1 and sorted([v0, v1, a, b]) == [0, 1, 2, 3]
    fan = torch.tensor([[0, 1, 2], [0, 1, 3], [1, 0, 4]])       # three faces around the edge (0, 1); vertex 5 unused
    t = ml.topology(fan, 6)
    assert t.pairs.shape == (3, 4) and all(r[:2] == [0, 1] for r in t.pairs.tolist())
    assert sorted(tuple(sorted(r[2:])) for r in t.pairs.tolist()) == [(2, 3), (2, 4), (3, 4)]
    assert t.adj_ptr.tolist() == [0, 4, 8, 10, 12, 14, 14] and t.adj_idx.shape[0] == 2 * t.edges.shape[0] == 14
    empty = ml.topology(torch.zeros((0, 3), dtype=torch.long), 3)
    assert empty.edges.shape == (0, 2) and empty.pairs.shape == (0, 4) and empty.adj_ptr.tolist() == [0, 0, 0, 0]
    # the cache follows the storage and the version of the face tensor, not the Python object
    faces = tet.clone()
    a = ml._cached(faces, 4)
    assert ml._cached(faces.detach(), 4) is a
    faces[0, 0] = 0
    faces.add_(0)
    assert ml._cached(faces, 4) is not a

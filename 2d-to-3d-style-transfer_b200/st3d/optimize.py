"""The style-optimisation iteration itself -- the body of the reference's hot loop
(second_approach.py:147-190 for the `texture` target), batched over views and running on libst3d
kernels, with view-sharded data parallelism over NCCL (one process per GPU).

One iteration = render the content mesh (no grad) -> render the current mesh -> VGG features of
content / style / current -> content MSE + Gram style loss (+ mesh regularisers) -> backward into the
texture / vertices -> (all-reduce of the flat gradient when sharded) -> Adam step.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import functional as Fn
from . import losses


def allreduce_gradients(params, group=None, flat: Optional[torch.Tensor] = None, async_op: bool = False):
    """SUM all-reduce of the gradients over the view-sharding ranks; NCCL over NVLink when the process group is NCCL.

    flat: the ONE fp32 buffer [texture S*S*3 | verts V*3] the gradients are views of (StyleOptimizer keeps one):
    reduced in place as a single collective, no gather / scatter passes around it.  Without it every gradient is
    reduced in place on its own.  async_op=True returns the work handle(s) to `.wait()` on, so the caller can run
    view-independent work (the mesh regularisers) while the collective is in flight."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    if flat is not None:
        bufs = [flat]
    else:
        bufs = [p.grad for p in params if p.grad is not None]
    works = [dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=async_op) for b in bufs]
    return [w for w in works if w is not None]


class CapturedIteration:
    """`fn()` recorded ONCE into a CUDA graph and replayed: the loop body costs its kernels, not the host's launch rate
    (a step of the 2D style-transfer loop or of the style optimisation is ~150 launches of 5-300 us).

    `fn` must be capture-safe: static input / output tensors, no host reads, optimiser state on the device
    (`capturable=True`).  `warmup` eager calls run first on a side stream (lazy initialisation, cuDNN algorithm
    selection, allocator warm-up) -- they are REAL iterations, the caller counts them.  libst3d ops are capture-safe:
    their deferred workspace-header checks are skipped under capture, the owner of the graph calls `check()`."""

    def __init__(self, fn, device, warmup: int = 3):
        from . import ops
        self.device = torch.device(device)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(0, warmup)):
                self.out = fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        ops.poll_overflow(block=True)
        self.graph = torch.cuda.CUDAGraph()
        with ops.collect_captured_headers() as collected, torch.cuda.graph(self.graph):
            self.out = fn()
        self.headers = collected.headers        # pinned copies of the raster workspace headers, refreshed by every replay
        self.replays = 0
        self._replayed = None

    def replay(self):
        """Runs the recorded iteration; returns fn's (static) outputs, overwritten by the next replay."""
        self.graph.replay()
        self.replays += 1
        if self.headers:
            self._replayed = torch.cuda.Event()
            self._replayed.record(torch.cuda.current_stream(self.device))
        return self.out

    def check(self) -> None:
        """Waits for the last replay and raises if one of its renders overflowed the work lists recorded in the graph
        (geometry that moves between replays: the sizes seen when the graph was recorded are not guaranteed)."""
        if self._replayed is not None:
            from . import ops
            self._replayed.synchronize()
            self._replayed = None
            ops.check_captured_headers(self.headers)


DEFAULT_WEIGHTS = {"main_loss_weight": 3.0, "mesh_verts_weight": 1.0, "mesh_edge_loss_weight": 1.0,
                   "mesh_laplacian_smoothing_weight": 1.0, "mesh_normal_consistency_weight": 1.0}


class StyleOptimizer:
    """"Approach 2" of the reference (second_approach.py): optimise the texture and / or the vertices of a mesh so
    that its renders match a style image under a VGG perceptual loss.

    target   'texture' | 'mesh' | 'both'  (utils.py:173-204, losses.py:101-126)
    texture  UV map (T,T,3) with verts_uvs / faces_uvs, or per-vertex colours via `verts_rgb` (V,3)
    style    one image (1,3,H,W), or J images with `style_weights` -> blended Gram targets (BASELINE configs[3])
    """

    def __init__(self, verts, faces, vgg, image_size, *, verts_uvs=None, faces_uvs=None, texture=None, verts_rgb=None,
                 target: str = "texture", lr: float = 0.01, style_weight: float = 1e6, content_weight: float = 1.0,
                 weights=None, style_weights=None, precision=None, cache_constants: bool = False, world_size: int = 1,
                 group=None, channels_last: bool = True, fuse_conv_relu: bool = True, content_background: str = "white",
                 current_background: str = "white"):
        dev = verts.device
        if dev.type != "cuda":
            raise RuntimeError("StyleOptimizer needs CUDA tensors: libst3d has no CPU path")
        for b in (content_background, current_background):
            if b not in ("white", "noise", "style"):
                raise ValueError(f"unknown background type {b!r} (utils.py:19-30: 'white', 'noise' or 'style')")
        # apply_background of second_approach.py:161,166, composited inside the render epilogue
        self.content_background, self.current_background = content_background, current_background
        if target not in ("texture", "mesh", "both"):
            raise ValueError(f"unknown optimisation target {target!r}")
        if (texture is None) == (verts_rgb is None):
            raise ValueError("give either texture (+ verts_uvs, faces_uvs) or verts_rgb")
        self.target = target
        self.verts0 = verts.detach().float().contiguous()                       # content geometry / regulariser anchor
        self.faces = faces.to(torch.int32).contiguous()
        self.uv_mode = texture is not None
        if self.uv_mode:
            self.face_uvs = verts_uvs.float()[faces_uvs.long()].contiguous()    # (F,3,2)
            colour0 = texture
        else:
            self.face_uvs = None
            colour0 = verts_rgb
        self.content_colour = colour0.detach().clone().float().contiguous()     # the mesh as loaded
        self.colour = colour0.detach().clone().float().contiguous()
        self.verts = self.verts0.clone()
        params = []
        if target in ("mesh", "both"):
            params.append(self.verts.requires_grad_(True))
        if target in ("texture", "both"):
            params.append(self.colour.requires_grad_(True))
        self.params = params
        # ONE flat fp32 gradient buffer; every parameter's .grad is a view of it for the life of the optimiser:
        # zeroed with one fill, all-reduced in place as one collective, read by Adam where it lies
        # (each parameter starts on a 256-byte boundary so vectorised kernels keep their fast path)
        starts, off = [], 0
        for p in params:
            starts.append(off)
            off += -(-p.numel() // 64) * 64
        self._flat_grad = torch.zeros(off, device=dev, dtype=torch.float32)
        for p, o in zip(params, starts):
            p.grad = self._flat_grad[o:o + p.numel()].view_as(p)
        self.weights = dict(DEFAULT_WEIGHTS, **(weights or {}))
        # cuDNN's tensor-core convolutions are NHWC kernels: with NCHW tensors torch wraps every conv in
        # nchw<->nhwc transposes (~20 % of a step).  channels_last keeps activations NHWC end to end.
        self.channels_last = channels_last
        if fuse_conv_relu:          # cuDNN's fused conv + bias + ReLU entry point (same numbers, fewer kernels)
            from .vgg import fuse_vgg_features
            self.vgg = fuse_vgg_features(vgg, channels_last)
        else:
            self.vgg = vgg.to(memory_format=torch.channels_last) if channels_last else vgg
        self.image_size = image_size
        self.style_weight, self.content_weight = style_weight, content_weight
        # on the device from the start: a captured iteration cannot upload them (no pageable H2D copy under capture)
        self.style_weights = (None if style_weights is None else
                              torch.as_tensor(style_weights, dtype=torch.float32, device=verts.device))
        self.precision = precision
        self.cache_constants = cache_constants
        self.world_size, self.group = world_size, group
        # stock torch Adam (utils.py:185-195), its single-kernel CUDA implementation: the default multi-tensor one
        # spends 7 launches / 0.19 ms per step on a 512^2 texture
        self.optimizer = torch.optim.Adam(params, lr=lr, fused=True)
        self._cache = {}
        self._copy_stream = None
        self.images_ready: Optional[torch.cuda.Event] = None
        self._mesh_tables = None
        self.last_images: Optional[torch.Tensor] = None
        self._eager_steps = 0
        self._captured: Optional[CapturedIteration] = None
        self._grad_seed = torch.full((), 1.0 / self.world_size, device=self.verts0.device, dtype=torch.float32)

    # kept for the first API of this module
    @property
    def texture(self):
        return self.colour

    def _nn_input(self, images):
        return images.contiguous(memory_format=torch.channels_last) if self.channels_last else images

    def _background(self, kind, n_views, style_img):
        """utils.py:19-30: None for 'white' (the shader's own background), fresh uniform noise per call for 'noise'
        (torch's generator, as the reference draws it), the style image for 'style'."""
        if kind == "white":
            return None
        S = self.image_size
        H, W = (S, S) if isinstance(S, int) else tuple(S)
        if kind == "noise":
            return torch.rand((n_views, 3, H, W), device=self.verts0.device)
        if style_img.shape[0] != 1 or tuple(style_img.shape[-2:]) != (H, W):
            raise ValueError("background 'style' needs ONE style image of the render size")
        return style_img

    def _render(self, verts, colour, R, T, background_image=None):
        kw = dict(texture=colour, face_uvs=self.face_uvs) if self.uv_mode else dict(verts_rgb=colour)
        # channels_last VGG: the renderer writes (and its backward reads) the images in that storage, no layout copies
        images, masks, _ = Fn.render_views(verts, self.faces, R, T, self.image_size, background_image=background_image,
                                           channels_last=self.channels_last, **kw)
        return images, masks

    def _regularisers(self):
        """losses.py:85-87 / 113-115: the vertex MSE plus the three mesh regularisers -- the latter from ONE libst3d
        launch (and one for their backward) over topology tables built once (st3d/mesh_losses.py)."""
        from . import mesh_losses as ml
        if self._mesh_tables is None:
            w = self.weights
            self._mesh_tables = (ml.topology(self.faces, self.verts.shape[0]),
                           torch.tensor([w["mesh_edge_loss_weight"], w["mesh_laplacian_smoothing_weight"],
                                         w["mesh_normal_consistency_weight"]], device=self.verts.device))
        topo, w3 = self._mesh_tables
        return (self.weights["mesh_verts_weight"] * Fn.mse_loss(self.verts, self.verts0)
                + torch.dot(ml.regularizers(self.verts, self.faces, 0.0, 7, topo=topo), w3))

    # ---- cache of the loop constants (content feature, style Grams): keyed on CONTENT, never on addresses ----------
    # The caching allocator hands the address of a freed per-batch temporary (R[idx].to(dev), the reference's own
    # batching loop) to the next one, and an in-place edit keeps the address: a data_ptr key would return the
    # constants of other cameras or another style image without an error.
    def _remember_constants(self, slot, R, T, style_img, content_feat, grams) -> None:
        inputs = (R, T, style_img)
        self._cache[slot] = dict(inputs=inputs, versions=tuple(t._version for t in inputs),
                                 copies=tuple(t.detach().clone() for t in inputs), content_feat=content_feat,
                                 grams=grams)

    def _constants_cached_for(self, slot, R, T, style_img) -> bool:
        """slot: which micro-batch of the step (its first view); every slot keeps its own constants."""
        c = self._cache.get(slot)
        if not c:
            return False
        inputs = (R, T, style_img)
        # fast path, no device work: the very tensor objects of the cached call (the cache keeps them alive, so their
        # storage cannot have been recycled) and no in-place write since (tensor._version)
        if all(a is b for a, b in zip(inputs, c["inputs"])) and tuple(t._version for t in inputs) == c["versions"]:
            return True
        # otherwise compare the values with the private copies (one small device reduction + host read per call)
        same = all(a.shape == b.shape and a.device == b.device and a.dtype == b.dtype and bool(torch.equal(a, b))
                   for a, b in zip(inputs, c["copies"]))
        if same:        # re-arm the fast path for these objects
            c["inputs"], c["versions"] = inputs, tuple(t._version for t in inputs)
        return same

    def _export_images(self, images: torch.Tensor, images_out: torch.Tensor) -> None:
        """Device -> pinned-host copy of this step's renders on a side stream, so that it overlaps the VGG passes
        that follow instead of trailing the step (the reference dumps every view every step,
        second_approach.py:183-185).  `self.images_ready` is the event to wait on before reading `images_out`."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=images.device)
        cur = torch.cuda.current_stream(images.device)
        self._copy_stream.wait_stream(cur)
        with torch.cuda.stream(self._copy_stream):
            images_out.copy_(images, non_blocking=True)
            self.images_ready = torch.cuda.Event()
            self.images_ready.record(self._copy_stream)
        if not torch.cuda.is_current_stream_capturing():    # (a captured copy stream joins back before the graph ends)
            images.record_stream(self._copy_stream)

    def _views_loss(self, R, T, style_img, images_out, cache_slot, loss_scale=1.0):
        """Perceptual loss (losses.py:12-44) of the views (R, T), times `loss_scale`: content render + constants, current
        render, VGG walk."""
        cacheable = self.cache_constants and self.content_background != "noise"     # fresh noise: nothing is constant
        if cacheable and self._constants_cached_for(cache_slot, R, T, style_img):
            c = self._cache[cache_slot]
            content_feat, grams = c["content_feat"], c["grams"]
        else:
            with torch.no_grad():
                content_imgs, _ = self._render(self.verts0, self.content_colour, R, T,          # second_approach.py:160-161
                                               self._background(self.content_background, R.shape[0], style_img))
            # losses.py:18-25: content features + style Grams, one VGG walk for both constant branches
            content_feat, grams = losses.content_and_style_constants(content_imgs, style_img, self.vgg, self.precision,
                                                                     self.style_weights)
            if cacheable:
                self._remember_constants(cache_slot, R, T, style_img, content_feat, grams)
        current_imgs, _ = self._render(self.verts, self.colour, R, T,                           # :165-166
                                       self._background(self.current_background, R.shape[0], style_img))
        if images_out is not None:
            self._export_images(current_imgs.detach(), images_out)
        self.last_images = current_imgs.detach()
        return losses.perceptual_loss_of_images(self._nn_input(current_imgs), self.vgg, content_feat, grams,
                                                self.style_weight, self.content_weight, self.precision, loss_scale)

    def _accumulate_gradients(self, R, T, style_img, images_out=None, micro_batch=None) -> torch.Tensor:
        """Everything of an iteration that depends on this rank's views: zero the flat gradient, then per micro-batch
        render -> VGG walk -> loss -> backward (gradients accumulate).  Returns the perceptual part of the loss."""
        B = R.shape[0]
        mb = B if not micro_batch else max(1, min(int(micro_batch), B))
        self._flat_grad.zero_()         # the .grad views stay in place (optimizer.zero_grad(set_to_none=False) as one fill)
        main_w = self.weights["main_loss_weight"] if self.target != "texture" else 1.0          # losses.py:108-124
        loss = None
        for s in range(0, B, mb):
            Rm, Tm = (R, T) if mb == B else (R[s:s + mb], T[s:s + mb])
            out_m = None if images_out is None else (images_out if mb == B else images_out[s:s + mb])
            # the weights of this chunk in the iteration's loss go down to the kernels that accumulate it, and the 1 / world
            # of the gradient average is the seed of the backward: no scalar kernels around the loss
            part = self._views_loss(Rm, Tm, style_img, out_m, s, loss_scale=main_w * (Rm.shape[0] / B))
            part.backward(gradient=self._grad_seed)                                             # :188
            loss = part.detach() if loss is None else loss + part.detach()
        if self._copy_stream is not None and torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream().wait_stream(self._copy_stream)      # a captured side stream must join back
        return loss

    def _reduce_and_update(self, loss: torch.Tensor) -> torch.Tensor:
        """The part of an iteration that involves every rank: ONE collective, issued asynchronously -- the regularisers
        are view-independent and identical on every rank, so their forward + backward run while the reduce is in
        flight and are added once, after it -- then the Adam step."""
        works = allreduce_gradients(self.params, self.group, flat=self._flat_grad, async_op=True)
        if self.target != "texture":
            reg = self._regularisers()
            (g_verts,) = torch.autograd.grad(reg, self.verts)
            for w in works:
                w.wait()
            self.verts.grad.add_(g_verts)
            loss = loss + reg.detach()
        else:
            for w in works:
                w.wait()
        self.optimizer.step()                                                                   # :189
        return loss

    def step(self, R: torch.Tensor, T: torch.Tensor, style_img: torch.Tensor,
             images_out: Optional[torch.Tensor] = None, micro_batch: Optional[int] = None) -> torch.Tensor:
        """One optimisation iteration over the views (R, T) held by this rank; returns the loss (device scalar).
        images_out: optional pinned host tensor (B,3,H,W) that receives this step's rendered views asynchronously
        (wait on `self.images_ready` before reading it).
        micro_batch: render / walk the VGG over this many views at a time and ACCUMULATE their gradients into the one
        Adam step (the loss is a mean over views, losses.py:31,38, so a chunk of b of the B views weighs b / B): the
        same iteration as one B-view batch with the activations of only `micro_batch` views alive."""
        self._eager_steps += 1
        return self._reduce_and_update(self._accumulate_gradients(R, T, style_img, images_out, micro_batch))

    def capture(self, R: torch.Tensor, T: torch.Tensor, style_img: torch.Tensor,
                images_out: Optional[torch.Tensor] = None, micro_batch: Optional[int] = None, warmup: int = 3) -> None:
        """Records this rank's share of the iteration (renders, VGG walks, losses, backward, the copy of the rendered
        views to `images_out`) as ONE CUDA graph over the STATIC tensors R, T, style_img: new cameras or a new style
        image are copied INTO them before `step_captured()`.  The warm-up runs `warmup` real iterations.  What stays
        outside the graph is what involves other ranks or the host: the NCCL all-reduce, the regularisers that overlap
        it (two launches, csrc/mesh_reg.cu), and the Adam step.  All three targets can be captured."""
        if self._eager_steps:
            # the leaves' gradient accumulators were bound to the stream the eager steps ran on (normally the legacy
            # default stream); autograd would make that stream wait on the capturing one, which CUDA refuses
            raise RuntimeError("capture() must be called before any eager step(): build a fresh optimiser for the captured loop")
        if self.cache_constants:
            # the constants would be computed in the warm-up and left OUT of the recorded iteration: new cameras copied
            # into the static tensors would then meet the old content feature without any error
            raise ValueError("capture() records the whole iteration, constants included: build the optimiser with "
                             "cache_constants=False")
        # `mesh` / `both`: a moving mesh changes the sizes of the rasterizer's work lists from step to step.  The graph
        # records a copy of every workspace header to pinned memory; step_captured() reads the copies of the previous
        # replay (work lists are sized at twice what the warm-up needed) and raises if one overflowed.

        def grads():
            return self._accumulate_gradients(R, T, style_img, images_out, micro_batch)

        def whole():
            return self._reduce_and_update(grads())

        dev = self.verts0.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                whole()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self._captured = CapturedIteration(grads, dev, warmup=0)

    def step_captured(self) -> torch.Tensor:
        """Replays the captured gradient computation, then all-reduce / regularisers / Adam; returns the loss."""
        if self._captured is None:
            raise RuntimeError("call capture(R, T, style_img) first")
        if self.target != "texture":
            self._captured.check()      # the previous replay (its regularisers / Adam tail is still queued behind it)
        loss = self._captured.replay()
        return self._reduce_and_update(loss)


class TextureStyleOptimizer(StyleOptimizer):
    """Texture-only optimisation of a UV-mapped mesh (the `texture` target, BASELINE configs[1])."""

    def __init__(self, verts, faces, verts_uvs, faces_uvs, texture, vgg, image_size, lr: float = 0.01,
                 style_weight: float = 1e6, content_weight: float = 1.0, precision=None, cache_constants: bool = False,
                 world_size: int = 1, group=None, channels_last: bool = True, fuse_conv_relu: bool = True):
        super().__init__(verts, faces, vgg, image_size, verts_uvs=verts_uvs, faces_uvs=faces_uvs, texture=texture,
                         target="texture", lr=lr, style_weight=style_weight, content_weight=content_weight,
                         precision=precision, cache_constants=cache_constants, world_size=world_size, group=group,
                         channels_last=channels_last, fuse_conv_relu=fuse_conv_relu)


class GraphedTextureFit:
    """The texture-fit loop of "approach 1" (first_approach.py:191-213, `texture` target: render -> masked MSE against
    the target renders -> backward -> Adam) with the whole iteration captured in ONE CUDA graph.

    At the reference's own size (1 view x 256^2, BASELINE configs[0]) an iteration is ~10 short kernels: launched
    one by one from Python the loop is bound by the host (~1 ms / iteration); replayed as a graph it costs the
    kernels' own time.  Cameras and target images are fixed for the life of the object (the reference re-uses the
    same views every epoch); `step()` replays the graph and returns the loss tensor of that iteration."""

    def __init__(self, verts, faces, verts_uvs, faces_uvs, texture, R, T, target_images, image_size, lr: float = 0.01,
                 warmup: int = 3):
        from . import ops
        dev = verts.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTextureFit needs CUDA tensors: libst3d has no CPU path")
        self.verts = verts.detach().float().contiguous()
        self.faces = faces.to(torch.int32).contiguous()
        self.face_uvs = verts_uvs.float()[faces_uvs.long()].contiguous()
        self.texture = texture.detach().clone().float().contiguous().requires_grad_(True)
        self.R, self.T = R.detach().float().contiguous().to(dev), T.detach().float().contiguous().to(dev)
        self.target = target_images.detach().float().contiguous().to(dev)
        self.image_size = image_size
        # capturable: the step counter lives on the device, so the update can be replayed
        self.optimizer = torch.optim.Adam([self.texture], lr=lr, capturable=True, fused=True)
        self._workspaces = []
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):               # warm-up outside the capture (lazy initialisation, autotuning)
            for _ in range(max(1, warmup)):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        ops.poll_overflow(block=True)
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._iteration()
        self.iterations = max(1, warmup)

    def _iteration(self):
        self.optimizer.zero_grad(set_to_none=True)
        images, masks, _ = Fn.render_views(self.verts, self.faces, self.R, self.T, self.image_size,
                                           texture=self.texture, face_uvs=self.face_uvs)
        loss = Fn.masked_mse_loss(images, self.target, masks)                       # losses.py:71-75
        loss.backward()
        self.optimizer.step()
        self.images = images.detach()
        return loss.detach()

    def step(self) -> torch.Tensor:
        """Replays one iteration; the returned tensor (and `self.images`) are overwritten by the next replay."""
        self.graph.replay()
        self.iterations += 1
        return self.loss

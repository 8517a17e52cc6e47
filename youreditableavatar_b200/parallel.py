"""Data-parallel multi-view batches: views are independent units, gradients are summed.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests).
Every rank holds all Gaussian parameters, renders its share of the batch's views, and the per-view backward
kernels ADD their gradients straight into one flat fp32 buffer (`GradBucket`) — the eight gradient tensors the
rasterizer returns are views into that buffer — so the only exchange step is a single all-reduce of the buffer
with no pack / unpack copies.  The reference has no multi-GPU path at all (one view per optimizer step,
Edit_core/tetgs_texture/refine.py:54); this is new.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch

GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations")


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """View v goes to rank v mod world_size (round-robin keeps per-rank work even for orbit batches)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_views, world_size))


def grad_shapes(P: int, M: int) -> Tuple[Tuple[int, ...], ...]:
    """Shapes of the 8-tuple `_C.rasterize_gaussians_backward` returns (rasterize_points.cu:151-159)."""
    return ((P, 3), (P, 3), (P, 1), (P, 3), (P, 6), (P, M, 3), (P, 3), (P, 4))


class GradBucket:
    """One flat fp32 buffer holding the gradient tensors back to back; `.views` are the eight gradient tensors (in
    the order `_C.rasterize_gaussians_backward` returns them) aliasing it.  Memory order: the small tensors first,
    dL_dsh (81 % of the bytes at SH degree 3) LAST, so that "everything but SH" is one contiguous slice and SH rows
    of a range of Gaussians are another — the two shapes the overlapped all-reduce sends.  Offsets are multiples of
    4 floats (16-byte vector stores in the kernels need the SH / rotation views 16-byte aligned)."""

    # what an optimizer over (means/delta, opacity, SH, scales, rotations) consumes; the other three exist in the
    # reference's return tuple only because its kernels produce them on the way
    TRAINING = ("dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations")

    def __init__(self, P: int, M: int, device="cuda", names: Optional[Sequence[str]] = None):
        self.P, self.M = P, M
        self.shapes = grad_shapes(P, M)
        self.names = tuple(GRAD_NAMES if names is None else names)
        offs, total = [None] * len(GRAD_NAMES), 0
        order = [i for i, n in enumerate(GRAD_NAMES) if n != "dL_dsh"] + [GRAD_NAMES.index("dL_dsh")]
        for i in order:
            if GRAD_NAMES[i] not in self.names:
                continue
            offs[i] = total
            n = 1
            for d in self.shapes[i]:
                n *= d
            total += (n + 3) // 4 * 4
        self.offsets = offs
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.views = tuple(None if o is None else self._view(o, shp) for o, shp in zip(offs, self.shapes))
        self.sh_offset = offs[GRAD_NAMES.index("dL_dsh")]   # None when SH gradients are not in the bucket
        self._pending = []

    def _view(self, off, shp):
        n = 1
        for d in shp:
            n *= d
        return self.flat[off:off + n].view(*shp)

    def named(self) -> Dict[str, torch.Tensor]:
        return dict(zip(GRAD_NAMES, self.views))

    @staticmethod
    def _distributed(group=None) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def all_reduce(self, group=None):
        """Sum over ranks in one call — the only collective of a step."""
        import torch.distributed as dist
        if self._distributed(group):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self

    # -- overlapped variant: ranges of SH rows leave while the backward still computes the next range ---------
    def all_reduce_sh_rows_async(self, first: int, count: int, group=None):
        """Starts the all-reduce of dL_dsh[first:first+count] (contiguous in `flat`).  Stream-ordered after the work
        already queued on the current stream, runs on the process group's own stream."""
        import torch.distributed as dist
        if self.sh_offset is None or not self._distributed(group):
            return
        row = self.M * 3
        sl = self.flat[self.sh_offset + first * row: self.sh_offset + (first + count) * row]
        self._pending.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=group, async_op=True))

    def all_reduce_rows_async(self, first: int, count: int, group=None):
        """Starts the all-reduce of rows [first, first+count) of EVERY gradient tensor in the bucket (one coalesced
        NCCL launch for the up-to-eight slices).  With this variant nothing but the last range is left for the end
        of the backward; `wait()` joins."""
        import torch.distributed as dist
        if not self._distributed(group):
            return
        slices = [t[first:first + count] for t in self.views if t is not None]
        if dist.get_backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
            with dist._coalescing_manager(group=group, device=self.flat.device, async_ops=True) as cm:
                for sl in slices:
                    dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=group)
            self._pending.append(cm)
        else:                                               # e.g. gloo in the CPU tests: one op per slice
            for sl in slices:
                self._pending.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=group, async_op=True))

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        return self

    def all_reduce_rest_and_wait(self, group=None):
        """All-reduces everything that is not SH (one contiguous slice) and joins the pending SH ranges."""
        import torch.distributed as dist
        if self._distributed(group):
            end = self.sh_offset if self.sh_offset is not None else self.flat.numel()
            if end > 0:
                self._pending.append(dist.all_reduce(self.flat[:end], op=dist.ReduceOp.SUM, group=group, async_op=True))
            for w in self._pending:
                w.wait()
        self._pending = []
        return self


class SymmGradBucket(GradBucket):
    """GradBucket whose flat buffer lives in symmetric memory (torch.distributed._symmetric_memory: allocation,
    rendezvous, multicast mapping and the cross-rank barrier are plumbing).  `all_reduce` is then ONE kernel of this
    library — `tgr_multimem_allreduce_f32`, a two-shot all-reduce with in-switch reduction over NVSwitch
    (csrc/collective.cu) — instead of an NCCL call.  Falls back to NCCL when the devices have no multicast support
    (`self.nvls` tells which path is live)."""

    def __init__(self, P: int, M: int, device="cuda", names: Optional[Sequence[str]] = None, group=None):
        import torch.distributed as dist
        super().__init__(P, M, device, names)
        self.group = group
        self.hdl, self.nvls = None, False
        if self._distributed(group):
            import torch.distributed._symmetric_memory as symm_mem
            flat = symm_mem.empty(self.flat.numel(), dtype=torch.float32, device=self.flat.device)
            flat.zero_()
            self.flat = flat
            self.views = tuple(None if o is None else self._view(o, shp) for o, shp in zip(self.offsets, self.shapes))
            g = group if group is not None else dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.flat, g)
            self.nvls = bool(getattr(self.hdl, "multicast_ptr", 0))

    def _mc(self, offset_floats: int = 0) -> int:
        h = self.hdl
        return h.multicast_ptr + (self.flat.data_ptr() - h.buffer_ptrs[h.rank]) + 4 * offset_floats

    def all_reduce(self, group=None):
        if not self.nvls:
            return super().all_reduce(group if group is not None else self.group)
        from . import _lib
        h = self.hdl
        stream = torch.cuda.current_stream(self.flat.device).cuda_stream
        h.barrier(channel=0)   # every rank's gradients are written (stream-ordered on each rank)
        _lib.check(_lib.lib().tgr_multimem_allreduce_f32(self._mc(), self.flat.numel(), h.rank, h.world_size, stream),
                   "tgr_multimem_allreduce_f32")
        h.barrier(channel=1)   # every rank's slice has been broadcast
        return self

    # -- overlapped variant: the SH rows of a range of Gaussians are summed in the switch, on a high-priority side stream
    #    with a capped grid, while the backward's per-Gaussian kernel computes the next range ---------------------------
    import os as _os
    MAX_CTAS_OVERLAPPED = int(_os.environ.get("TGR_NVLS_CTAS", "32"))

    def _side(self) -> torch.cuda.Stream:
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.flat.device, priority=-1)
            self._channel = 0
        return self._side_stream

    def _next_channel(self) -> int:
        self._channel = (self._channel + 1) % 8
        return self._channel

    def all_reduce_sh_rows_async(self, first: int, count: int, group=None):
        if not self.nvls:
            return super().all_reduce_sh_rows_async(first, count, group if group is not None else self.group)
        if self.sh_offset is None:
            return
        from . import _lib
        h, side = self.hdl, self._side()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))        # this range's rows are written (stream order)
        row = self.M * 3
        with torch.cuda.stream(side):
            side.wait_event(ev)
            h.barrier(channel=self._next_channel())                   # ... on every rank
            _lib.check(_lib.lib().tgr_multimem_allreduce_f32_capped(self._mc(self.sh_offset + first * row), count * row,
                                                                    h.rank, h.world_size, self.MAX_CTAS_OVERLAPPED,
                                                                    side.cuda_stream), "tgr_multimem_allreduce_f32_capped")

    def all_reduce_rest_and_wait(self, group=None):
        if not self.nvls:
            return super().all_reduce_rest_and_wait(group if group is not None else self.group)
        from . import _lib
        h, side = self.hdl, self._side()
        main = torch.cuda.current_stream(self.flat.device)
        ev = torch.cuda.Event()
        ev.record(main)                                               # the last range (and with it every small tensor) is written
        end = self.sh_offset if self.sh_offset is not None else self.flat.numel()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            h.barrier(channel=self._next_channel())
            if end > 0:                                               # everything that is not SH: one contiguous slice
                _lib.check(_lib.lib().tgr_multimem_allreduce_f32(self._mc(0), end, h.rank, h.world_size, side.cuda_stream),
                           "tgr_multimem_allreduce_f32")
            h.barrier(channel=self._next_channel())                   # every rank's slices have been broadcast
        main.wait_stream(side)
        return self


def render_batch_fwd_bwd(inp: Dict[str, torch.Tensor], cams: Sequence[Dict[str, object]], degree: int,
                         upstream, bucket: GradBucket, extras: bool = False, keep_images: bool = False):
    """Forward + backward of this rank's views; gradients accumulate in `bucket` (first view overwrites, the
    others add — no zero-fill, no separate accumulation pass).  `upstream(i, color, depth, alpha)` returns
    (dL_dcolor, dL_ddepth | None, dL_dalpha | None) for local view i.  Returns the rendered images if asked."""
    from . import rasterizer as rz
    e = torch.Tensor([])
    g = lambda k: inp[k] if inp.get(k) is not None else e
    images = []
    for i, cam in enumerate(cams):
        fwd = rz.c_rasterize_gaussians(
            cam["bg"], g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
            cam.get("scale_modifier", 1.0), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"],
            cam["tanfovy"], cam["image_height"], cam["image_width"], g("shs"), degree, cam["campos"], False, False,
            extras=extras)
        R, color, radii, geom, binning, img = fwd[:6]
        depth, alpha = (fwd[6], fwd[7]) if extras else (None, None)
        dLc, dLd, dLa = upstream(i, color, depth, alpha)
        kw = dict(accumulate_into=bucket.views) if i > 0 else dict(out=bucket.views)
        rz.c_rasterize_gaussians_backward(
            cam["bg"], g("means3D"), radii, g("colors_precomp"), g("scales"), g("rotations"),
            cam.get("scale_modifier", 1.0), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"],
            cam["tanfovy"], dLc, g("shs"), degree, cam["campos"], geom, R, binning, img, False,
            dL_dout_depth=dLd, dL_dout_alpha=dLa, **kw)
        if keep_images:
            images.append(color)
    return images


def settings_from_cam(cam: Dict[str, object], degree: int):
    """Camera dict (scene.orbit_camera) -> the reference's GaussianRasterizationSettings."""
    from .rasterizer import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=int(cam["image_height"]), image_width=int(cam["image_width"]), tanfovx=cam["tanfovx"],
        tanfovy=cam["tanfovy"], bg=cam["bg"], scale_modifier=cam.get("scale_modifier", 1.0),
        viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=degree, campos=cam["campos"],
        prefiltered=False, debug=False)


def render_views_fwd_bwd(inp: Dict[str, torch.Tensor], cams: Sequence[Dict[str, object]], degree: int, upstream,
                         bucket: GradBucket, extras: bool = False, n_streams: int = 0, accumulate: bool = False,
                         all_reduce: bool = False, comm_chunks: int = 4, group=None, comm_mode: str = "rows"):
    """Batched forward + backward of this rank's views (youreditableavatar_b200.multiview): one preprocess launch
    for all views, binning / blending fused per stage (n_streams = 0) or per view on `n_streams` streams, one
    backward-preprocess launch that writes the summed gradients into `bucket` (overwrite, or add with accumulate=True).
    `upstream(color[V,3,H,W], depth[V,1,H,W] | None, alpha | None[, view_events])` -> (dL_dcolor[V,3,H,W],
    dL_ddepth | None, dL_dalpha | None); `view_events[v]` fires when view v's images are complete (the forward has
    already been joined on the current stream when `upstream` runs, the events only matter to side streams).
    all_reduce=True also sums `bucket` over the ranks, overlapped with the backward: the per-Gaussian kernel runs
    over `comm_chunks` ranges of Gaussians and each range's gradient rows are all-reduced (NCCL, on the process
    group's stream; comm_mode "rows": all tensors' rows of the range in one coalesced launch, "sh": the SH rows
    per range and the small tensors at the end, "flat": no overlap) while the next range is computed.
    Returns the rendered images."""
    from . import multiview as mv
    e = torch.Tensor([])
    g = lambda k: inp[k] if inp.get(k) is not None else e
    settings = [settings_from_cam(c, degree) for c in cams]
    res = mv.c_rasterize_views(settings, g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
                               g("cov3D_precomp"), g("shs"), extras=extras, n_streams=n_streams)
    state, color = res[0], res[1]
    depth, alpha = (res[3], res[4]) if extras else (None, None)
    code = getattr(upstream, "__code__", None)       # (inspect.signature costs 30 us per call: a tenth of a C1 step's host time)
    wants_events = code is not None and (code.co_argcount - (1 if hasattr(upstream, "__self__") else 0)) >= 4
    if wants_events:
        dLc, dLd, dLa = upstream(color, depth, alpha, state.view_events)
    else:
        dLc, dLd, dLa = upstream(color, depth, alpha)
    kw = dict(accumulate_into=bucket.views) if accumulate else dict(out=bucket.views)
    if all_reduce and GradBucket._distributed(group):
        if comm_mode == "rows":     # every range carries all its gradient rows: only the last range is exposed
            mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, chunks=comm_chunks,
                                          on_chunk=lambda first, count: bucket.all_reduce_rows_async(first, count, group),
                                          **kw)
            bucket.wait()
        elif comm_mode == "sh":     # SH rows per range, the small tensors in one slice at the end
            mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, chunks=comm_chunks,
                                          on_chunk=lambda first, count: bucket.all_reduce_sh_rows_async(first, count, group),
                                          **kw)
            bucket.all_reduce_rest_and_wait(group)
        else:                       # "flat": one all-reduce of the whole buffer after the backward
            mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, **kw)
            bucket.all_reduce(group)
    else:
        mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, **kw)
    return color

"""Multi-view batches of the rasterizer: V cameras over the SAME Gaussians in one operator call.

The reference renders one view per call and per optimizer step (Edit_core/tetgs_texture/refine.py:54,
tetgs_scene/tetgs_model.py:467-614); its SDS / inpainting stages loop over views in Python.  On a B200 that leaves
two things on the table, which this module picks up (C ABI: tgr_*_batch in include/tetgs_rast.h):

  * the per-Gaussian kernels are HBM-bound and re-read 236 B of parameters per Gaussian and view (and the
    backward re-writes as much gradient per view).  Here one launch projects every Gaussian into all V views
    (`tgr_forward_preprocess_batch`) and one launch chains all V packed 2-D gradients back to the parameters
    (`tgr_backward_preprocess_batch`): parameters are read once, gradients written once per batch;
  * the per-view stages in between (depth sort, binning, tile sort, blending) are independent across views and
    individually leave SMs idle (latency-bound sorts, the tail of the heaviest tile); they are issued round-robin
    on a few CUDA streams so one view's tail is filled by the next view's kernels.

PyTorch is plumbing only: memory, streams/events, autograd bookkeeping.  There is no CPU path.
"""
from typing import List, Optional, Sequence

import ctypes as C
import torch

from . import _lib
from ._lib import TgrParams, check
from . import rasterizer as rz

__all__ = ["c_rasterize_views", "c_rasterize_views_backward", "rasterize_views", "MultiViewRasterizer", "ViewBatchState"]

_side = {}


def _side_streams(device: torch.device, n: int) -> List[torch.cuda.Stream]:
    key = device.index if device.index is not None else torch.cuda.current_device()
    pool = _side.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


_pinned_counts = {}


def _count_slots(device: torch.device, n: int) -> torch.Tensor:
    """Pinned 8-word slots {num_rendered, overflow, num_visible, prefilter_violation, key_or, key_and, -, -} for the
    asynchronous header read-back of a batch (double-buffered)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ent = _pinned_counts.get(key)
    if ent is None or ent[0].shape[1] < n:
        ent = [torch.zeros(2, max(n, 64), 8, dtype=torch.int32).pin_memory(), 0]
        _pinned_counts[key] = ent
    ent[1] ^= 1
    return ent[0][ent[1], :n]


def _read_counts(slots: torch.Tensor):
    """-> (instance counts per view, depth-key bits the batch's depth sorts have to look at)."""
    rows = slots.tolist()
    for r in rows:
        rz.check_prefilter(r)
    return [int(r[0]) for r in rows], max(rz.depth_bits_needed(r) for r in rows)


class ViewBatchState:
    """Everything the backward needs from a batched forward (the analogue of the tensors the reference's
    autograd Function saves, diff_gaussian_rasterization/__init__.py:97): parameter structs, the per-view
    instance counts and the opaque workspaces."""

    def __init__(self):
        self.params = None      # (TgrParams * V)
        self.counts = None      # list[int]: true instance counts per view
        self.caps = None        # list[int]: instance capacity each view's binning buffer was sized (and carved) for
        self.tensors = None     # keeps every buffer the structs point into alive
        self.V = 0
        self.extras = False
        self.n_streams = 1
        self.binding = None
        self.view_events = []   # one CUDA event per view, recorded when its forward outputs are complete


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


# Capacity hints: the binning buffers of a batch are sized from the instance counts, which only exist on the device
# after the preprocess kernel.  The first batch of a given shape waits for them (one host<->device sync, where the
# reference has one per view, rasterizer_impl.cu:281).  Every later batch of that shape sizes its buffers from the
# previous batch's largest count plus slack and launches depth sort / binning / blending WITHOUT waiting, so the GPU
# never idles between the per-Gaussian kernel and the sorts; the counts are read afterwards (the host blocks while the
# GPU already renders) and, should a view have outgrown its capacity (the kernels detect that on the device and
# render nothing), the batch is redone the synchronous way.
_capacity = {"enabled": True, "slack": 1.2, "margin": 65536, "hints": {}}
host_wait_seconds = 0.0   # time the host spent blocked in tgr_wait_num_rendered (diagnostics: issue time minus this is CPU work)


def _wait_counts(L):
    global host_wait_seconds
    import time
    t0 = time.perf_counter()
    check(L.tgr_wait_num_rendered(), "tgr_wait_num_rendered")
    host_wait_seconds += time.perf_counter() - t0


def set_capacity_hints(enabled: bool = True, slack: float = 1.2) -> None:
    """Turns the sync-free steady state of `c_rasterize_views` on / off (and forgets the recorded hints)."""
    _capacity["enabled"], _capacity["slack"] = bool(enabled), float(slack)
    _capacity["hints"].clear()


def c_rasterize_views(settings: Sequence, means3D, colors, opacity, scales, rotations, cov3D_precomp, sh,
                      extras: bool = False, n_streams: int = 0, binding=None, _use_hint: bool = True):
    """Forward of V views.  `settings` is a sequence of GaussianRasterizationSettings (same image size, SH degree
    and scale_modifier; cameras differ).  Returns (state, color[V,3,H,W], radii[V,P] int32[, depth[V,1,H,W],
    alpha[V,1,H,W]]).  The per-view results are bit-identical to V single-view `c_rasterize_gaussians` calls.
    n_streams = 0 (default): fused schedule — every per-view stage is one launch for the whole batch
    (tgr_forward_render_batch); n_streams >= 1: the per-view stages are issued view by view, round-robin on that
    many CUDA streams."""
    V = len(settings)
    if V == 0:
        raise RuntimeError("rasterize_views: empty batch")
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    rz._require_cuda(means3D)
    s0 = settings[0]
    H, W = int(s0.image_height), int(s0.image_width)
    for st in settings:
        if int(st.image_height) != H or int(st.image_width) != W or st.sh_degree != s0.sh_degree or \
                float(st.scale_modifier) != float(s0.scale_modifier):
            raise RuntimeError("rasterize_views: all views of a batch must share image size, SH degree and scale_modifier")
    device = means3D.device
    P = means3D.size(0)
    f32 = dict(dtype=torch.float32, device=device)
    u8 = dict(dtype=torch.uint8, device=device)
    state = ViewBatchState()
    state.V, state.extras, state.binding = V, extras, binding

    color = torch.empty(V, 3, H, W, **f32) if P else torch.zeros(V, 3, H, W, **f32)
    radii = torch.empty(V, P, dtype=torch.int32, device=device)
    depth = alpha = None
    if extras:
        depth = torch.empty(V, 1, H, W, **f32) if P else torch.zeros(V, 1, H, W, **f32)
        alpha = torch.empty(V, 1, H, W, **f32) if P else torch.zeros(V, 1, H, W, **f32)
    if P == 0:
        state.counts = state.caps = [0] * V
        return (state, color, radii) + ((depth, alpha) if extras else ())

    with torch.cuda.device(device):
        L = _lib.lib()
        means3D = rz._f32c(means3D, "means3D", device)
        colors = rz._f32c(colors, "colors_precomp", device)
        opacity = rz._f32c(opacity, "opacities", device)
        scales = rz._f32c(scales, "scales", device)
        rotations = rz._f32c(rotations, "rotations", device)
        cov3D_precomp = rz._f32c(cov3D_precomp, "cov3D_precomp", device)
        sh = rz._f32c(sh, "sh", device)

        gbytes = _align(L.tgr_geom_bytes(P))
        ibytes = _align(L.tgr_image_bytes(W, H))
        geom = torch.empty(V, gbytes, **u8)
        img = torch.empty(V, ibytes, **u8)
        slots = _count_slots(device, V)
        params = (TgrParams * V)()
        cams = []
        for v, st in enumerate(settings):
            bg = rz._f32c(st.bg, "bg", device)
            vm = rz._f32c(st.viewmatrix, "viewmatrix", device)
            pm = rz._f32c(st.projmatrix, "projmatrix", device)
            cp = rz._f32c(st.campos, "campos", device)
            cams.append((bg, vm, pm, cp))
            p = params[v]
            rz._fill_common(p, bg, means3D, colors, opacity, scales, rotations, st.scale_modifier, cov3D_precomp, vm, pm,
                            st.tanfovx, st.tanfovy, H, W, sh, st.sh_degree, cp, st.debug)
            p.prefiltered = 1 if st.prefiltered else 0
            p.extras = 1 if extras else 0
            p.geom_buffer, p.geom_bytes = geom[v].data_ptr(), gbytes
            p.image_buffer, p.image_bytes = img[v].data_ptr(), ibytes
            p.out_color = color[v].data_ptr()
            p.radii = radii[v].data_ptr()
            if extras:
                p.out_depth = depth[v].data_ptr()
                p.out_alpha = alpha[v].data_ptr()
            p.host_num_rendered = slots[v].data_ptr()

        main = torch.cuda.current_stream(device)
        bptr = C.byref(binding) if binding is not None else None
        check(L.tgr_forward_preprocess_batch(params, V, bptr, main.cuda_stream), "tgr_forward_preprocess_batch")
        S = max(0, min(int(n_streams), V))
        hkey = (device.index if device.index is not None else torch.cuda.current_device(), P, W, H, V)
        hint = _capacity["hints"].get(hkey) if (_capacity["enabled"] and _use_hint and S == 0) else None
        if hint is None:
            # the one host<->device sync of the batch (the reference has one per view, rasterizer_impl.cu:281)
            _wait_counts(L)
            counts, depth_bits = _read_counts(slots)
            cap_list = counts
        else:
            counts = None
            cap_list = [int(hint[0] * _capacity["slack"]) + _capacity["margin"]] * V
            depth_bits = min(32, hint[1] + 1)       # one bit of slack over what the previous batch of this shape needed
        for v in range(V):
            params[v].depth_key_bits = depth_bits

        binnings = []
        view_events = []
        for v in range(V):
            binning = torch.empty(L.tgr_binning_bytes(P, cap_list[v], W, H), **u8)
            binnings.append(binning)
            params[v].binning_buffer, params[v].binning_bytes = binning.data_ptr(), binning.numel()
        if S == 0:
            caps = (C.c_uint64 * V)(*cap_list)
            check(L.tgr_forward_render_batch(params, caps, V, main.cuda_stream), "tgr_forward_render_batch")
            if counts is None:
                # launched ahead of the counts: read them now (the GPU is already sorting / blending) and verify
                _wait_counts(L)
                counts, need_bits = _read_counts(slots)
                if max(counts) > cap_list[0] or need_bits > depth_bits:
                    _capacity["hints"].pop(hkey, None)   # outgrown: redo this batch with exactly sized buffers / sort bits
                    return c_rasterize_views(settings, means3D, colors, opacity, scales, rotations, cov3D_precomp, sh,
                                             extras=extras, n_streams=n_streams, binding=binding, _use_hint=False)
            else:
                need_bits = depth_bits
            if _capacity["enabled"]:
                _capacity["hints"][hkey] = (max(counts), need_bits)
            ev = torch.cuda.Event()
            ev.record(main)
            view_events = [ev] * V
        else:
            streams = _side_streams(device, S) if S > 1 else [main]
            fork = torch.cuda.Event()
            fork.record(main)
            for v in range(V):
                p = params[v]
                s = streams[v % S]
                if S > 1 and v < S:
                    s.wait_event(fork)
                check(L.tgr_forward_depth_sort(C.byref(p), s.cuda_stream), "tgr_forward_depth_sort")
                check(L.tgr_forward_render(C.byref(p), counts[v], s.cuda_stream), "tgr_forward_render")
                ev = torch.cuda.Event()
                ev.record(s)
                view_events.append(ev)   # view v's images are complete: lets callers drain them while later views render
            if S > 1:
                for s in streams:
                    main.wait_stream(s)

    state.params, state.counts, state.caps, state.n_streams = params, counts, list(cap_list), S
    state.view_events = view_events
    state.tensors = (means3D, colors, opacity, scales, rotations, cov3D_precomp, sh, cams, geom, img, binnings, radii)
    return (state, color, radii) + ((depth, alpha) if extras else ())


def c_rasterize_views_backward(state: ViewBatchState, dL_dout_color, dL_dout_depth=None, dL_dout_alpha=None,
                               accumulate_into=None, out=None, chunks: int = 1, on_chunk=None):
    """Backward of a batch: dL_dout_color[V,3,H,W] (+ depth / alpha gradients [V,1,H,W]) -> the 8-tuple of
    `_C.rasterize_gaussians_backward` (rasterize_points.cu:117-196), summed over the V views.  `out` /
    `accumulate_into` as in `c_rasterize_gaussians_backward`."""
    V = state.V
    means3D, colors, opacity, scales, rotations, cov3D_precomp, sh = state.tensors[:7] if state.tensors else (None,) * 7
    params = state.params
    if params is None:  # P == 0
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dL_dout_color.device)
        return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, 0, 3), z(0, 3), z(0, 4)
    device = means3D.device
    P = means3D.size(0)
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L = _lib.lib()
        dL_dout_color = rz._f32c(dL_dout_color, "dL_dout_color", device)
        if dL_dout_color.shape[0] != V:
            raise RuntimeError("dL_dout_color must have one image per view")
        if dL_dout_depth is not None:
            dL_dout_depth = rz._f32c(dL_dout_depth, "dL_dout_depth", device)
        if dL_dout_alpha is not None:
            dL_dout_alpha = rz._f32c(dL_dout_alpha, "dL_dout_alpha", device)
        sh_path = M > 0 and colors.numel() == 0
        if accumulate_into is not None or out is not None:
            grads = accumulate_into if accumulate_into is not None else out
        else:
            grads = (torch.empty(P, 3, **f32), torch.empty(P, 3, **f32), torch.empty(P, 1, **f32), torch.empty(P, 3, **f32),
                     torch.empty(P, 6, **f32), torch.empty(P, M, 3, **f32) if sh_path else torch.zeros(P, M, 3, **f32),
                     torch.empty(P, 3, **f32), torch.empty(P, 4, **f32))
        (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations) = grads
        use_extras = dL_dout_depth is not None or dL_dout_alpha is not None
        for v in range(V):
            p = params[v]
            p.extras = 1 if use_extras else 0
            p.dL_dout_color = dL_dout_color[v].data_ptr()
            p.dL_dout_depth = dL_dout_depth[v].data_ptr() if dL_dout_depth is not None else None
            p.dL_dout_alpha = dL_dout_alpha[v].data_ptr() if dL_dout_alpha is not None else None
        p0 = params[0]
        p0.accumulate = 1 if accumulate_into is not None else 0
        p0.dL_dmeans2D = rz._ptr(dL_dmeans2D)
        p0.dL_dcolors = rz._ptr(dL_dcolors)
        p0.dL_dopacity = rz._ptr(dL_dopacity)
        p0.dL_dmeans3D = rz._ptr(dL_dmeans3D)
        p0.dL_dcov3D = rz._ptr(dL_dcov3D)
        p0.dL_dsh = rz._ptr(dL_dsh) if sh_path else None
        p0.dL_dscales = rz._ptr(dL_dscales)
        p0.dL_drotations = rz._ptr(dL_drotations)

        main = torch.cuda.current_stream(device)
        S = state.n_streams
        caps = (C.c_uint64 * V)(*state.caps)
        if S == 0:
            check(L.tgr_backward_blend_batch(params, caps, V, main.cuda_stream), "tgr_backward_blend_batch")
        else:
            streams = _side_streams(device, S) if S > 1 else [main]
            if S > 1:
                fork = torch.cuda.Event()
                fork.record(main)
            for v in range(V):
                s = streams[v % S]
                if S > 1 and v < S:
                    s.wait_event(fork)
                check(L.tgr_backward_blend(C.byref(params[v]), state.caps[v], s.cuda_stream), "tgr_backward_blend")
            if S > 1:
                for s in streams:
                    main.wait_stream(s)
        caps = (C.c_uint64 * V)(*state.caps)
        bptr = C.byref(state.binding) if state.binding is not None else None
        nch = max(1, min(int(chunks), (P + 255) // 256))
        per = ((P + nch - 1) // nch + 255) // 256 * 256          # range starts must be multiples of 256
        first = 0
        while first < P:
            count = min(per, P - first)
            check(L.tgr_backward_preprocess_batch(params, caps, V, bptr, first, count, main.cuda_stream),
                  "tgr_backward_preprocess_batch")
            if on_chunk is not None:
                on_chunk(first, count)
            first += count
        state.keep_bwd = (dL_dout_color, dL_dout_depth, dL_dout_alpha)
    return grads


class _RasterizeViews(torch.autograd.Function):
    """Batched counterpart of `_RasterizeGaussians` (diff_gaussian_rasterization/__init__.py:48-155): same
    tensor arguments, a sequence of settings instead of one."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, settings,
                extras, n_streams):
        res = c_rasterize_views(settings, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, sh,
                                extras=extras, n_streams=n_streams)
        ctx.state = res[0]      # ctypes structs, counts, workspaces (opaque byte buffers autograd has no business with)
        ctx.extras = extras
        # the parameter tensors go through save_for_backward so that autograd's version counters catch an in-place
        # update between forward and backward (e.g. FusedAdam.step(), which writes parameters in place)
        ctx.save_for_backward(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        ctx.mark_non_differentiable(res[2])
        return tuple(res[1:])

    @staticmethod
    def backward(ctx, grad_color, _grad_radii, grad_depth=None, grad_alpha=None):
        _ = ctx.saved_tensors   # raises autograd's own errors: freed graph (second backward), in-place modification
        kw = dict(dL_dout_depth=grad_depth, dL_dout_alpha=grad_alpha) if ctx.extras else {}
        # the state is kept: with retain_graph=True the backward can run again (it re-zeroes its accumulators)
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rots) = c_rasterize_views_backward(
            ctx.state, grad_color, **kw)
        return (g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rots, g_cov3D, None, None, None)


def rasterize_views(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, settings,
                    extras=False, n_streams=0):
    return _RasterizeViews.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                 list(settings), extras, n_streams)


class MultiViewRasterizer(torch.nn.Module):
    """`GaussianRasterizer` for a batch of cameras: forward(...) takes the same keyword arguments
    (diff_gaussian_rasterization/__init__.py:186-220) and returns (color[V,3,H,W], radii[V,P]) — plus
    depth[V,1,H,W], alpha[V,1,H,W] with extra_outputs=True."""

    def __init__(self, raster_settings: Sequence, extra_outputs: bool = False, n_streams: int = 0):
        super().__init__()
        self.raster_settings = list(raster_settings)
        self.extra_outputs = extra_outputs
        self.n_streams = n_streams

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        return rasterize_views(means3D, means2D, e if shs is None else shs, e if colors_precomp is None else colors_precomp,
                               opacities, e if scales is None else scales, e if rotations is None else rotations,
                               e if cov3D_precomp is None else cov3D_precomp, self.raster_settings,
                               self.extra_outputs, self.n_streams)

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
    """One flat fp32 buffer laid out [means2D | colors | opacity | means3D | cov3D | sh | scales | rotations];
    `.views` are the eight gradient tensors aliasing it.  Offsets are multiples of 4 floats (16-byte vector
    stores in the kernels need the SH / rotation views 16-byte aligned)."""

    # what an optimizer over (means/delta, opacity, SH, scales, rotations) consumes; the other three exist in the
    # reference's return tuple only because its kernels produce them on the way
    TRAINING = ("dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations")

    def __init__(self, P: int, M: int, device="cuda", names: Optional[Sequence[str]] = None):
        self.P, self.M = P, M
        self.shapes = grad_shapes(P, M)
        self.names = tuple(GRAD_NAMES if names is None else names)
        offs, total = [], 0
        for name, shp in zip(GRAD_NAMES, self.shapes):
            if name not in self.names:
                offs.append(None)
                continue
            offs.append(total)
            n = 1
            for d in shp:
                n *= d
            total += (n + 3) // 4 * 4
        self.offsets = offs
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.views = tuple(None if o is None else self._view(o, shp) for o, shp in zip(offs, self.shapes))

    def _view(self, off, shp):
        n = 1
        for d in shp:
            n *= d
        return self.flat[off:off + n].view(*shp)

    def named(self) -> Dict[str, torch.Tensor]:
        return dict(zip(GRAD_NAMES, self.views))

    def all_reduce(self, group=None):
        """Sum over ranks — the only collective of a step."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
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
                         bucket: GradBucket, extras: bool = False, n_streams: int = 4, accumulate: bool = False):
    """Batched forward + backward of this rank's views (youreditableavatar_b200.multiview): one preprocess launch
    for all views, per-view binning / blending on `n_streams` streams, one backward-preprocess launch that writes
    the summed gradients into `bucket` (overwrite, or add with accumulate=True).
    `upstream(color[V,3,H,W], depth[V,1,H,W] | None, alpha | None[, view_events])` -> (dL_dcolor[V,3,H,W],
    dL_ddepth | None, dL_dalpha | None); `view_events[v]` fires when view v's images are complete (the forward has
    already been joined on the current stream when `upstream` runs, the events only matter to side streams).
    Returns the rendered images."""
    from . import multiview as mv
    e = torch.Tensor([])
    g = lambda k: inp[k] if inp.get(k) is not None else e
    settings = [settings_from_cam(c, degree) for c in cams]
    res = mv.c_rasterize_views(settings, g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
                               g("cov3D_precomp"), g("shs"), extras=extras, n_streams=n_streams)
    state, color = res[0], res[1]
    depth, alpha = (res[3], res[4]) if extras else (None, None)
    import inspect
    if len(inspect.signature(upstream).parameters) >= 4:
        dLc, dLd, dLa = upstream(color, depth, alpha, state.view_events)
    else:
        dLc, dLd, dLa = upstream(color, depth, alpha)
    kw = dict(accumulate_into=bucket.views) if accumulate else dict(out=bucket.views)
    mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, **kw)
    return color

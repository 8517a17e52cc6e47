"""Fused mesh binding: render straight from the raw TetGS parameters.

Replaces the eager PyTorch binding + activations the reference runs before every rasterizer call
(Edit_core/tetgs_scene/tetgs_model.py:252-286: points = ori + normals*delta, exp / normalize / sigmoid;
barycentric attributes :328-377) by the `BOUND` variants of the preprocess kernels (csrc/preprocess.cu,
csrc/preprocess_bwd.cu), reached through `tgr_binding` in the C ABI.  Gradients come back with respect to the
raw parameters (delta, log-scales, raw quaternions, opacity logits, SH) and optionally the mesh vertices.
"""
from typing import Dict, NamedTuple, Optional

import torch

from . import rasterizer as rz
from ._lib import TgrBinding


class MeshBinding(NamedTuple):
    verts: torch.Tensor         # [Nv,3] f32
    vert_normals: torch.Tensor  # [Nv,3] f32 unit vertex normals
    faces: torch.Tensor         # [Nf,3] i32
    face_index: torch.Tensor    # [P]   i32
    bary: torch.Tensor          # [P,3] f32

    @staticmethod
    def from_scene(gs: Dict[str, torch.Tensor], device="cuda") -> "MeshBinding":
        f = lambda k, dt: gs[k].to(device=device, dtype=dt).contiguous()
        return MeshBinding(f("verts", torch.float32), f("vert_normals", torch.float32), f("faces", torch.int32),
                           f("face_index", torch.int32), f("bary", torch.float32))


class DirectBinding(NamedTuple):
    """mean = origins + normals * delta with per-Gaussian constants — how every scene model of the reference holds its
    binding between re-meshings: `TetGS.ori_points` / `.normals` (tetgs_model.py:156-172, 252-258), the edit models'
    cat(keep_points, ori_edit_points + _edit_normals * _edit_points) (tetgs_edit_3d.py:272-280; keep Gaussians: their
    world position as origin, zero normal) and the fixed points of the flat 2-D Gaussians (tetgs_edit_2d.py:279-282:
    normals=None).  Gaussians [0, n_frozen) are the frozen `keep_*` part (requires_grad=False,
    tetgs_edit_2d.py:237-267): they get zero gradient rows and cost the backward nothing."""
    origins: torch.Tensor                 # [P,3] f32
    normals: Optional[torch.Tensor]       # [P,3] f32 or None
    n_frozen: int = 0

    @staticmethod
    def from_keep_edit(keep_xyz: torch.Tensor, edit_origins: torch.Tensor, edit_normals: Optional[torch.Tensor] = None,
                       device="cuda") -> "DirectBinding":
        """keep Gaussians first (frozen), then the edit Gaussians — the order of the reference's torch.cat calls."""
        f = lambda t: t.to(device=device, dtype=torch.float32)
        origins = torch.cat([f(keep_xyz), f(edit_origins)]).contiguous()
        normals = None
        if edit_normals is not None:
            normals = torch.cat([torch.zeros_like(f(keep_xyz)), f(edit_normals)]).contiguous()
        return DirectBinding(origins, normals, int(keep_xyz.shape[0]))


def _binding_struct(mesh, delta, log_scales, raw_quats, opacity_logits, act, grads=None) -> TgrBinding:
    b = TgrBinding()
    if isinstance(mesh, DirectBinding):
        b.origins = mesh.origins.data_ptr()
        b.normals = mesh.normals.data_ptr() if mesh.normals is not None else None
        b.n_frozen = int(mesh.n_frozen)
    else:
        b.n_verts, b.n_faces = mesh.verts.shape[0], mesh.faces.shape[0]
        b.verts, b.vert_normals = mesh.verts.data_ptr(), mesh.vert_normals.data_ptr()
        b.faces, b.face_index, b.bary = mesh.faces.data_ptr(), mesh.face_index.data_ptr(), mesh.bary.data_ptr()
    b.delta = delta.data_ptr() if delta is not None else None
    b.log_scales = log_scales.data_ptr()
    b.raw_quats, b.opacity_logits = raw_quats.data_ptr(), opacity_logits.data_ptr()
    b.out_means3D, b.out_scales = act["means3D"].data_ptr(), act["scales"].data_ptr()
    b.out_rotations, b.out_opacities = act["rotations"].data_ptr(), act["opacities"].data_ptr()
    if grads is not None:
        b.dL_ddelta = grads["delta"].data_ptr() if grads.get("delta") is not None else None
        b.dL_dlog_scales = grads["log_scales"].data_ptr()
        b.dL_draw_quats, b.dL_dopacity_logits = grads["raw_quats"].data_ptr(), grads["opacity_logits"].data_ptr()
        if grads.get("verts") is not None:
            b.dL_dverts = grads["verts"].data_ptr()
    return b


class _RasterizeBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, delta, log_scales, raw_quats, opacity_logits, shs, mesh, raster_settings, extras, verts_grad):
        P = log_scales.shape[0]
        dev = log_scales.device
        # the reference keeps delta and the opacity logits as [P,1] (`_points`, `all_densities`, tetgs_model.py:172,202);
        # gradients go back in whatever shape / dtype the caller's tensors have.  delta=None: no offsets (the flat 2-D
        # Gaussians of tetgs_edit_2d.py sit at fixed points)
        ctx.in_meta = tuple((t.shape, t.dtype) if t is not None else None for t in (delta, log_scales, raw_quats, opacity_logits, shs))
        c = lambda t: t.detach().contiguous().float()
        log_scales, raw_quats, opacity_logits, shs = map(c, (log_scales, raw_quats, opacity_logits.reshape(-1), shs))
        delta = c(delta.reshape(-1)) if delta is not None else None
        f32 = dict(dtype=torch.float32, device=dev)
        act = {"means3D": torch.empty(P, 3, **f32), "scales": torch.empty(P, 3, **f32),
               "rotations": torch.empty(P, 4, **f32), "opacities": torch.empty(P, 1, **f32)}
        b = _binding_struct(mesh, delta, log_scales, raw_quats, opacity_logits, act)
        s = raster_settings
        e = torch.Tensor([])
        res = rz.c_rasterize_gaussians(s.bg, act["means3D"], e, act["opacities"], act["scales"], act["rotations"],
                                       s.scale_modifier, e, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy,
                                       s.image_height, s.image_width, shs, s.sh_degree, s.campos, s.prefiltered, s.debug,
                                       extras=extras, binding=b)
        R, color, radii, geom, binning, img = res[:6]
        ctx.settings, ctx.R, ctx.extras, ctx.mesh, ctx.verts_grad = s, R, extras, mesh, verts_grad
        ctx.has_delta = delta is not None
        ctx.save_for_backward(delta if delta is not None else log_scales.new_empty(0), log_scales, raw_quats, opacity_logits,
                              shs, radii, geom, binning, img, act["means3D"], act["scales"], act["rotations"], act["opacities"])
        ctx.mark_non_differentiable(radii)
        if extras:
            return color, radii, res[6], res[7]
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth=None, g_alpha=None):
        (delta, log_scales, raw_quats, opacity_logits, shs, radii, geom, binning, img, means3D, scales, rotations,
         opacities) = ctx.saved_tensors
        s, mesh = ctx.settings, ctx.mesh
        P, dev = log_scales.shape[0], log_scales.device
        f32 = dict(dtype=torch.float32, device=dev)
        if not ctx.has_delta:
            delta = None
        grads = {"delta": torch.empty(P, **f32) if ctx.has_delta else None, "log_scales": torch.empty(P, 3, **f32),
                 "raw_quats": torch.empty(P, 4, **f32), "opacity_logits": torch.empty(P, **f32),
                 "verts": torch.zeros_like(mesh.verts) if (ctx.verts_grad and not isinstance(mesh, DirectBinding)) else None}
        act = {"means3D": means3D, "scales": scales, "rotations": rotations, "opacities": opacities}
        b = _binding_struct(mesh, delta, log_scales, raw_quats, opacity_logits, act, grads)
        e = torch.Tensor([])
        kw = dict(dL_dout_depth=g_depth, dL_dout_alpha=g_alpha) if ctx.extras else {}
        out = rz.c_rasterize_gaussians_backward(s.bg, means3D, radii, e, scales, rotations, s.scale_modifier, e,
                                                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, g_color, shs,
                                                s.sh_degree, s.campos, geom, ctx.R, binning, img, s.debug, binding=b,
                                                **kw)
        ctx.extra_grads = grads
        _RasterizeBound.last_vertex_grad = grads["verts"]
        back = (grads["delta"], grads["log_scales"], grads["raw_quats"], grads["opacity_logits"], out[5])
        back = tuple(g.reshape(m[0]).to(m[1]) if (g is not None and m is not None) else None for g, m in zip(back, ctx.in_meta))
        return back + (None, None, None, None)


def rasterize_bound(delta, log_scales, raw_quats, opacity_logits, shs, mesh: MeshBinding, raster_settings,
                    extras: bool = False, verts_grad: bool = False):
    """color, radii (, depth, alpha) = render of Gaussians bound to `mesh`; differentiable w.r.t. the five raw
    parameter tensors.  With verts_grad=True the vertex gradient of the last backward is available as
    `_RasterizeBound.last_vertex_grad` (vertices are frozen in the reference's texture stages)."""
    return _RasterizeBound.apply(delta, log_scales, raw_quats, opacity_logits, shs, mesh, raster_settings, extras,
                                 verts_grad)

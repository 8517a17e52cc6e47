"""Host-side mirror of the reference rasterizer's operator API, on top of the C ABI.

Mirrors (same names, argument order, return arity and error behaviour):
  * `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward`, `_C.mark_visible`
    (Edit_core/thirdparties/diff-gaussian-rasterization/ext.cpp:15-19, rasterize_points.cu:35-217)
  * `rasterize_gaussians`, `_RasterizeGaussians`, `GaussianRasterizationSettings`, `GaussianRasterizer`
    (diff_gaussian_rasterization/__init__.py:21-220)
New, opt-in: `GaussianRasterizer(settings, extra_outputs=True)` additionally returns depth and alpha images.

PyTorch is used only for device memory, streams and autograd plumbing; all computation is in
libtetgs_rast.so (hand-written sm_100a CUDA).  There is no CPU path: non-CUDA inputs raise.
"""
from typing import NamedTuple, Optional

import ctypes as C
import torch
import torch.nn as nn

from . import _lib
from ._lib import TgrParams, TgrBinding, check

__all__ = [
    "GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
    "c_rasterize_gaussians", "c_rasterize_gaussians_backward", "c_mark_visible",
]

_pinned = {}


PREFILTER_MESSAGE = "Point is filtered although prefiltered is set. This shouldn't happen!"   # auxiliary.h:158


def _pinned_slot(device: torch.device) -> torch.Tensor:
    """Ring of pinned 8-word slots per device for the asynchronous read-back of the geom header
    {num_rendered, overflow, num_visible, prefilter_violation, key_or, key_and, -, -}."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ent = _pinned.get(key)
    if ent is None:
        ent = [torch.zeros(64, 8, dtype=torch.int32).pin_memory(), 0]
        _pinned[key] = ent
    ent[1] = (ent[1] + 1) % 64
    return ent[0][ent[1]]


def depth_bits_needed(header_words) -> int:
    """Number of low bits in which the fp32 depth keys of the visible Gaussians differ (words [4], [5] of the header
    mirror are their OR / AND): all the depth sort has to look at."""
    diff = (int(header_words[4]) ^ int(header_words[5])) & 0xffffffff
    return max(1, diff.bit_length())


# depth-key bits of the previous call per (device, P): the depth sort of a single-view call is launched before the
# header comes back, so it sorts as many bits (+ 1 of slack) as the last view of the same scene needed; should this view
# need more, the forward starts over with the exact number before anything consumes the wrong order
_depth_bits_hint = {}

# Capacity hints of the single-view calls, per (device, P, W, H): the reference stalls the GPU once per forward — it
# reads num_rendered back to size the binning buffer before it can launch the sort (rasterizer_impl.cu:277-281).  Here
# the first call of a shape does the same; every later one sizes the buffer from the previous count x slack and launches
# binning / sort / blending right behind the preprocess kernel.  The host still reads the count (the operator returns
# it), but only AFTER everything is queued: the GPU renders while the host waits, and should the view have outgrown
# its buffer (the kernels notice on the device and render nothing) the call is redone the synchronous way.
_capacity = {"enabled": True, "slack": 1.25, "margin": 32768, "hints": {}}


def set_capacity_hints(enabled: bool = True) -> None:
    _capacity["enabled"] = bool(enabled)
    _capacity["hints"].clear()


def check_prefilter(header_words) -> None:
    """The reference prints this message and __trap()s (which kills the CUDA context) when `prefiltered=True` was
    promised but a Gaussian fails the near-plane test (auxiliary.h:154-160); here it is an exception."""
    if int(header_words[3]) != 0:
        raise RuntimeError(PREFILTER_MESSAGE)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, name: str, device: torch.device) -> torch.Tensor:
    """fp32 + contiguous + same device (absent optionals arrive as 0-element CPU tensors and pass through)."""
    if t.numel() == 0:
        return t
    if t.device != device:
        raise RuntimeError("%s is on %s but means3D is on %s" % (name, t.device, device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "tetgs rasterizer: means3D must be a CUDA tensor — this library has no CPU fallback "
            "(got device %s)" % t.device)


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _fill_common(p: TgrParams, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                 viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, debug):
    P = means3D.size(0)
    p.P = P
    p.D = int(degree)
    p.M = int(sh.size(1)) if sh.numel() != 0 else 0
    p.W = int(W)
    p.H = int(H)
    p.tan_fovx = float(tan_fovx)
    p.tan_fovy = float(tan_fovy)
    p.scale_modifier = float(scale_modifier)
    p.debug = 1 if debug else 0
    p.background = _ptr(bg)
    p.viewmatrix = _ptr(viewmatrix)
    p.projmatrix = _ptr(projmatrix)
    p.campos = _ptr(campos)
    p.means3D = _ptr(means3D)
    p.shs = _ptr(sh)
    p.colors_precomp = _ptr(colors)
    p.opacities = _ptr(opacity)
    p.scales = _ptr(scales)
    p.rotations = _ptr(rotations)
    p.cov3D_precomp = _ptr(cov3D_precomp)


def c_rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                          viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                          prefiltered, debug, extras=False, binding=None, _use_hint=True):
    """`_C.rasterize_gaussians` (rasterize_points.cu:35-115): same 19 positional args, same 6-tuple result
    (num_rendered, color[3,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer); with extras=True the
    tuple is extended by (depth[1,H,W], alpha[1,H,W])."""
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    _require_cuda(means3D)
    device = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    f32 = dict(dtype=torch.float32, device=device)
    u8 = dict(dtype=torch.uint8, device=device)

    if P == 0:  # rasterize_points.cu:81 — nothing is launched, outputs stay zero, buffers empty
        out = (0, torch.zeros(3, H, W, **f32), torch.zeros(0, dtype=torch.int32, device=device),
               torch.empty(0, **u8), torch.empty(0, **u8), torch.empty(0, **u8))
        if extras:
            out = out + (torch.zeros(1, H, W, **f32), torch.zeros(1, H, W, **f32))
        return out

    with torch.cuda.device(device):
        L = _lib.lib()
        means3D = _f32c(means3D, "means3D", device)
        bg = _f32c(bg, "bg", device)
        colors = _f32c(colors, "colors_precomp", device)
        opacity = _f32c(opacity, "opacities", device)
        scales = _f32c(scales, "scales", device)
        rotations = _f32c(rotations, "rotations", device)
        cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp", device)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", device)
        projmatrix = _f32c(projmatrix, "projmatrix", device)
        sh = _f32c(sh, "sh", device)
        campos = _f32c(campos, "campos", device)

        out_color = torch.empty(3, H, W, **f32)
        radii = torch.empty(P, dtype=torch.int32, device=device)
        geom = torch.empty(L.tgr_geom_bytes(P), **u8)
        img = torch.empty(L.tgr_image_bytes(W, H), **u8)
        depth = alpha = None
        if extras:
            depth = torch.empty(1, H, W, **f32)
            alpha = torch.empty(1, H, W, **f32)
        slot = _pinned_slot(device)

        p = TgrParams()
        _fill_common(p, bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                     projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, debug)
        p.prefiltered = 1 if prefiltered else 0
        p.extras = 1 if extras else 0
        p.geom_buffer, p.geom_bytes = geom.data_ptr(), geom.numel()
        p.image_buffer, p.image_bytes = img.data_ptr(), img.numel()
        p.out_color = out_color.data_ptr()
        p.radii = radii.data_ptr()
        p.out_depth = _ptr(depth)
        p.out_alpha = _ptr(alpha)
        p.host_num_rendered = slot.data_ptr()
        stream = _stream(device)
        bptr = C.byref(binding) if binding is not None else None
        hkey = (device.index if device.index is not None else torch.cuda.current_device(), P)
        ckey = hkey + (W, H)
        p.depth_key_bits = min(32, _depth_bits_hint.get(hkey, 31) + 1)
        hint = _capacity["hints"].get(ckey) if (_capacity["enabled"] and _use_hint) else None

        check(L.tgr_forward_preprocess(C.byref(p), bptr, stream), "tgr_forward_preprocess")
        cap = None
        if hint is not None:
            # launch-ahead: binning buffer from the previous count, everything queued before the host looks at the count
            cap = (int(hint * _capacity["slack"]) + _capacity["margin"] + 31) // 32 * 32
            binning = torch.empty(L.tgr_binning_bytes(P, cap, W, H), **u8)
            p.binning_buffer, p.binning_bytes = binning.data_ptr(), binning.numel()
            check(L.tgr_forward_render(C.byref(p), cap, stream), "tgr_forward_render")
        # the host reads the count: the operator returns it (the reference stalls the GPU here, rasterizer_impl.cu:281;
        # with a hint the GPU is already sorting and blending while the host waits)
        check(L.tgr_wait_num_rendered(), "tgr_wait_num_rendered")
        R = int(slot[0])
        check_prefilter(slot)
        need = depth_bits_needed(slot)
        _depth_bits_hint[hkey] = need
        if _capacity["enabled"]:
            _capacity["hints"][ckey] = R
        if cap is not None and R > cap:    # outgrown: nothing was rendered (device-side check); redo with the exact size
            return c_rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                                         viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                                         campos, prefiltered, debug, extras=extras, binding=binding, _use_hint=False)
        if need > p.depth_key_bits:        # this view spans more depth bits than the hint covered: the order is wrong.
            # Start over (the hint now holds what this view needs); rare — the camera would have to move from a view
            # where all visible depths share their leading bits to one where they do not
            return c_rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                                         viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                                         campos, prefiltered, debug, extras=extras, binding=binding, _use_hint=_use_hint)
        if cap is None:
            binning = torch.empty(L.tgr_binning_bytes(P, R, W, H), **u8)
            p.binning_buffer, p.binning_bytes = binning.data_ptr(), binning.numel()
            check(L.tgr_forward_render(C.byref(p), R, stream), "tgr_forward_render")

    out = (R, out_color, radii, geom, binning, img)
    if extras:
        out = out + (depth, alpha)
    return out


def c_rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                   viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
                                   geomBuffer, R, binningBuffer, imageBuffer, debug, dL_dout_depth=None,
                                   dL_dout_alpha=None, binding=None, accumulate_into=None, out=None):
    """`_C.rasterize_gaussians_backward` (rasterize_points.cu:117-196): same 21 positional args, same 8-tuple
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations).
    New: accumulate_into=<that 8-tuple from an earlier call> adds this view's gradients into those tensors
    inside the kernels (multi-view batches) and returns them; out=<8-tuple> overwrites caller-provided tensors
    (e.g. views into one flat all-reduce buffer) instead of allocating."""
    _require_cuda(means3D)
    device = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=device)
    if P == 0:
        z = lambda *s: torch.zeros(*s, **f32)
        return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, M, 3), z(0, 3), z(0, 4)

    with torch.cuda.device(device):
        L = _lib.lib()
        means3D = _f32c(means3D, "means3D", device)
        bg = _f32c(bg, "bg", device)
        colors = _f32c(colors, "colors_precomp", device)
        scales = _f32c(scales, "scales", device)
        rotations = _f32c(rotations, "rotations", device)
        cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp", device)
        viewmatrix = _f32c(viewmatrix, "viewmatrix", device)
        projmatrix = _f32c(projmatrix, "projmatrix", device)
        sh = _f32c(sh, "sh", device)
        campos = _f32c(campos, "campos", device)
        dL_dout_color = _f32c(dL_dout_color, "dL_dout_color", device)
        radii = radii.contiguous()

        sh_path = M > 0 and colors.numel() == 0
        if accumulate_into is not None or out is not None:
            (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
             dL_drotations) = accumulate_into if accumulate_into is not None else out
        else:
            # every element is written by the kernels -> empty, not zeros (the reference fills nine zero tensors)
            dL_dmeans3D = torch.empty(P, 3, **f32)
            dL_dmeans2D = torch.empty(P, 3, **f32)
            dL_dcolors = torch.empty(P, 3, **f32)
            dL_dopacity = torch.empty(P, 1, **f32)
            dL_dcov3D = torch.empty(P, 6, **f32)
            dL_dsh = torch.empty(P, M, 3, **f32) if sh_path else torch.zeros(P, M, 3, **f32)
            dL_dscales = torch.empty(P, 3, **f32)
            dL_drotations = torch.empty(P, 4, **f32)

        p = TgrParams()
        _fill_common(p, bg, means3D, colors, None_t, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                     projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, debug)
        p.extras = 1 if (dL_dout_depth is not None or dL_dout_alpha is not None) else 0
        p.accumulate = 1 if accumulate_into is not None else 0
        p.geom_buffer, p.geom_bytes = geomBuffer.data_ptr(), geomBuffer.numel()
        p.binning_buffer, p.binning_bytes = binningBuffer.data_ptr(), binningBuffer.numel()
        p.image_buffer, p.image_bytes = imageBuffer.data_ptr(), imageBuffer.numel()
        p.radii = radii.data_ptr()
        p.dL_dout_color = dL_dout_color.data_ptr()
        if dL_dout_depth is not None:
            dL_dout_depth = _f32c(dL_dout_depth, "dL_dout_depth", device)
            p.dL_dout_depth = dL_dout_depth.data_ptr()
        if dL_dout_alpha is not None:
            dL_dout_alpha = _f32c(dL_dout_alpha, "dL_dout_alpha", device)
            p.dL_dout_alpha = dL_dout_alpha.data_ptr()
        # a None entry (only possible through out= / accumulate_into=) means "do not produce this gradient"
        p.dL_dmeans2D = _ptr(dL_dmeans2D)
        p.dL_dcolors = _ptr(dL_dcolors)
        p.dL_dopacity = _ptr(dL_dopacity)
        p.dL_dmeans3D = _ptr(dL_dmeans3D)
        p.dL_dcov3D = _ptr(dL_dcov3D)
        p.dL_dsh = _ptr(dL_dsh) if sh_path else None
        p.dL_dscales = _ptr(dL_dscales)
        p.dL_drotations = _ptr(dL_drotations)
        bptr = C.byref(binding) if binding is not None else None
        # the binning buffer may be larger than R asks for (forward launched from a capacity hint): its layout follows
        # from its size, as the reference's follows from R (rasterizer_impl.cu:371-373)
        cap = L.tgr_binning_capacity(P, binningBuffer.numel(), W, H)
        if cap < int(R):
            raise RuntimeError("binningBuffer (%d bytes) is too small for %d rendered instances" % (binningBuffer.numel(), int(R)))
        check(L.tgr_backward(C.byref(p), bptr, cap, _stream(device)), "tgr_backward")

    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


None_t = torch.empty(0)


def c_mark_visible(means3D, viewmatrix, projmatrix):
    """`_C.mark_visible` (rasterize_points.cu:198-217) -> bool[P]."""
    _require_cuda(means3D)
    device = means3D.device
    P = means3D.size(0)
    present = torch.zeros(P, dtype=torch.bool, device=device)
    if P != 0:
        with torch.cuda.device(device):
            m = _f32c(means3D, "means3D", device)
            v = _f32c(viewmatrix, "viewmatrix", device)
            pr = _f32c(projmatrix, "projmatrix", device)
            check(_lib.lib().tgr_mark_visible(P, m.data_ptr(), v.data_ptr(), pr.data_ptr(), present.data_ptr(),
                                              _stream(device)), "tgr_mark_visible")
    return present


# ------------------------------------------------------------------------------------------------
# autograd Function + module, mirroring diff_gaussian_rasterization/__init__.py:17-220
# ------------------------------------------------------------------------------------------------
def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _fwd_args(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp):
    return (
        raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations, raster_settings.scale_modifier,
        cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
        raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width, sh,
        raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug,
    )


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, extras=False):
        args = _fwd_args(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                res = c_rasterize_gaussians(*args, extras=extras)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            res = c_rasterize_gaussians(*args, extras=extras)
        num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = res[:6]
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.extras = extras
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        if extras:
            return color, radii, res[6], res[7]
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_out_depth=None, grad_out_alpha=None):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, grad_out_color, sh, raster_settings.sh_degree, raster_settings.campos,
                geomBuffer, num_rendered, binningBuffer, imgBuffer, raster_settings.debug)
        kw = {}
        if ctx.extras:
            kw = dict(dL_dout_depth=grad_out_depth, dL_dout_alpha=grad_out_alpha)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads8 = c_rasterize_gaussians_backward(*args, **kw)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads8 = c_rasterize_gaussians_backward(*args, **kw)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = grads8
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, extras=False):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, extras)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, extra_outputs: bool = False):
        super().__init__()
        self.raster_settings = raster_settings
        self.extra_outputs = extra_outputs

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = c_mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings, self.extra_outputs)

"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes as C

import numpy as np
import torch

from youreditableavatar_b200 import _lib, scene
from youreditableavatar_b200._lib import TgrParams, check
from youreditableavatar_b200 import rasterizer as rz


def to_dev(d, device):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def small_scene(P=3000, g=32, res=128, view=0, nviews=4, device="cuda", seed=0, sh_coeffs=16):
    gs = scene.make_scene(P, g=g, seed=seed, device="cpu", sh_coeffs=sh_coeffs)
    act = scene.activate(gs)
    cam = scene.orbit_camera(view, nviews, res, res, device="cpu")
    return gs, to_dev(act, device), to_dev(cam, device)


def random_cloud(P, res_w, res_h, device="cuda", seed=0, big=False):
    """Unstructured Gaussians in front of (and partly behind / outside) a camera: exercises culling,
    clamping, large radii and ragged image sizes."""
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(P, 3, generator=g) - 0.5) * torch.tensor([3.0, 3.0, 3.0])
    scales = torch.exp(torch.randn(P, 3, generator=g) * 0.7 + (-2.0 if big else -3.5))
    rots = torch.randn(P, 4, generator=g)
    rots = rots / rots.norm(dim=-1, keepdim=True) * (1 + 0.05 * torch.randn(P, 1, generator=g))
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 2)
    shs = torch.randn(P, 16, 3, generator=g) * 0.3
    cam = scene.orbit_camera(1, 5, res_h, res_w, radius=2.0, device="cpu")
    inp = {"means3D": means, "scales": scales, "rotations": rots, "opacities": opac, "shs": shs}
    return to_dev(inp, device), to_dev(cam, device)


def ours_forward(inp, cam, degree, extras=False):
    e = torch.Tensor([])
    g = lambda k: inp[k] if inp.get(k) is not None else e
    return rz.c_rasterize_gaussians(
        cam["bg"], g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
        cam.get("scale_modifier", 1.0), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"],
        cam["tanfovy"], cam["image_height"], cam["image_width"], g("shs"), degree, cam["campos"], False, False,
        extras=extras)


def ours_backward(inp, cam, degree, fwd, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
    e = torch.Tensor([])
    g = lambda k: inp[k] if inp.get(k) is not None else e
    R, color, radii, geom, binning, img = fwd[:6]
    return rz.c_rasterize_gaussians_backward(
        cam["bg"], g("means3D"), radii, g("colors_precomp"), g("scales"), g("rotations"),
        cam.get("scale_modifier", 1.0), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"],
        cam["tanfovy"], dL_dcolor, g("shs"), degree, cam["campos"], geom, R, binning, img, False,
        dL_dout_depth=dL_ddepth, dL_dout_alpha=dL_dalpha)


def _params_for_export(P, W, H, geom, binning, img):
    p = TgrParams()
    p.P, p.W, p.H = P, W, H
    p.geom_buffer, p.geom_bytes = geom.data_ptr(), geom.numel()
    p.image_buffer, p.image_bytes = img.data_ptr(), img.numel()
    if binning is not None and binning.numel() > 0:
        p.binning_buffer, p.binning_bytes = binning.data_ptr(), binning.numel()
    return p


def export_binning(P, W, H, fwd, cap=None):
    """cap: instance capacity the binning buffer was sized for (default: recovered from the buffer's size)."""
    R, color, radii, geom, binning, img = fwd[:6]
    dev = geom.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = torch.empty(R, dtype=torch.int64, device=dev)
    ids = torch.empty(R, dtype=torch.int32, device=dev)
    ranges = torch.empty(T, 2, dtype=torch.int32, device=dev)
    p = _params_for_export(P, W, H, geom, binning, img)
    st = torch.cuda.current_stream(dev).cuda_stream
    if cap is None:   # a forward launched from a capacity hint has a buffer larger than R: the layout follows from its size
        cap = _lib.lib().tgr_binning_capacity(P, binning.numel(), W, H)
    check(_lib.lib().tgr_export_binning(C.byref(p), cap, R, keys.data_ptr(), ids.data_ptr(),
                                        ranges.data_ptr(), st))
    torch.cuda.synchronize()
    return keys, ids, ranges


def export_geom(P, W, H, fwd):
    R, color, radii, geom, binning, img = fwd[:6]
    dev = geom.device
    f = dict(dtype=torch.float32, device=dev)
    depth, xy, co, rgb = torch.empty(P, **f), torch.empty(P, 2, **f), torch.empty(P, 4, **f), torch.empty(P, 3, **f)
    tiles = torch.empty(P, dtype=torch.int32, device=dev)
    p = _params_for_export(P, W, H, geom, None, img)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().tgr_export_geom(C.byref(p), depth.data_ptr(), xy.data_ptr(), co.data_ptr(), rgb.data_ptr(),
                                     tiles.data_ptr(), st))
    torch.cuda.synchronize()
    return {"depths": depth, "means2D": xy, "conic_opacity": co, "rgb": rgb, "tiles_touched": tiles}


def export_image_state(P, W, H, fwd):
    R, color, radii, geom, binning, img = fwd[:6]
    dev = geom.device
    fT = torch.empty(H * W, dtype=torch.float32, device=dev)
    nc = torch.empty(H * W, dtype=torch.int32, device=dev)
    p = _params_for_export(P, W, H, geom, None, img)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().tgr_export_image_state(C.byref(p), fT.data_ptr(), nc.data_ptr(), st))
    torch.cuda.synchronize()
    return fT, nc


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    d = (a - b).norm()
    n = b.norm()
    return float(d / n) if n > 0 else float(d)

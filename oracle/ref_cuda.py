"""Loader for the UNMODIFIED reference CUDA extensions built by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY
(imported by tests/, smoke() and `bench.py --impl reference`; never by the product package).

`ref_dgr_C` exposes the reference's own pybind entry points (ext.cpp:15-19); the helpers below call them with
the reference's argument order and decode its opaque byte buffers (layout: rasterizer_impl.h:21-72,
rasterizer_impl.cu:155-194) so keys / sorted ids / tile ranges / n_contrib can be compared bit for bit.
"""
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_DGR = os.path.join(_HERE, "_ref", "dgr", "ref_dgr_C.so")
_KNN = os.path.join(_HERE, "_ref", "knn", "ref_knn_C.so")
_mods = {}


def available() -> bool:
    return os.path.exists(_DGR) and os.path.exists(_KNN)


def _load(name, path):
    if name not in _mods:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def dgr():
    return _load("ref_dgr_C", _DGR)


def knn():
    return _load("ref_knn_C", _KNN)


def _e(device):
    # absent optionals arrive exactly as the reference's own Python wrapper passes them (__init__.py:197-207)
    return torch.Tensor([])


def forward(inp, cam, degree):
    """inp: dict with means3D, opacities and (scales, rotations | cov3D_precomp), (shs | colors_precomp)."""
    dev = inp["means3D"].device
    g = lambda k: inp[k].contiguous() if inp.get(k) is not None else _e(dev)
    return dgr().rasterize_gaussians(
        cam["bg"], g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
        float(cam.get("scale_modifier", 1.0)), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"],
        float(cam["tanfovx"]), float(cam["tanfovy"]), int(cam["image_height"]), int(cam["image_width"]), g("shs"),
        int(degree), cam["campos"], False, False)


def backward(inp, cam, degree, fwd_out, dL_dcolor):
    R, color, radii, geom, binning, img = fwd_out
    dev = inp["means3D"].device
    g = lambda k: inp[k].contiguous() if inp.get(k) is not None else _e(dev)
    return dgr().rasterize_gaussians_backward(
        cam["bg"], g("means3D"), radii, g("colors_precomp"), g("scales"), g("rotations"),
        float(cam.get("scale_modifier", 1.0)), g("cov3D_precomp"), cam["viewmatrix"], cam["projmatrix"],
        float(cam["tanfovx"]), float(cam["tanfovy"]), dL_dcolor.contiguous(), g("shs"), int(degree), cam["campos"],
        geom, R, binning, img, False)


def _align128(a: int) -> int:
    return (a + 127) & ~127


def _view(buf: torch.Tensor, addr: int, nbytes: int, dtype) -> torch.Tensor:
    off = addr - buf.data_ptr()
    return buf[off:off + nbytes].view(dtype)


def decode_binning(binning: torch.Tensor, R: int):
    """-> (sorted keys uint64 as int64 tensor, sorted ids int32 tensor). rasterizer_impl.cu:181-194"""
    a = _align128(binning.data_ptr())
    point_list = _view(binning, a, 4 * R, torch.int32)
    a = _align128(a + 4 * R)            # point_list_unsorted
    a = _align128(a + 4 * R)            # point_list_keys
    keys = _view(binning, a, 8 * R, torch.int64)
    return keys, point_list


def decode_image(img: torch.Tensor, N: int, T: int):
    """-> (accum_alpha f32[N], n_contrib i32[N], ranges i32[T,2]). rasterizer_impl.cu:172-179"""
    a = _align128(img.data_ptr())
    accum = _view(img, a, 4 * N, torch.float32)
    a = _align128(a + 4 * N)
    ncon = _view(img, a, 4 * N, torch.int32)
    a = _align128(a + 4 * N)
    ranges = _view(img, a, 8 * T, torch.int32).reshape(T, 2)
    return accum, ncon, ranges


def decode_geom(geom: torch.Tensor, P: int):
    """-> dict(depths, means2D, conic_opacity, rgb, tiles_touched). rasterizer_impl.cu:155-170"""
    a = _align128(geom.data_ptr())
    depths = _view(geom, a, 4 * P, torch.float32)
    a = _align128(a + 4 * P)            # clamped (bool x 3P)
    a = _align128(a + 3 * P)            # internal_radii
    a = _align128(a + 4 * P)            # means2D
    means2D = _view(geom, a, 8 * P, torch.float32).reshape(P, 2)
    a = _align128(a + 8 * P)            # cov3D
    a = _align128(a + 24 * P)           # conic_opacity
    co = _view(geom, a, 16 * P, torch.float32).reshape(P, 4)
    a = _align128(a + 16 * P)           # rgb
    rgb = _view(geom, a, 12 * P, torch.float32).reshape(P, 3)
    a = _align128(a + 12 * P)           # tiles_touched
    tiles = _view(geom, a, 4 * P, torch.int32)
    return {"depths": depths, "means2D": means2D, "conic_opacity": co, "rgb": rgb, "tiles_touched": tiles}

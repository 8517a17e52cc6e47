"""`diff_gaussian_rasterization._C` — the three entry points of the reference's pybind module
(ext.cpp:15-19), same positional argument order, implemented over the C ABI (include/tetgs_rast.h)."""
from youreditableavatar_b200.rasterizer import (
    c_rasterize_gaussians as rasterize_gaussians,
    c_rasterize_gaussians_backward as rasterize_gaussians_backward,
    c_mark_visible as mark_visible,
)

__all__ = ["rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"]

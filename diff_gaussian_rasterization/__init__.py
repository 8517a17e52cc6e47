"""Drop-in `diff_gaussian_rasterization` package backed by the B200-native rasterizer.

`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer` — the import the
Edit_core scene models use (tetgs_model.py:7, tetgs_edit_2d.py:6, tetgs_edit_3d.py:5) — resolves here.
"""
from youreditableavatar_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    rasterize_gaussians,
    cpu_deep_copy_tuple,
)
from . import _C  # noqa: F401

"""youreditableavatar_b200 — B200-native (sm_100a) differentiable Gaussian rasterizer for TetGS avatars.

Only the hot path of liuhx02/YourEditableAvatar is here: the drop-in
`diff_gaussian_rasterization` / `simple_knn` operator API over a plain C-ABI CUDA library.
"""
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)
from .knn import distCUDA2  # noqa: F401

__version__ = "0.1.0"

"""Device-side camera / rasterization-settings construction for a batch of views.

The reference rebuilds the camera of every rendered view on the host, per call: cat + axis flip + `torch.inverse` +
getWorld2View + getProjectionMatrix + bmm, two `.item()` syncs and an H2D copy
(Edit_core/tetgs_scene/tetgs_model.py:467-521; the edit variants go device -> numpy -> device,
tetgs_edit_3d.py:508-518).  Here `build_cameras` turns the camera-to-world matrices of V views (the tensor the
reference indexes, `nerf_cameras.camera_to_worlds`) into the packed records the kernels read in ONE launch
(`tgr_build_cameras`, csrc/cameras.cu) with no host synchronisation, and `settings_for_views` wraps slices of that
block into the reference's `GaussianRasterizationSettings`, ready for `GaussianRasterizer` / `MultiViewRasterizer`.
"""
import math
from typing import List, Optional, Sequence, Union

import torch

from . import _lib
from ._lib import CAMERA_FLOATS, check
from .rasterizer import GaussianRasterizationSettings

__all__ = ["build_cameras", "settings_for_views", "CameraBlock"]


class CameraBlock:
    """[V, 40] fp32 on the device: viewmatrix (16, transposed W2C), projmatrix (16), campos (3), tan(fov/2) (2), pad."""

    def __init__(self, data: torch.Tensor):
        self.data = data

    def __len__(self):
        return self.data.shape[0]

    def viewmatrix(self, v: int) -> torch.Tensor:
        return self.data[v, 0:16].view(4, 4)

    def projmatrix(self, v: int) -> torch.Tensor:
        return self.data[v, 16:32].view(4, 4)

    def campos(self, v: int) -> torch.Tensor:
        return self.data[v, 32:35]

    def tanfov(self) -> torch.Tensor:
        return self.data[:, 35:37]


def build_cameras(camera_to_worlds: torch.Tensor, fovx: Union[float, torch.Tensor], fovy: Union[float, torch.Tensor],
                  cx: Union[float, torch.Tensor] = 0.0, cy: Union[float, torch.Tensor] = 0.0, znear: float = 1e-4,
                  zfar: float = 100.0) -> CameraBlock:
    """camera_to_worlds: [V,3,4] (or [V,4,4], last row ignored) CUDA fp32, nerfstudio / OpenGL axes as held by the
    reference's camera object; fovx / fovy in radians and the NDC principal point (K[0,0,2], K[0,1,2] of the
    reference's p3d camera, tetgs_model.py:498-499) as Python floats or per-view tensors.  znear / zfar defaults are
    the reference's p3d camera values."""
    if not camera_to_worlds.is_cuda:
        raise RuntimeError("build_cameras: camera_to_worlds must be a CUDA tensor (there is no CPU path)")
    if camera_to_worlds.dim() != 3 or camera_to_worlds.shape[1] not in (3, 4) or camera_to_worlds.shape[2] != 4:
        raise RuntimeError("camera_to_worlds must have dimensions (num_views, 3, 4)")
    dev = camera_to_worlds.device
    V = camera_to_worlds.shape[0]
    c2w = camera_to_worlds[:, :3, :].to(torch.float32).contiguous()
    intr = torch.empty(V, 4, dtype=torch.float32, device=dev)
    for i, x in enumerate((fovx, fovy, cx, cy)):
        intr[:, i] = x.to(device=dev, dtype=torch.float32).reshape(-1) if isinstance(x, torch.Tensor) else float(x)
    out = torch.empty(V, CAMERA_FLOATS, dtype=torch.float32, device=dev)
    if V:
        with torch.cuda.device(dev):
            check(_lib.lib().tgr_build_cameras(V, c2w.data_ptr(), intr.data_ptr(), float(znear), float(zfar),
                                               out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                  "tgr_build_cameras")
    return CameraBlock(out)


def settings_for_views(block: CameraBlock, image_height: int, image_width: int, fovx: float, fovy: float,
                       bg: torch.Tensor, sh_degree: int, scale_modifier: float = 1.0,
                       views: Optional[Sequence[int]] = None) -> List[GaussianRasterizationSettings]:
    """One `GaussianRasterizationSettings` per view, as tetgs_model.py:504-517 fills it: matrices and camera centre
    are slices of the device block (no copies), tanfov are host floats computed from the scalar fovs (the reference
    caches them on the model: tetgs_model.py `self.tanfovx`), prefiltered / debug False."""
    tanfovx, tanfovy = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    idx = range(len(block)) if views is None else views
    return [GaussianRasterizationSettings(
        image_height=int(image_height), image_width=int(image_width), tanfovx=tanfovx, tanfovy=tanfovy, bg=bg,
        scale_modifier=scale_modifier, viewmatrix=block.viewmatrix(v), projmatrix=block.projmatrix(v),
        sh_degree=sh_degree, campos=block.campos(v), prefiltered=False, debug=False) for v in idx]

"""distCUDA2: mean squared distance to the 3 nearest neighbours (simple-knn/spatial.cu:15-26)."""
import torch

from . import _lib
from ._lib import check


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Tensor[P,3] (CUDA, fp32) -> Tensor[P]; mirrors `simple_knn._C.distCUDA2`."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (no CPU fallback)")
    P = points.size(0)
    device = points.device
    means = torch.zeros(P, dtype=torch.float32, device=device)
    if P == 0:
        return means
    with torch.cuda.device(device):
        L = _lib.lib()
        pts = points.contiguous().float()
        ws = torch.empty(L.tgr_knn_bytes(P), dtype=torch.uint8, device=device)
        check(L.tgr_dist2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(), ws.numel(),
                          torch.cuda.current_stream(device).cuda_stream), "tgr_dist2")
    return means

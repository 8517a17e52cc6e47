"""Image loss of the texture / refinement stages on the B200 path.

Host-side mirror of Edit_core/utils/loss_utils.py (same names, arguments and return values: `l1_loss`, `l2_loss`,
`ssim`) plus `image_loss`, the closure the training loops build from them
(Edit_core/tetgs_texture/refine.py:241-247, paint_2dgs.py:341-347, refine_3dgs.py:273-279), all routed through the
C ABI (`tgr_image_loss_forward` / `tgr_image_loss_backward`, csrc/loss.cu): one launch for the loss of a whole batch
of views, one for its gradient, instead of ~25 ATen kernels per view each way.  No CPU path: CPU tensors raise, a
missing libtetgs_rast.so raises.
"""
from typing import Optional, Sequence, Union

import torch

from . import _lib
from ._lib import check

__all__ = ["l1_loss", "l2_loss", "ssim", "image_loss", "image_loss_and_grad"]


def _as_batch(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.size(1) != 3:
        raise RuntimeError("%s must have dimensions (3, H, W) or (V, 3, H, W)" % what)
    return t


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weights, l1_weight, l2_weight, dssim_weight):
        if not pred.is_cuda:
            raise RuntimeError("image_loss: tensors must be CUDA tensors (there is no CPU path)")
        if target.device != pred.device or target.shape != pred.shape:
            raise RuntimeError("image_loss: prediction and target must have the same shape and device")
        V, _, H, W = pred.shape
        pred = pred.contiguous().float()
        u8 = target.dtype == torch.uint8
        target = target.contiguous() if u8 else target.contiguous().float()
        L = _lib.lib()
        with torch.cuda.device(pred.device):
            nbytes = L.tgr_image_loss_bytes(V, W, H)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
            out = torch.empty(1 + V, dtype=torch.float32, device=pred.device)
            stream = torch.cuda.current_stream(pred.device).cuda_stream
            check(L.tgr_image_loss_forward(V, W, H, pred.data_ptr(), target.data_ptr(), 1 if u8 else 0,
                                           0 if weights is None else weights.data_ptr(), l1_weight, l2_weight,
                                           dssim_weight, out.data_ptr(), ws.data_ptr(), nbytes, stream),
                  "tgr_image_loss_forward")
        ctx.save_for_backward(pred, target, weights, ws)
        ctx.cfg = (V, W, H, u8, l1_weight, l2_weight, dssim_weight, nbytes)
        return out[0].clone(), out[1:].clone()

    @staticmethod
    def backward(ctx, g_total, g_per_view):
        pred, target, weights, ws = ctx.saved_tensors
        V, W, H, u8, l1_weight, l2_weight, dssim_weight, nbytes = ctx.cfg
        w = weights if weights is not None else torch.full((V,), 1.0 / V, dtype=torch.float32, device=pred.device)
        scale = (g_total.float() * w + g_per_view.float()).contiguous()      # dL / d loss_v, on the device
        grad = torch.empty_like(pred)
        L = _lib.lib()
        with torch.cuda.device(pred.device):
            stream = torch.cuda.current_stream(pred.device).cuda_stream
            check(L.tgr_image_loss_backward(V, W, H, pred.data_ptr(), target.data_ptr(), 1 if u8 else 0,
                                            scale.data_ptr(), l1_weight, l2_weight, dssim_weight, grad.data_ptr(),
                                            ws.data_ptr(), nbytes, stream), "tgr_image_loss_backward")
        return grad, None, None, None, None, None


def image_loss(pred: torch.Tensor, target: torch.Tensor, l1_weight: float = 0.8, l2_weight: float = 0.0,
               dssim_weight: float = 0.2, view_weights: Optional[Union[torch.Tensor, Sequence[float]]] = None,
               return_per_view: bool = False):
    """`(1 - dssim_factor) * l1_loss(pred, gt) + dssim_factor * (1 - ssim(pred, gt))` (refine.py:247; the defaults
    are its dssim_factor = 0.2) for a batch of views in one launch.  pred: [V,3,H,W] or [3,H,W] fp32 (the
    rasterizer's colour output); target: same shape, fp32 or uint8 (the 8-bit training image, read as u/255).
    Per view the loss is exactly the reference's; the total is sum_v w_v loss_v with w_v = 1/V (the reference's mean
    over its image batch) or `view_weights` (e.g. the x10 canonical view).  Returns the 0-dim total, or
    (total, per_view[V]) — both differentiable wrt `pred`."""
    pred, target = _as_batch(pred, "pred"), _as_batch(target, "target")
    w = None
    if view_weights is not None:
        w = torch.as_tensor(view_weights, dtype=torch.float32).to(pred.device).contiguous()
        if w.numel() != pred.size(0):
            raise RuntimeError("view_weights must hold one weight per view")
    total, per_view = _ImageLoss.apply(pred, target, w, float(l1_weight), float(l2_weight), float(dssim_weight))
    return (total, per_view) if return_per_view else total


def image_loss_and_grad(pred: torch.Tensor, target: torch.Tensor, l1_weight: float = 0.8, l2_weight: float = 0.0,
                        dssim_weight: float = 0.2, view_weights: Optional[torch.Tensor] = None,
                        workspace: Optional[torch.Tensor] = None):
    """Training-step form without autograd bookkeeping: one C-ABI call (`tgr_image_loss`: loss launch + gradient
    launch) returning (losses[1 + V] = {total, per view}, d total / d pred [V,3,H,W]) — the upstream gradient the
    rasterizer's backward consumes (`render_views_fwd_bwd(upstream=...)`).  `workspace` (uint8, >=
    tgr_image_loss_bytes) can be passed to reuse one buffer across steps; view_weights is a CUDA fp32 tensor."""
    pred, target = _as_batch(pred, "pred"), _as_batch(target, "target")
    if not pred.is_cuda:
        raise RuntimeError("image_loss: tensors must be CUDA tensors (there is no CPU path)")
    if target.device != pred.device or target.shape != pred.shape:
        raise RuntimeError("image_loss: prediction and target must have the same shape and device")
    V, _, H, W = pred.shape
    pred = pred.detach().contiguous().float()
    u8 = target.dtype == torch.uint8
    target = target.contiguous() if u8 else target.contiguous().float()
    L = _lib.lib()
    with torch.cuda.device(pred.device):
        nbytes = L.tgr_image_loss_bytes(V, W, H)
        if workspace is None or workspace.numel() < nbytes:
            workspace = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
        out = torch.empty(1 + V, dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        check(L.tgr_image_loss(V, W, H, pred.data_ptr(), target.data_ptr(), 1 if u8 else 0,
                               0 if view_weights is None else view_weights.data_ptr(), float(l1_weight),
                               float(l2_weight), float(dssim_weight), out.data_ptr(), grad.data_ptr(),
                               workspace.data_ptr(), workspace.numel(),
                               torch.cuda.current_stream(pred.device).cuda_stream), "tgr_image_loss")
    return out, grad


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:17-18 — mean |network_output - gt| over every element."""
    return image_loss(network_output, gt, 1.0, 0.0, 0.0)


def l2_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:20-21"""
    return image_loss(network_output, gt, 0.0, 1.0, 0.0)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """loss_utils.py:33-63 — mean SSIM (size_average) or one mean per image of the batch."""
    if window_size != 11:
        raise RuntimeError("ssim: only the reference's window_size = 11 is built (loss_utils.py:33)")
    total, per_view = image_loss(img1, img2, 0.0, 0.0, 1.0, return_per_view=True)   # dssim term = 1 - mean ssim
    return 1.0 - total if size_average else 1.0 - per_view

"""CPU float64 oracle of the steps either side of the rasterizer.  TEST INFRASTRUCTURE ONLY.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs may import this module;
the product path (youreditableavatar_b200/) never does.

Restated here (paths relative to /root/reference/Edit_core/):
  image loss   utils/loss_utils.py:17-63 (l1_loss, l2_loss, gaussian, create_window, ssim, _ssim) and the closure
               at tetgs_texture/refine.py:245-247; gradients by torch.autograd on this forward
  Adam         tetgs_scene/tetgs_optimizer.py:66-108 drives torch.optim.Adam(l, lr=0.0, eps=1e-15); the arithmetic
               lives in a third-party dependency, torch (requirements.txt: unpinned; torch 2.11.0 here) —
               restated from its published algorithm (torch/optim/adam.py `_single_tensor_adam`), anchored on the
               call site above and pinned against torch.optim.Adam itself (tests/golden/train_adam.npz)
  lr schedule  utils/general_utils.py:25-58 (get_expon_lr_func)
  cameras      tetgs_scene/tetgs_model.py:479-503, utils/graphics_utils.py:39-49,68-86

Pinning: the reference ships no tests for these either; tests/golden/train_*.npz hold outputs of the reference's own
Python functions imported from /root/reference on CPU (oracle/make_golden_train.py, committed), and
tests/test_train_oracle.py checks this module against them.
"""
import math
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

F64 = torch.float64


# ---------------------------------------------------------------------------------------- image loss
def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """loss_utils.py:23-25 — NOTE the weights are float32 values normalised in float32; kept so in the oracle
    (they are data of the algorithm), everything after is float64."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return (g / g.sum()).to(F64)


def ssim_map(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """loss_utils.py:44-59 on [B,C,H,W] (float64 unless the caller asks otherwise); the 2-D window is the float32
    outer product (loss_utils.py:27-31)."""
    C = img1.size(-3)
    w1 = gaussian_window(window_size).to(torch.float32).unsqueeze(1)
    w2 = w1.mm(w1.t()).to(device=img1.device, dtype=img1.dtype)
    window = w2.unsqueeze(0).unsqueeze(0).expand(C, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=C)
    mu2 = F.conv2d(img2, window, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, window, padding=pad, groups=C) - mu1_sq
    s2 = F.conv2d(img2 * img2, window, padding=pad, groups=C) - mu2_sq
    s12 = F.conv2d(img1 * img2, window, padding=pad, groups=C) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def image_loss(pred: torch.Tensor, target: torch.Tensor, l1_weight: float = 0.8, l2_weight: float = 0.0,
               dssim_weight: float = 0.2, view_weights: Optional[Sequence[float]] = None, dtype=F64):
    """pred, target: [V,3,H,W].  Returns (total, per_view[V]) in float64 with autograd attached to `pred`.
    Per view: refine.py:247 with l1_loss / l2_loss / ssim of loss_utils.py:17-21,33-42 applied to that view alone
    (the reference renders one view per step); total = sum_v w_v loss_v, w_v = 1/V by default.
    dtype=torch.float32 on CUDA tensors is the reference's own precision and op sequence (five grouped conv2d +
    elementwise ATen kernels) — what bench.py's `--impl reference` arm times."""
    p, g = pred.to(dtype), target.to(dtype)
    V = p.shape[0]
    d = p - g
    per = l1_weight * d.abs().mean(dim=(1, 2, 3)) + l2_weight * (d * d).mean(dim=(1, 2, 3))
    if dssim_weight != 0.0:
        per = per + dssim_weight * (1.0 - ssim_map(p, g).mean(dim=(1, 2, 3)))
    w = torch.full((V,), 1.0 / V, dtype=dtype, device=p.device) if view_weights is None else \
        torch.as_tensor(view_weights, dtype=dtype).to(p.device)
    return (w * per).sum(), per


# ---------------------------------------------------------------------------------------- Adam
def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-15):
    """One torch.optim.Adam step (no weight decay / amsgrad) in float64 numpy; `lr` scalar or per-element array.
    torch/optim/adam.py `_single_tensor_adam`: lerp, mul+addcmul, bias corrections, addcdiv."""
    p, g, m, v = (np.asarray(a, dtype=np.float64) for a in (p, g, m, v))
    m = m + (g - m) * (1.0 - beta1)
    v = v * beta2 + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (np.asarray(lr, dtype=np.float64) / bc1) * (m / denom)
    return p, m, v


def expon_lr(step: int, lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
             max_steps: int = 1000000) -> float:
    """general_utils.py:25-58"""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0), 1)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


# ---------------------------------------------------------------------------------------- cameras
def build_camera(c2w_3x4: np.ndarray, fovx: float, fovy: float, cx: float = 0.0, cy: float = 0.0, znear: float = 1e-4,
                 zfar: float = 100.0):
    """tetgs_model.py:479-503 in float64 numpy.  Returns (viewmatrix[4,4], projmatrix[4,4], campos[3], tanfovx,
    tanfovy) with viewmatrix = W2C^T and projmatrix = W2C^T P^T (the transposed matrices the kernels read)."""
    c2w = np.concatenate([np.asarray(c2w_3x4, dtype=np.float64), np.array([[0.0, 0.0, 0.0, 1.0]])], axis=0)
    c2w[:3, 1:3] *= -1                                   # :485
    w2c = np.linalg.inv(c2w)                             # :488
    R = w2c[:3, :3].T                                    # :489
    T = w2c[:3, 3]
    Rt = np.zeros((4, 4))                                # graphics_utils.py:39-45
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    view = Rt.T                                          # :491-492
    thy, thx = math.tan(fovy / 2), math.tan(fovx / 2)    # graphics_utils.py:68-86
    top, right = thy * znear, thx * znear
    P = np.zeros((4, 4))
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.T.copy()                                    # :493-497
    proj[2, 0] = -cx                                     # :498-499
    proj[2, 1] = -cy
    full = view @ proj                                   # :501
    campos = c2w[:3, 3].copy()                           # :502 camera centre
    return view, full, campos, thx, thy

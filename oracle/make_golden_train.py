"""Generates tests/golden/train_*.npz by running the reference's OWN Python functions (imported from /root/reference,
CPU, fp32) for the steps either side of the rasterizer:

  train_loss.npz     Edit_core/utils/loss_utils.py l1_loss / l2_loss / ssim and the refine.py:247 closure, with the
                     gradient of the loss wrt the prediction from torch.autograd through the reference code
  train_adam.npz     torch.optim.Adam(l, lr=0.0, eps=1e-15) driven exactly as TetGSOptimizer does
                     (Edit_core/tetgs_scene/tetgs_optimizer.py:66-117), position lr from the reference's
                     get_expon_lr_func (utils/general_utils.py:25-58); three steps
  train_binding.npz  utils/graphics_utils.py triangle_area / calculate_distances (the binding rule's two primitives)
  train_cameras.npz  the camera preamble of tetgs_model.py:479-503 executed line by line with the reference's
                     getWorld2View / getProjectionMatrix (utils/graphics_utils.py)

Run in the build container (needs /root/reference; not on the GPU box):  python oracle/make_golden_train.py
pytorch3d / open3d are absent here and are only imported (never called) by the two utils modules: they are stubbed.
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Edit_core/utils"
OUT = os.path.join(ROOT, "tests", "golden")


def _ref_module(name):
    for stub in ("pytorch3d", "pytorch3d.transforms", "open3d"):
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.matrix_to_quaternion = None
            sys.modules[stub] = m
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def smooth_images(V, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(V, 3, H, W, generator=g)
    k = torch.ones(3, 1, 5, 5) / 25.0
    gt = torch.nn.functional.conv2d(x, k, padding=2, groups=3).clamp(0, 1)
    gt = (gt - gt.amin()) / (gt.amax() - gt.amin())
    pred = (gt + 0.08 * torch.randn(V, 3, H, W, generator=g) + 0.05).clamp(0, 1.2)
    return pred.contiguous(), gt.contiguous()


def make_loss():
    lu = _ref_module("loss_utils")
    d = {}
    cases = [("a", 2, 45, 77, 3), ("b", 1, 64, 64, 4), ("c", 3, 33, 100, 5), ("d", 1, 9, 7, 6)]
    d["cases"] = np.array([c[0] for c in cases])
    for name, V, H, W, seed in cases:
        pred, gt = smooth_images(V, H, W, seed)
        gt_u8 = (gt * 255).round().to(torch.uint8)
        gt = gt_u8.float() / 255.0                    # general_utils.py:8 — training targets are 8-bit images
        d[name + "_pred"], d[name + "_gt_u8"] = pred.numpy(), gt_u8.numpy()
        per_l1, per_l2, per_ss, per_loss, grads = [], [], [], [], []
        for v in range(V):                            # the reference evaluates one view per call (refine.py:54)
            p = pred[v:v + 1].clone().requires_grad_(True)
            l1, l2, ss = lu.l1_loss(p, gt[v:v + 1]), lu.l2_loss(p, gt[v:v + 1]), lu.ssim(p, gt[v:v + 1])
            loss = (1.0 - 0.2) * l1 + 0.2 * (1.0 - ss)   # refine.py:247, dssim_factor 0.2 (refine.py:57)
            loss.backward()
            per_l1.append(l1.item()); per_l2.append(l2.item()); per_ss.append(ss.item()); per_loss.append(loss.item())
            grads.append(p.grad[0].numpy())
        d[name + "_l1"], d[name + "_l2"] = np.array(per_l1), np.array(per_l2)
        d[name + "_ssim"], d[name + "_loss"] = np.array(per_ss), np.array(per_loss)
        d[name + "_grad"] = np.stack(grads)           # d loss_v / d pred_v per view (not yet weighted by 1/V)
        # the batched call of the reference (mean over the batch) for the total
        d[name + "_ssim_batched"] = np.float64(lu.ssim(pred, gt).item())
    np.savez_compressed(os.path.join(OUT, "train_loss.npz"), **d)
    print("train_loss.npz", {k: d[k] for k in d if k.endswith("_loss")})


def make_adam():
    gu = _ref_module("general_utils")
    P, M = 37, 16
    g = torch.Generator().manual_seed(11)
    shapes = {"points": (P, 3), "sh_dc": (P, 1, 3), "sh_rest": (P, M - 1, 3), "densities": (P, 1), "scales": (P, 3),
              "quaternions": (P, 4)}
    # OptimizationParams defaults (tetgs_optimizer.py:8-31) and spatial_lr_scale
    opt = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
               position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
    spatial = 2.5
    params = {k: torch.randn(*s, generator=g).requires_grad_(True) for k, s in shapes.items()}
    d = {"spatial_lr_scale": np.float64(spatial)}
    d.update({"opt_" + k: np.float64(v) for k, v in opt.items()})
    for k, p in params.items():
        d["p0_" + k] = p.detach().numpy().copy()
    l = [{"params": [params["points"]], "lr": opt["position_lr_init"] * spatial, "name": "points"},
         {"params": [params["sh_dc"]], "lr": opt["feature_lr"], "name": "sh_coordinates_dc"},
         {"params": [params["sh_rest"]], "lr": opt["feature_lr"] / 20.0, "name": "sh_coordinates_rest"},
         {"params": [params["densities"]], "lr": opt["opacity_lr"], "name": "all_densities"},
         {"params": [params["scales"]], "lr": opt["scaling_lr"], "name": "scales"},
         {"params": [params["quaternions"]], "lr": opt["rotation_lr"], "name": "quaternions"}]
    optim = torch.optim.Adam(l, lr=0.0, eps=1e-15)                                   # tetgs_optimizer.py:92
    sched = gu.get_expon_lr_func(lr_init=opt["position_lr_init"] * spatial, lr_final=opt["position_lr_final"] * spatial,
                                 lr_delay_mult=opt["position_lr_delay_mult"], max_steps=opt["position_lr_max_steps"])
    iters = [1, 2, 700]          # update_learning_rate(iteration) is called with the loop's iteration (refine.py:268)
    d["iterations"] = np.array(iters)
    lrs = []
    for s, it in enumerate(iters):
        for grp in optim.param_groups:                                               # tetgs_optimizer.py:110-117
            if grp["name"] == "points":
                grp["lr"] = sched(it)
        lrs.append(float(optim.param_groups[0]["lr"]))
        for k, p in params.items():
            gr = torch.randn(*shapes[k], generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
            if s == 1 and k == "densities":
                gr.zero_()                                                           # an all-zero gradient step
            p.grad = gr
            d["g%d_%s" % (s, k)] = gr.numpy().copy()
        optim.step()
        for k, p in params.items():
            st = optim.state[p]
            d["p%d_%s" % (s + 1, k)] = p.detach().numpy().copy()
            d["m%d_%s" % (s + 1, k)] = st["exp_avg"].numpy().copy()
            d["v%d_%s" % (s + 1, k)] = st["exp_avg_sq"].numpy().copy()
    d["points_lr"] = np.array(lrs)
    np.savez_compressed(os.path.join(OUT, "train_adam.npz"), **d)
    print("train_adam.npz points lr", lrs)


def make_cameras():
    gr = _ref_module("graphics_utils")
    V = 6
    g = torch.Generator().manual_seed(21)
    c2ws, views, fulls, centres = [], [], [], []
    fovx, fovy = 2 * math.atan(0.5 / 1.4), 2 * math.atan(0.45 / 1.4)
    cxcy = torch.tensor([[0.0, 0.0], [0.0, 0.0], [0.01, -0.02], [0.0, 0.0], [-0.03, 0.005], [0.0, 0.0]])
    znear, zfar = 1e-4, 100.0
    for v in range(V):
        # orbit camera looking at the origin (cameras.py:281-345 style), OpenGL axes, as a nerfstudio c2w [3,4]
        az, el, r = 2 * math.pi * v / V + 0.3, math.radians([5, -15, 25][v % 3]), 3.0
        eye = torch.tensor([r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el)])
        back = eye / eye.norm()
        right = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0]), back)
        right = right / right.norm()
        up = torch.linalg.cross(back, right)
        c2w34 = torch.stack([right, up, back, eye], dim=1) + 1e-3 * torch.randn(3, 4, generator=g)
        # ---- tetgs_model.py:479-503, line by line ----
        c2w = torch.cat([c2w34, torch.Tensor([[0, 0, 0, 1]])], dim=0)
        c2w[:3, 1:3] *= -1
        c2w = c2w.squeeze()
        w2c = torch.inverse(c2w)
        R = w2c[:3, :3].T
        T = w2c[:3, 3]
        world_view_transform = torch.Tensor(gr.getWorld2View(R=R, t=T, tensor=True)).transpose(0, 1)
        proj_transform = gr.getProjectionMatrix(znear, zfar, fovx, fovy).transpose(0, 1)
        proj_transform[..., 2, 0] = -cxcy[v, 0]
        proj_transform[..., 2, 1] = -cxcy[v, 1]
        full = (world_view_transform.unsqueeze(0).bmm(proj_transform.unsqueeze(0))).squeeze(0)
        c2ws.append(c2w34.numpy()); views.append(world_view_transform.contiguous().numpy())
        fulls.append(full.numpy()); centres.append(c2w[:3, 3].numpy())   # p3d get_camera_center == c2w translation
    np.savez_compressed(os.path.join(OUT, "train_cameras.npz"), c2w=np.stack(c2ws), fovx=np.float64(fovx),
                        fovy=np.float64(fovy), cxcy=cxcy.numpy(), znear=np.float64(znear), zfar=np.float64(zfar),
                        viewmatrix=np.stack(views), projmatrix=np.stack(fulls), campos=np.stack(centres),
                        tanfovx=np.float64(math.tan(fovx / 2)), tanfovy=np.float64(math.tan(fovy / 2)))
    print("train_cameras.npz", V, "views")


def make_binding():
    """utils/graphics_utils.py triangle_area / calculate_distances on a random triangle soup."""
    import contextlib
    import io
    gr = _ref_module("graphics_utils")
    g = torch.Generator().manual_seed(31)
    A, B, C = (torch.randn(200, 3, generator=g) for _ in range(3))
    B[:20] = A[:20] + 1e-4 * torch.randn(20, 3, generator=g)            # slivers
    pts = (A + B + C) / 3 + 0.01 * torch.randn(200, 3, generator=g)
    with contextlib.redirect_stdout(io.StringIO()):                      # calculate_distances prints a shape
        dist = gr.calculate_distances(pts, A, B, C)
    np.savez_compressed(os.path.join(OUT, "train_binding.npz"), A=A.numpy(), B=B.numpy(), C=C.numpy(), points=pts.numpy(),
                        area=gr.triangle_area(A, B, C).numpy(), distances=dist.numpy())
    print("train_binding.npz")


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_loss()
    make_adam()
    make_cameras()
    make_binding()

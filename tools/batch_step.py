"""Runs a few multi-view batches (fwd+bwd, fused schedule) of one config — the command ncu wraps."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import scene  # noqa: E402
from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--streams", type=int, default=0)
ap.add_argument("--train", action="store_true",
                help="whole training step: fused image loss against uint8 targets as the upstream gradient (no extras), "
                     "one-launch Adam after the backward (SURVEY §8 f1-f2)")
args = ap.parse_args()
P, res, _, g = scene.CONFIGS[args.config]
V = args.views
gs = scene.make_scene(args.config, device="cuda")
act = scene.activate(gs)
cams = [scene.orbit_camera(v, V, res, res, device="cuda") for v in range(V)]
gen = torch.Generator().manual_seed(1)
N = res * res
ups = ((torch.randn(V, 3, res, res, generator=gen) / (3 * N)).cuda(), (torch.randn(V, 1, res, res, generator=gen) / N).cuda(),
       (torch.randn(V, 1, res, res, generator=gen) / N).cuda())
bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
torch.cuda.synchronize()
if args.train:
    from youreditableavatar_b200 import loss_utils
    from youreditableavatar_b200.optimizer import OptimizationParams, TetGSOptimizer
    params = {k: act[k].clone() for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    gv = bucket.named()
    o = OptimizationParams()
    opt = TetGSOptimizer({"points": params["means3D"], "sh": params["shs"], "all_densities": params["opacities"],
                          "scales": params["scales"], "quaternions": params["rotations"]},
                         OptimizationParams(position_lr_init=o.position_lr_init * 1e-3, position_lr_final=o.position_lr_final * 1e-3,
                                            feature_lr=o.feature_lr * 1e-3, opacity_lr=o.opacity_lr * 1e-3,
                                            scaling_lr=o.scaling_lr * 1e-3, rotation_lr=o.rotation_lr * 1e-3), 1.0,
                         grads={"points": gv["dL_dmeans3D"], "sh": gv["dL_dsh"], "all_densities": gv["dL_dopacity"],
                                "scales": gv["dL_dscales"], "quaternions": gv["dL_drotations"]})
    targets = torch.randint(0, 256, (V, 3, res, res), generator=gen, dtype=torch.uint8).cuda()
    for _ in range(args.steps):
        render_views_fwd_bwd(params, cams, 3, lambda c, d, a: (loss_utils.image_loss_and_grad(c, targets)[1], None, None),
                             bucket, extras=False, n_streams=args.streams)
        opt.step()
else:
    for _ in range(args.steps):
        render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, bucket, extras=True, n_streams=args.streams)
torch.cuda.synchronize()
print("done")

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
for _ in range(args.steps):
    render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, bucket, extras=True, n_streams=args.streams)
torch.cuda.synchronize()
print("done")

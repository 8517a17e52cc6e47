"""Runs a few fwd+bwd steps of one implementation on one config — the command ncu wraps."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from youreditableavatar_b200 import scene  # noqa: E402
from helpers import ours_forward, ours_backward  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--impl", default="ours")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--view", type=int, default=0)
ap.add_argument("--extras", type=int, default=0)
args = ap.parse_args()
P, res, _, g = scene.CONFIGS[args.config]
gs = scene.make_scene(args.config, device="cuda")
act = scene.activate(gs)
cam = scene.orbit_camera(args.view, 8, res, res, device="cuda")
gen = torch.Generator().manual_seed(1)
dL = (torch.randn(3, res, res, generator=gen) / (3 * res * res)).cuda()
torch.cuda.synchronize()
if args.impl == "ours":
    for _ in range(args.steps):
        fo = ours_forward(act, cam, 3, extras=bool(args.extras))
        ours_backward(act, cam, 3, fo, dL)
else:
    from oracle import ref_cuda
    for _ in range(args.steps):
        fr = ref_cuda.forward(act, cam, 3)
        ref_cuda.backward(act, cam, 3, fr, dL)
torch.cuda.synchronize()
print("done")

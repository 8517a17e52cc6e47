"""How even is the work of the 8 ranks of the 64-view C4 batch?  Renders every rank's 8 views (view v -> rank v mod 8) on ONE
GPU, one subset after the other: instance counts and fwd+bwd time per subset."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import scene  # noqa: E402
from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd, shard_views  # noqa: E402

cfg, V, G = "C3", 8, 8
P, res, _, _ = scene.CONFIGS[cfg]
act = scene.activate(scene.make_scene(cfg, device="cuda"))
N = res * res
gen = torch.Generator().manual_seed(1)
ups = ((torch.randn(V, 3, res, res, generator=gen) / (3 * N)).cuda(), (torch.randn(V, 1, res, res, generator=gen) / N).cuda(),
       (torch.randn(V, 1, res, res, generator=gen) / N).cuda())
bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
from youreditableavatar_b200 import multiview as mv
from youreditableavatar_b200.parallel import settings_from_cam
for policy in ("v mod G", "contiguous"):
    print("policy:", policy)
    for r in range(G):
        ids = shard_views(V * G, r, G) if policy == "v mod G" else list(range(r * V, (r + 1) * V))
        cams = [scene.orbit_camera(v, V * G, res, res, device="cuda") for v in ids]
        for _ in range(3):
            render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, bucket, extras=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, bucket, extras=True)
        e1.record()
        torch.cuda.synchronize()
        e = torch.Tensor([])
        st = mv.c_rasterize_views([settings_from_cam(c, 3) for c in cams], act["means3D"], e, act["opacities"], act["scales"],
                                  act["rotations"], e, act["shs"], extras=True)[0]
        print("  rank %d: %.3f ms/step, sum R = %.2f M, views %s" % (r, e0.elapsed_time(e1) / 5, sum(st.counts) / 1e6, ids))

"""cProfile of the host side of the e2e step at C1 (10 k Gaussians, 256^2, one view): where does the Python time go?"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    impl, per_view_api, streams, comm, comm_chunks = "ours", False, 0, "auto", 2


torch.cuda.set_device(0)
P, res, act, cams, up_host, up_dev, targets_host = bench.build_workload("C1", 1, 0, 1)
runner = bench.OursRunner(P, res, act, n_streams=0, comm="nccl")
feeder = bench.BatchFeeder([bench.cam_to_host(c) for c in cams], targets_host)
ups = tuple(torch.stack([u[k] for u in up_host]).cuda() for k in range(3))
for _ in range(20):
    runner.step(cams, ups, 1, feeder)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    runner.step(cams, ups, 1, feeder)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)

"""Experiment: batched multi-view path (one preprocess / preprocess_bwd launch per batch, S streams in between)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import scene, _lib
from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd
import ctypes as C

V = int(os.environ.get("V", 8))
cfg = "C3"
P, res, _, g = scene.CONFIGS[cfg]
gs = scene.make_scene(cfg, device="cuda"); act = scene.activate(gs)
cams = [scene.orbit_camera(v, V, res, res, device="cuda") for v in range(V)]
gen = torch.Generator().manual_seed(1)
N = res * res
dLc = (torch.randn(V, 3, res, res, generator=gen) / (3 * N)).cuda()
dLd = (torch.randn(V, 1, res, res, generator=gen) / N).cuda()
dLa = (torch.randn(V, 1, res, res, generator=gen) / N).cuda()
bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
L = _lib.lib()
for S in [int(x) for x in sys.argv[1:]] or [1, 2, 4]:
    step = lambda: render_views_fwd_bwd(act, cams, 3, lambda c, d, a: (dLc, dLd, dLa), bucket, extras=True, n_streams=S)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    line = "batched V=%d streams %d: %.3f ms/step  %.1f views/s" % (V, S, ms / K, V * K / ms * 1000)
    if S <= 1:
        L.tgr_profile_enable(1)
        step(); torch.cuda.synchronize()
        sums = (C.c_float * _lib.NUM_STAGES)(); cnts = (C.c_int32 * _lib.NUM_STAGES)()
        L.tgr_profile_collect(sums, cnts); L.tgr_profile_enable(0)
        line += "  stages(ms per launch): " + " ".join("%s=%.3f/%d" % (n, sums[i] / max(cnts[i], 1), cnts[i]) for i, n in enumerate(_lib.STAGE_NAMES))
    print(line)

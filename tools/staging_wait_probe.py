"""Experiment for profiles/r02_staging_wait.txt: with the library built with TGR_NVCC_DEFINES=-DTGR_MEASURE_STAGING, how
much of the forward blend's consumer time is spent waiting for staged records (= the most a bulk-copy / TMA producer
could win)?   TGR_NVCC_DEFINES=-DTGR_MEASURE_STAGING python -m youreditableavatar_b200.build -f && python tools/staging_wait_probe.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import _lib, scene  # noqa: E402
from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd  # noqa: E402

L = _lib.lib()
fn = L.tgr_debug_staging_cycles
fn.argtypes = [C.POINTER(C.c_ulonglong * 4), C.c_int]
for cfg, V in (("C3", 8), ("C5", 2), ("C1", 1)):
    P, res, _, _ = scene.CONFIGS[cfg]
    act = scene.activate(scene.make_scene(cfg, device="cuda"))
    cams = [scene.orbit_camera(v, V, res, res, device="cuda") for v in range(V)]
    ups = (torch.zeros(V, 3, res, res, device="cuda"), None, None)
    bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
    for it in range(3):
        out = (C.c_ulonglong * 4)()
        fn(out, 1)
        render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, bucket, extras=False)
    out = (C.c_ulonglong * 4)()
    fn(out, 1)
    w, t, pw, pt = [int(x) for x in out]
    print("%s, %d views: consumer warps blocked on full[] %.1f %% of their cycles; producer warps blocked on empty[] (data "
          "ready, ring full: back-pressure of the slowest consumer) %.1f %% of theirs" % (cfg, V, 100.0 * w / max(t, 1), 100.0 * pw / max(pt, 1)))
    scene._scene_cache.clear()
    del act, bucket
    torch.cuda.empty_cache()

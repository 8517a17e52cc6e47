"""Development probe: stage timings of ours vs the reference CUDA build on one config (run under gpurun)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from youreditableavatar_b200 import scene  # noqa: E402
from helpers import ours_forward, ours_backward, to_dev, rel_l2  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--views", type=int, default=3)
    ap.add_argument("--ref", type=int, default=1)
    args = ap.parse_args()
    P, res, _, g = scene.CONFIGS[args.config]
    t0 = time.time()
    gs = scene.make_scene(args.config, device="cuda")
    act = scene.activate(gs)
    torch.cuda.synchronize()
    print("scene %s: P=%d g=%d faces=%d gen %.1fs" % (args.config, P, g, gs["faces"].shape[0], time.time() - t0))
    ref = None
    if args.ref:
        from oracle import ref_cuda
        if ref_cuda.available():
            ref = ref_cuda
    for v in range(args.views):
        cam = scene.orbit_camera(v, 8, res, res, device="cuda")
        gen = torch.Generator().manual_seed(1)
        dL = (torch.randn(3, res, res, generator=gen) / (3 * res * res)).cuda()
        fo = ours_forward(act, cam, 3)
        go = ours_backward(act, cam, 3, fo, dL)
        torch.cuda.synchronize()
        R = fo[0]
        vis = int((fo[2] > 0).sum())
        print("view %d: R=%d visible=%d tiles/gauss=%.2f" % (v, R, vis, R / max(vis, 1)))
        t_f = timeit(lambda: ours_forward(act, cam, 3))
        t_fb = timeit(lambda: ours_backward(act, cam, 3, ours_forward(act, cam, 3), dL))
        print("  ours: fwd %.3f ms  fwd+bwd %.3f ms  (%.1f views/s)" % (t_f, t_fb, 1000.0 / t_fb))
        if ref is not None:
            fr = ref.forward(act, cam, 3)
            gr = ref.backward(act, cam, 3, fr, dL)
            torch.cuda.synchronize()
            print("  parity: R %s img max-abs %.3g radii-eq %s" % (fr[0] == R, float((fr[1] - fo[1]).abs().max()),
                                                                 bool(torch.equal(fr[2], fo[2]))))
            for n, a, b in zip(["m2D", "col", "op", "m3D", "cov", "sh", "sc", "rot"], go, gr):
                print("    grad %-4s rel-L2 %.3g" % (n, rel_l2(a, b)))
            r_f = timeit(lambda: ref.forward(act, cam, 3))
            r_fb = timeit(lambda: ref.backward(act, cam, 3, ref.forward(act, cam, 3), dL))
            print("  ref : fwd %.3f ms  fwd+bwd %.3f ms  (%.1f views/s)  speedup %.2fx" % (r_f, r_fb, 1000.0 / r_fb, r_fb / t_fb))


if __name__ == "__main__":
    main()

"""Launches the kernels of the rows either side of the rasterizer at C3 / C4 size a few times:
image loss forward + gradient for 8 views of 1024^2 with uint8 targets, one Adam step over the C3 parameter set
(59 M floats, SH rows with dc / rest rates), camera construction for 64 views.  Run it under ncu for profiles/:

  ncu --set full --clock-control none --import-source on -k regex:"loss_|adam_|build_cameras" -c 8 \
      -o gpurun_out/prof_train python tools/train_kernels_probe.py

Without a profiler it prints CUDA-event timings (never taken under ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from youreditableavatar_b200 import _lib, cameras, loss_utils  # noqa: E402
from youreditableavatar_b200.optimizer import OptimizationParams, TetGSOptimizer  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    V, H, W, P, M = 8, 1024, 1024, 1_000_000, 16
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.rand(V, 3, H, W, device="cuda", generator=g)
    tgt = (torch.rand(V, 3, H, W, device="cuda", generator=g) * 255).to(torch.uint8)
    ws = torch.empty(L.tgr_image_loss_bytes(V, W, H), dtype=torch.uint8, device="cuda")
    out, grad = torch.empty(1 + V, device="cuda"), torch.empty_like(pred)
    st = torch.cuda.current_stream().cuda_stream
    a = (V, W, H, pred.data_ptr(), tgt.data_ptr(), 1, 0, 0.8, 0.0, 0.2)
    fwd = lambda: L.tgr_image_loss_forward(*a, out.data_ptr(), ws.data_ptr(), ws.numel(), st)
    bwd = lambda: L.tgr_image_loss_backward(*a, grad.data_ptr(), ws.data_ptr(), ws.numel(), st)
    params = {"points": torch.randn(P, 3, device="cuda"), "sh": torch.randn(P, M, 3, device="cuda"),
              "all_densities": torch.randn(P, 1, device="cuda"), "scales": torch.randn(P, 3, device="cuda"),
              "quaternions": torch.randn(P, 4, device="cuda")}
    grads = {k: torch.randn_like(v) * 1e-3 for k, v in params.items()}
    opt = TetGSOptimizer(params, OptimizationParams(), 1.0, grads=grads)
    c2w = torch.eye(4, device="cuda")[:3].repeat(64, 1, 1) + 0.01 * torch.randn(64, 3, 4, device="cuda")
    cam = lambda: cameras.build_cameras(c2w, 0.7, 0.7)
    if reps > 2:   # plain timing run
        n_px = V * 3 * H * W
        n_par = sum(p.numel() for p in params.values())
        for name, fn, nbytes in (("loss_fwd", fwd, n_px * 17), ("loss_bwd", bwd, n_px * 21), ("adam", opt.step, n_par * 28),
                                 ("build_cameras(64)", cam, 64 * (48 + 16 + 160))):
            ms = timed(fn, reps)
            print("%-20s %.4f ms  %.1f GB/s algorithmic" % (name, ms, nbytes / ms / 1e6))
    else:
        for _ in range(reps):
            fwd(); bwd(); opt.step(); cam()
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()

"""Experiment: views round-robin over S CUDA streams (tail filling between independent views)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import scene, rasterizer as rz
from youreditableavatar_b200.parallel import GradBucket

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
V = 8
cfg = "C3"
P, res, _, g = scene.CONFIGS[cfg]
gs = scene.make_scene(cfg, device="cuda"); act = scene.activate(gs)
cams = [scene.orbit_camera(v, V, res, res, device="cuda") for v in range(V)]
gen = torch.Generator().manual_seed(1)
N = res * res
ups = [((torch.randn(3, res, res, generator=gen) / (3 * N)).cuda(), (torch.randn(1, res, res, generator=gen) / N).cuda(),
        (torch.randn(1, res, res, generator=gen) / N).cuda()) for _ in range(V)]
buckets = [GradBucket(P, 16, "cuda", names=GradBucket.TRAINING) for _ in range(S)]
streams = [torch.cuda.Stream() for _ in range(S)]
e = torch.Tensor([])

def step():
    main = torch.cuda.current_stream()
    ev0 = torch.cuda.Event(); ev0.record(main)
    keep = []
    for i in range(V):
        s = i % S
        st = streams[s]
        if i < S:
            st.wait_event(ev0)
        with torch.cuda.stream(st):
            cam, up = cams[i], ups[i]
            fwd = rz.c_rasterize_gaussians(cam["bg"], act["means3D"], e, act["opacities"], act["scales"], act["rotations"],
                                           1.0, e, cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
                                           res, res, act["shs"], 3, cam["campos"], False, False, extras=True)
            R, color, radii, geom, binning, img = fwd[:6]
            kw = dict(accumulate_into=buckets[s].views) if i >= S else dict(out=buckets[s].views)
            rz.c_rasterize_gaussians_backward(cam["bg"], act["means3D"], radii, e, act["scales"], act["rotations"], 1.0, e,
                                              cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], up[0],
                                              act["shs"], 3, cam["campos"], geom, R, binning, img, False,
                                              dL_dout_depth=up[1], dL_dout_alpha=up[2], **kw)
            keep.append(fwd)
    for st in streams:
        main.wait_stream(st)
    for b in buckets[1:]:
        buckets[0].flat.add_(b.flat)
    return keep

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
print("streams %d: %.3f ms/step  %.1f views/s" % (S, ms / K, V * K / ms * 1000))

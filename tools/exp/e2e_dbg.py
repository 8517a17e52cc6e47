"""Debug the e2e leg: per-step wall/GPU times under feeder variants."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
args = type("A", (), dict(config="C3"))
V, world, rank = 8, 1, 0
torch.cuda.set_device(0)
P, res, act, cams, up_host, up_dev = bench.build_workload("C3", V, 0, 1)
up_stack_host = tuple(torch.stack([u[k] for u in up_host]).pin_memory() for k in range(3))
up_stack_dev = tuple(t.cuda() for t in up_stack_host)
runner = bench.OursRunner(P, res, act, n_streams=4)
host_cams = [bench.cam_to_host(c) for c in cams]
out_pinned = torch.empty(V, 3, res, res).pin_memory()
for mode in sys.argv[1:]:
    feeder = bench.BatchFeeder(host_cams, up_stack_host, out_pinned)
    if "nod2h" in mode:
        feeder.images_out = lambda color, ev: feeder.keep.append(color)
    if "noh2d" in mode:
        fixed = feeder._issue()
        feeder._issue = lambda: fixed
    if "syncloss" in mode:
        orig = feeder.end_step
        def end_step(result, orig=orig):
            orig(result); torch.cuda.synchronize(); return 0.0
        feeder.end_step = end_step
    for _ in range(3):
        runner.step(cams, up_stack_dev, 1, feeder)
    torch.cuda.synchronize()
    t0 = time.time(); ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        runner.step(cams, up_stack_dev, 1, feeder); ts.append(time.time() - t0)
    feeder.drain(); e1.record(); torch.cuda.synchronize()
    print(mode, "gpu ms/step %.3f" % (e0.elapsed_time(e1) / 10), "host issue times(ms):", " ".join("%.1f" % (1e3 * (b - a)) for a, b in zip([0] + ts[:-1], ts)))

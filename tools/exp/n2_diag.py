"""torchrun diag: per-rank compute time without the all-reduce, all-reduce alone, and both (overlapped / not)."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
V = 8
P, res, act, cams, up_host, up_dev, _tg = bench.build_workload("C3", V, rank, world)
ups = tuple(torch.stack([u[k] for u in up_host]).cuda() for k in range(3))
from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd
bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)

def run(name, fn, K=15):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device="cuda"); g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    if rank == 0: print(name, " ".join("%.3f" % x.item() for x in g), "ms/step per rank", flush=True)

up = lambda c, d, a: ups
run("compute only      ", lambda: render_views_fwd_bwd(act, cams, 3, up, bucket, extras=True, n_streams=0))
run("allreduce only    ", lambda: bucket.all_reduce())
run("compute + ar      ", lambda: (render_views_fwd_bwd(act, cams, 3, up, bucket, extras=True, n_streams=0), bucket.all_reduce()))
for mode in ("sh", "rows"):
    for c in (4, 8):
        run("overlapped %-4s chunks %2d" % (mode, c), lambda: render_views_fwd_bwd(act, cams, 3, up, bucket, extras=True, all_reduce=True, comm_chunks=c, comm_mode=mode))
dist.destroy_process_group()

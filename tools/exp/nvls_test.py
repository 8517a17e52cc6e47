"""torchrun: NVLS two-shot all-reduce (tgr_multimem_allreduce_f32) vs NCCL on the gradient bucket: equality + timing."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from youreditableavatar_b200.parallel import GradBucket, SymmGradBucket
P = 1_000_000
b = SymmGradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
if rank == 0: print("nvls path:", b.nvls, "floats", b.flat.numel(), flush=True)
g = torch.Generator(device="cuda").manual_seed(rank)
src = torch.randn(b.flat.numel(), device="cuda", generator=g)
ref = src.clone(); dist.all_reduce(ref)
b.flat.copy_(src); b.all_reduce(); torch.cuda.synchronize()
err = (b.flat - ref).abs().max().item(); rel = err / ref.abs().max().item()
print("rank %d max abs diff vs NCCL %.3e (rel %.2e)" % (rank, err, rel), flush=True)
def timeit(fn, K=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(K)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
t_nvls = timeit(lambda: b.all_reduce())
t_nccl = timeit(lambda: dist.all_reduce(ref))
if rank == 0: print("all-reduce of %.0f MB: NVLS kernel %.3f ms, NCCL %.3f ms" % (b.flat.numel() * 4 / 1e6, t_nvls, t_nccl), flush=True)
dist.destroy_process_group()

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from youreditableavatar_b200 import _lib
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = torch.Generator().manual_seed(0)
keys = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64).to(torch.int32).cuda()
if bits < 31:
    keys = keys % (1 << bits)
vals = torch.arange(n, dtype=torch.int32, device="cuda")
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
temp = torch.empty(L.tgr_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    kin, vin = keys.clone(), vals.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.tgr_sort_pairs_u32(n, kin.data_ptr(), vin.data_ptr(), ko.data_ptr(), vo.data_ptr(), 0, bits, temp.data_ptr(), temp.numel(), st)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for _ in range(5): run()
ts = sorted(run() for _ in range(20))
print("n=%d bits=%d: median %.1f us (min %.1f)" % (n, bits, ts[10] * 1000, ts[0] * 1000))

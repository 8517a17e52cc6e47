import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import *
from oracle import ref_cuda
_, inp, cam = small_scene(50000, 96, 512, 2)
fo = ours_forward(inp, cam, 3); fr = ref_cuda.forward(inp, cam, 3)
torch.cuda.synchronize()
d = (fo[1] - fr[1]).abs()
print("nan ours", int(torch.isnan(fo[1]).sum()), "nan ref", int(torch.isnan(fr[1]).sum()))
bad = (d > 1e-4) | torch.isnan(d)
idx = bad.any(0).nonzero()
print("bad pixels", idx.shape[0])
fT, nc = export_image_state(50000, 512, 512, fo)
acc, ncr, rng = ref_cuda.decode_image(fr[5], 512 * 512, 32 * 32)
for y, x in idx[:12].tolist():
    p = y * 512 + x
    print((y, x), "tile", (y // 16, x // 16), "in-tile", (y % 16, x % 16), "ours", fo[1][:, y, x].tolist(), "ref", fr[1][:, y, x].tolist(), "T", float(fT[p]), float(acc[p]), "nc", int(nc[p]), int(ncr[p]))
ys = idx[:, 0] // 16; xs = idx[:, 1] // 16
tiles = torch.unique(ys * 32 + xs)
print("bad tiles", tiles.tolist()[:20], "lens", [(int(rng[t, 1] - rng[t, 0])) for t in tiles.tolist()[:20]])

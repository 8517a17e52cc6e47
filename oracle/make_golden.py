"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA rasterizer (oracle/_ref) on a B200.

Run on the GPU box:  python oracle/make_golden.py gpurun_out/golden   (then copy the files to tests/golden/)
Each fixture holds the fp32 inputs, the camera, and the reference's own outputs: image, radii, num_rendered,
per-Gaussian records (depth, means2D, conic+opacity, rgb, tiles_touched), sorted 64-bit keys, sorted ids,
tile ranges, final_T, n_contrib and the eight gradient tensors for a seeded upstream gradient.
These pin oracle/oracle.py (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import random_cloud, small_scene  # noqa: E402
from oracle import ref_cuda  # noqa: E402

GRADS = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]


def dump(name, inp, cam, degree, outdir):
    W, H = cam["image_width"], cam["image_height"]
    P = inp["means3D"].shape[0]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    g = torch.Generator().manual_seed(1)
    dL = (torch.randn(3, H, W, generator=g) / (3 * H * W)).cuda()
    fr = ref_cuda.forward(inp, cam, degree)
    gr = ref_cuda.backward(inp, cam, degree, fr, dL)
    torch.cuda.synchronize()
    R = fr[0]
    keys, ids = ref_cuda.decode_binning(fr[4], R)
    accum, ncon, ranges = ref_cuda.decode_image(fr[5], W * H, T)
    geo = ref_cuda.decode_geom(fr[3], P)
    vis = (fr[2] > 0)
    d = {"degree": np.int32(degree), "W": np.int32(W), "H": np.int32(H), "num_rendered": np.int64(R),
         "color": fr[1].cpu().numpy(), "radii": fr[2].cpu().numpy(), "dL_dcolor": dL.cpu().numpy(),
         "keys": keys.cpu().numpy().view(np.uint64), "ids": ids.cpu().numpy().view(np.uint32),
         "ranges": ranges.cpu().numpy().view(np.uint32), "final_T": accum.cpu().numpy(),
         "n_contrib": ncon.cpu().numpy().view(np.uint32)}
    for k, v in geo.items():  # culled entries of the reference's buffers are uninitialised: zero them
        a = v.clone()
        a[~vis] = 0
        d["geom_" + k] = a.cpu().numpy()
    for k, v in inp.items():
        d["in_" + k] = v.cpu().numpy()
    for k in ("viewmatrix", "projmatrix", "campos", "bg"):
        d["cam_" + k] = cam[k].cpu().numpy()
    d["cam_tanfovx"] = np.float64(cam["tanfovx"])
    d["cam_tanfovy"] = np.float64(cam["tanfovy"])
    for n, t in zip(GRADS, gr):
        d[n] = t.cpu().numpy()
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **d)
    print(name, "P", P, "R", R, "visible", int(vis.sum()))


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    _, inp, cam = small_scene(1500, 32, 64, 1)
    dump("shell_1500_64_sh3", inp, cam, 3, outdir)
    inp, cam = random_cloud(700, 77, 45, seed=4, big=True)
    dump("cloud_700_77x45_sh1", inp, cam, 1, outdir)
    _, inp, cam = small_scene(1200, 32, 48, 2)
    inp = dict(inp)
    inp["colors_precomp"] = torch.rand(1200, 3, generator=torch.Generator().manual_seed(5)).cuda()
    inp.pop("shs")
    dump("shell_1200_48_colors", inp, cam, 0, outdir)


if __name__ == "__main__":
    main()

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from youreditableavatar_b200 import scene
from helpers import ours_forward, export_binning, export_image_state
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
P, res, _, g = scene.CONFIGS[cfg]
gs = scene.make_scene(cfg, device="cuda"); act = scene.activate(gs)
for v in range(2):
    cam = scene.orbit_camera(v, 8, res, res, device="cuda")
    fo = ours_forward(act, cam, 3)
    keys, ids, ranges = export_binning(P, res, res, fo)
    fT, nc = export_image_state(P, res, res, fo)
    ln = (ranges[:, 1] - ranges[:, 0]).float()
    ne = ln[ln > 0]
    q = torch.quantile(ne, torch.tensor([0.5, 0.9, 0.99, 1.0], device=ne.device))
    T = res // 16
    ncm = nc.view(T, 16, T, 16).permute(0, 2, 1, 3).reshape(T * T, 256).float()
    last = ncm.max(1)[0]
    print("view", v, "R", fo[0], "nonempty tiles", ne.numel(), "len mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (ne.mean(), *q.tolist()))
    lne = last[ln > 0]
    q2 = torch.quantile(lne, torch.tensor([0.5, 0.9, 0.99, 1.0], device=ne.device))
    print("   tile_last mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f ; sum(last)/sum(len) %.2f" % (lne.mean(), *q2.tolist(), lne.sum() / ne.sum()))
    print("   pixels terminated early (T<1e-4 path): %.3f ; mean n_contrib over covered px %.0f" % (float((fT < 1e-3).float().mean()), float(nc[nc > 0].float().mean())))

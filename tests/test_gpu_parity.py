"""GPU parity tests (run with `-m gpu` on the B200 box).  Everything goes through the C ABI via the host-side
mirror of the reference operator API.  Three checkers, all test infrastructure:
  * the UNMODIFIED reference CUDA rasterizer built into oracle/_ref (bit-exact integer artefacts, images,
    gradients) — skipped if it was not built;
  * the float64 CPU oracle (oracle/oracle.py);
  * committed golden fixtures (tests/golden), see test_golden.py.
Tolerances (BASELINE.json north_star): keys / sort order / tile ranges bit-exact; images max-abs 1e-4;
gradients relative L2 1e-3.
"""
import numpy as np
import pytest
import torch

from youreditableavatar_b200 import scene
from helpers import (export_binning, export_geom, export_image_state, ours_backward, ours_forward, random_cloud,
                     rel_l2, small_scene, to_dev)

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3


def _ref():
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("reference CUDA build (oracle/_ref) not present")
    return ref_cuda


def _cases():
    # (name, builder) — avatar shells at several sizes plus unstructured clouds with ragged image sizes
    return [
        ("shell_3k_128", lambda: small_scene(3000, 32, 128, 0)[1:]),
        ("shell_10k_256", lambda: small_scene(10000, 32, 256, 1)[1:]),
        ("shell_50k_512", lambda: small_scene(50000, 96, 512, 2)[1:]),
        ("cloud_5k_200x120", lambda: random_cloud(5000, 200, 120, seed=3)),
        ("cloud_big_2k_77x45", lambda: random_cloud(2000, 77, 45, seed=4, big=True)),
    ]


@pytest.mark.parametrize("name,build", _cases())
@pytest.mark.parametrize("degree", [3, 0])
def test_forward_bitexact_vs_reference(name, build, degree):
    ref = _ref()
    inp, cam = build()
    P, W, H = inp["means3D"].shape[0], cam["image_width"], cam["image_height"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    ours = ours_forward(inp, cam, degree)
    theirs = ref.forward(inp, cam, degree)
    torch.cuda.synchronize()
    R_o, color_o, radii_o = ours[0], ours[1], ours[2]
    R_r, color_r, radii_r = theirs[0], theirs[1], theirs[2]

    # --- per-Gaussian records: bit-exact where visible ------------------------------------------
    assert torch.equal(radii_o, radii_r), "radii differ at %d Gaussians" % int((radii_o != radii_r).sum())
    vis = radii_r > 0
    go, gr = export_geom(P, W, H, ours), ref.decode_geom(theirs[3], P)
    assert torch.equal(go["tiles_touched"], gr["tiles_touched"].to(torch.int32))
    assert torch.equal(go["depths"][vis].view(torch.int32), gr["depths"][vis].view(torch.int32)), "depth bits"
    assert torch.equal(go["means2D"][vis].view(torch.int32), gr["means2D"][vis].view(torch.int32)), "means2D bits"
    assert torch.equal(go["conic_opacity"][vis].view(torch.int32), gr["conic_opacity"][vis].view(torch.int32)), "conic"
    assert (go["rgb"][vis] - gr["rgb"][vis]).abs().max().item() <= 1e-6
    assert R_o == R_r

    # --- keys, sorted order, tile ranges: bit-exact ----------------------------------------------
    keys_o, ids_o, ranges_o = export_binning(P, W, H, ours)
    keys_r, ids_r = ref.decode_binning(theirs[4], R_r)
    _, ncon_r, ranges_r = ref.decode_image(theirs[5], W * H, T)
    assert torch.equal(keys_o, keys_r), "sorted 64-bit keys differ"
    assert torch.equal(ids_o, ids_r), "sorted Gaussian ids differ"
    assert torch.equal(ranges_o, ranges_r), "tile ranges differ"

    # --- image + per-pixel state -------------------------------------------------------------------
    fT_o, ncon_o = export_image_state(P, W, H, ours)
    assert torch.equal(ncon_o, ncon_r), "n_contrib differs"
    err = (color_o - color_r).abs().max().item()
    assert err <= IMG_TOL, "image max-abs %g" % err


@pytest.mark.parametrize("name,build", _cases())
def test_backward_vs_reference(name, build):
    ref = _ref()
    inp, cam = build()
    W, H = cam["image_width"], cam["image_height"]
    g = torch.Generator().manual_seed(1)
    dL = (torch.randn(3, H, W, generator=g) / (3 * H * W)).cuda()
    fo = ours_forward(inp, cam, 3)
    fr = ref.forward(inp, cam, 3)
    go = ours_backward(inp, cam, 3, fo, dL)
    gr = ref.backward(inp, cam, 3, fr, dL)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for n, a, b in zip(names, go, gr):
        assert a.shape == b.shape, n
        assert torch.isfinite(a).all(), n
        e = rel_l2(a, b)
        assert e <= GRAD_TOL, "%s rel-L2 %g" % (n, e)


@pytest.mark.parametrize("variant", ["sh", "colors", "cov3d"])
def test_vs_float64_oracle(variant):
    from oracle import oracle
    _, inp, cam = small_scene(2500, 32, 96, 1)
    W, H = cam["image_width"], cam["image_height"]
    P = inp["means3D"].shape[0]
    inp = dict(inp)
    degree = 3
    if variant == "colors":
        inp["colors_precomp"] = torch.rand(P, 3, generator=torch.Generator().manual_seed(5)).cuda()
        inp.pop("shs")
    if variant == "cov3d":
        rec = oracle.preprocess(inp["means3D"].cpu(), inp["opacities"].cpu(), cam["viewmatrix"].cpu(),
                                cam["projmatrix"].cpu(), cam["campos"].cpu(), W, H, cam["tanfovx"], cam["tanfovy"],
                                scales=inp["scales"].cpu(), rotations=inp["rotations"].cpu(), shs=inp["shs"].cpu(),
                                degree=degree)
        s = inp["scales"].double().cpu()
        q = inp["rotations"].double().cpu()
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Rm = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
                          torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
                          torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], -2)
        S = Rm @ torch.diag_embed(s * s) @ Rm.transpose(1, 2)
        inp["cov3D_precomp"] = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1).float().cuda()
        inp.pop("scales"); inp.pop("rotations")

    g = torch.Generator().manual_seed(1)
    dLc = torch.randn(3, H, W, generator=g) / (3 * H * W)
    dLd = torch.randn(1, H, W, generator=g) / (H * W)
    dLa = torch.randn(1, H, W, generator=g) / (H * W)

    fo = ours_forward(inp, cam, degree, extras=True)
    go = ours_backward(inp, cam, degree, fo, dLc.cuda(), dLd.cuda(), dLa.cuda())
    torch.cuda.synchronize()
    geo = export_geom(P, W, H, fo)

    oin = {k: v.detach().double().cpu().requires_grad_(True) for k, v in inp.items()}
    out = oracle.rasterize(oin, to_dev(cam, "cpu"), degree,
                           xy_radii_override=(geo["means2D"].cpu().numpy(), fo[2].cpu().numpy()),
                           depth_override=geo["depths"].cpu().numpy())
    # integer artefacts given the same fp32 records: bit-exact
    keys_o, ids_o, ranges_o = export_binning(P, W, H, fo)
    assert np.array_equal(keys_o.cpu().numpy().view(np.uint64), out["keys"])
    assert np.array_equal(ids_o.cpu().numpy().view(np.uint32), out["ids"])
    assert np.array_equal(ranges_o.cpu().numpy().view(np.uint32), out["ranges"])
    assert fo[0] == out["num_rendered"]
    # the oracle's own radii agree except at fp32/fp64 ceil boundaries (none expected at this size)
    assert (fo[2].cpu().numpy() != out["radii"]).mean() <= 1e-3

    def img_ok(a, b, what):
        d = (a.double().cpu() - b.detach()).abs()
        # a pair whose alpha sits within fp32 rounding of the 1/255 cut-off may flip: allow <= 1e-4 of pixels
        assert (d > IMG_TOL).double().mean().item() <= 1e-4 and d.max().item() <= 1e-2, "%s max %g" % (what, d.max())

    img_ok(fo[1], out["color"], "color")
    img_ok(fo[6], out["depth"], "depth")
    img_ok(fo[7], out["alpha"], "alpha")
    loss = (out["color"] * dLc.double()).sum() + (out["depth"] * dLd.double()).sum() + (out["alpha"] * dLa.double()).sum()
    loss.backward()
    pairs = [("dL_dmeans3D", go[3], oin["means3D"].grad)]
    if variant != "cov3d":
        pairs += [("dL_dscales", go[6], oin["scales"].grad), ("dL_drotations", go[7], oin["rotations"].grad)]
    else:
        pairs += [("dL_dcov3D", go[4], oin["cov3D_precomp"].grad)]
    if variant == "colors":
        pairs += [("dL_dcolors", go[1], oin["colors_precomp"].grad)]
    else:
        pairs += [("dL_dsh", go[5], oin["shs"].grad)]
    pairs += [("dL_dopacity", go[2], oin["opacities"].grad)]
    for n, a, b in pairs:
        e = rel_l2(a, b)
        assert e <= 3 * GRAD_TOL, "%s rel-L2 vs oracle %g" % (n, e)


def test_extras_do_not_change_colour():
    _, inp, cam = small_scene(3000, 32, 128, 0)
    a = ours_forward(inp, cam, 3, extras=False)
    b = ours_forward(inp, cam, 3, extras=True)
    assert torch.equal(a[1], b[1])
    assert b[6].shape == (1, 128, 128) and b[7].shape == (1, 128, 128)
    assert float(b[7].min()) >= 0 and float(b[7].max()) <= 1


def test_empty_and_fully_culled():
    _, inp, cam = small_scene(3000, 32, 64, 0)
    dev = inp["means3D"].device
    # P = 0: zeros image, empty buffers, nothing launched (rasterize_points.cu:81)
    empty = {k: v[:0] for k, v in inp.items()}
    out = ours_forward(empty, cam, 3)
    assert out[0] == 0 and out[1].shape == (3, 64, 64) and float(out[1].abs().max()) == 0 and out[2].numel() == 0
    # everything behind the camera: background only, R = 0, gradients all zero
    behind = dict(inp)
    behind["means3D"] = inp["means3D"] + cam["campos"][None] * 3.0
    out = ours_forward(behind, cam, 3)
    assert out[0] == 0 and int(out[2].abs().sum()) == 0
    assert torch.allclose(out[1], cam["bg"][:, None, None].expand(3, 64, 64))
    grads = ours_backward(behind, cam, 3, out, torch.ones(3, 64, 64, device=dev))
    for g in grads:
        assert float(g.abs().max()) == 0


def test_dropin_module_autograd():
    """The call the Edit_core scene models make (tetgs_model.py:605-614), gradients through autograd."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    _, inp, cam = small_scene(3000, 32, 128, 0)
    P = inp["means3D"].shape[0]
    settings = GaussianRasterizationSettings(
        image_height=128, image_width=128, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=cam["bg"],
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=3,
        campos=cam["campos"], prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=settings)
    leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
    image, radii = rasterizer(means3D=leaves["means3D"], means2D=means2D, shs=leaves["shs"], colors_precomp=None,
                              opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"],
                              cov3D_precomp=None)
    assert image.shape == (3, 128, 128) and radii.shape == (P,) and radii.dtype == torch.int32
    image.square().mean().backward()
    for k, v in leaves.items():
        assert v.grad is not None and torch.isfinite(v.grad).all(), k
    assert means2D.grad is not None and float(means2D.grad.abs().sum()) > 0
    vis = rasterizer.markVisible(inp["means3D"])
    assert vis.dtype == torch.bool and vis.shape == (P,)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rasterizer(means3D=inp["means3D"], means2D=means2D, opacities=inp["opacities"], scales=inp["scales"],
                   rotations=inp["rotations"])
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rasterizer(means3D=inp["means3D"].reshape(-1), means2D=means2D, shs=inp["shs"], opacities=inp["opacities"],
                   scales=inp["scales"], rotations=inp["rotations"])


def test_mark_visible_matches_reference():
    """`_C.mark_visible` (rasterize_points.cu:198-217 -> checkFrustum, rasterizer_impl.cu:54-66): the near-plane test of
    in_frustum, bit for bit against the reference build — shells (all visible), a cloud straddling the camera, points
    exactly on the 0.2 plane, NaN / inf coordinates, and the empty set."""
    ref = _ref()
    from youreditableavatar_b200.rasterizer import c_mark_visible
    _, inp, cam = small_scene(20000, 48, 128, 1)
    cloud, ccam = random_cloud(50000, 200, 120, seed=9)
    for means, c in ((inp["means3D"], cam), (cloud["means3D"] * 3.0, ccam)):
        m = means.clone()
        # a handful of special rows: on the plane (z_view == 0.2 up to rounding), NaN, +-inf
        vm = c["viewmatrix"]
        Rm, t = vm[:3, :3], vm[3, :3]                       # row-vector convention: p_view = p @ Rm + t
        on_plane = (torch.tensor([[0.0, 0.0, 0.2], [0.3, -0.1, 0.2], [0.0, 0.0, 0.20000002]], device=m.device) - t) @ torch.linalg.inv(Rm)
        m[:3] = on_plane
        m[3] = float("nan")
        m[4, 0] = float("inf")
        m[5, 2] = float("-inf")
        mine = c_mark_visible(m, c["viewmatrix"], c["projmatrix"])
        theirs = ref.dgr().mark_visible(m, c["viewmatrix"], c["projmatrix"])
        assert mine.dtype == torch.bool and torch.equal(mine, theirs), int((mine != theirs).sum())
        assert 0 < int(mine.sum()) <= m.shape[0]
    empty = c_mark_visible(inp["means3D"][:0], cam["viewmatrix"], cam["projmatrix"])
    assert empty.shape == (0,) and empty.dtype == torch.bool
    # radii > 0 implies visible: the rasterizer's own culling uses the same test
    out = ours_forward(cloud, ccam, 3)
    assert bool((c_mark_visible(cloud["means3D"], ccam["viewmatrix"], ccam["projmatrix"]) | (out[2] == 0)).all())


def test_prefiltered_semantics():
    """`prefiltered=True` is the caller's promise that every Gaussian passes the near-plane test (it filtered with
    mark_visible).  Kept promise: identical results to prefiltered=False (the reference too: the flag only arms the
    check).  Broken promise: the reference prints "Point is filtered although prefiltered is set..." and __trap()s
    (auxiliary.h:154-160), which kills the CUDA context; here the same message is raised as an exception and the
    context stays usable — single-view operator, autograd module and multi-view batch."""
    ref = _ref()
    from youreditableavatar_b200 import rasterizer as rz, multiview as mv
    from youreditableavatar_b200.parallel import settings_from_cam
    cloud, cam = random_cloud(6000, 200, 120, seed=3)
    vis = rz.c_mark_visible(cloud["means3D"], cam["viewmatrix"], cam["projmatrix"])
    assert 0 < int(vis.sum()) < 6000
    kept = {k: v[vis].contiguous() for k, v in cloud.items()}
    e = torch.Tensor([])

    def fwd(inp, prefiltered):
        return rz.c_rasterize_gaussians(cam["bg"], inp["means3D"], e, inp["opacities"], inp["scales"], inp["rotations"], 1.0,
                                        e, cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
                                        cam["image_height"], cam["image_width"], inp["shs"], 3, cam["campos"], prefiltered, False)

    a, b = fwd(kept, True), fwd(kept, False)
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    theirs = ref.dgr().rasterize_gaussians(cam["bg"], kept["means3D"], e, kept["opacities"], kept["scales"], kept["rotations"],
                                           1.0, e, cam["viewmatrix"], cam["projmatrix"], float(cam["tanfovx"]),
                                           float(cam["tanfovy"]), cam["image_height"], cam["image_width"], kept["shs"], 3,
                                           cam["campos"], True, False)       # the reference with the promise kept
    assert theirs[0] == a[0] and torch.equal(theirs[2], a[2]) and (theirs[1] - a[1]).abs().max().item() <= IMG_TOL
    with pytest.raises(RuntimeError, match="Point is filtered although prefiltered is set"):
        fwd(cloud, True)
    c = fwd(cloud, False)                                   # the context is alive and the unfiltered call is unchanged
    assert c[0] == a[0] and torch.equal(c[1], a[1])
    sets = [settings_from_cam(cam, 3)._replace(prefiltered=True)] * 2
    with pytest.raises(RuntimeError, match="Point is filtered although prefiltered is set"):
        mv.c_rasterize_views(sets, cloud["means3D"], e, cloud["opacities"], cloud["scales"], cloud["rotations"], e, cloud["shs"])
    res = mv.c_rasterize_views(sets, kept["means3D"], e, kept["opacities"], kept["scales"], kept["rotations"], e, kept["shs"])
    assert torch.equal(res[1][0], a[1]) and torch.equal(res[1][1], a[1])


@pytest.mark.parametrize("n,bits", [(1, 32), (1000, 32), (4096, 12), (4097, 8), (1 << 20, 32), (3_000_001, 13), (50000, 31)])
def test_radix_sort_stable(n, bits):
    import ctypes as C
    from youreditableavatar_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 2 ** min(bits, 31), (n,), generator=g, dtype=torch.int64).to(torch.int32).cuda()
    if bits <= 16:   # many duplicates: stability matters
        keys = keys % (1 << bits)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    temp = torch.empty(L.tgr_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
    kin, vin = keys.clone(), vals.clone()
    rc = L.tgr_sort_pairs_u32(n, kin.data_ptr(), vin.data_ptr(), ko.data_ptr(), vo.data_ptr(), 0, bits, temp.data_ptr(),
                              temp.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.tgr_last_error()
    torch.cuda.synchronize()
    ks, order = torch.sort(keys.long() & ((1 << bits) - 1), stable=True)
    assert torch.equal(vo.long(), order), "not the stable order"
    assert torch.equal(ko.long() & ((1 << bits) - 1), ks)


def test_dist2_vs_reference_and_oracle():
    from oracle import oracle
    from simple_knn._C import distCUDA2
    g = torch.Generator().manual_seed(7)
    pts = torch.randn(20000, 3, generator=g).cuda()
    mine = distCUDA2(pts)
    want = oracle.dist2_knn3(pts.cpu()).float()
    assert torch.allclose(mine.cpu(), want, rtol=1e-4, atol=1e-9)
    from oracle import ref_cuda
    if ref_cuda.available():
        theirs = ref_cuda.knn().distCUDA2(pts)
        assert torch.allclose(mine, theirs, rtol=1e-5, atol=1e-9)
    # mesh-bound points (clustered, duplicates possible)
    _, inp, _ = small_scene(30000, 64, 64, 0)
    m = distCUDA2(inp["means3D"])
    w = oracle.dist2_knn3(inp["means3D"].cpu()).float()
    assert torch.allclose(m.cpu(), w, rtol=1e-4, atol=1e-10)


def test_fused_binding_matches_eager_binding_and_oracle():
    """BOUND kernels: same image as feeding the eagerly activated tensors; raw-parameter gradients vs the
    float64 oracle chained through oracle.bind (tetgs_model.py:252-286)."""
    from oracle import oracle
    from youreditableavatar_b200 import scene
    from youreditableavatar_b200.binding import MeshBinding, rasterize_bound
    from youreditableavatar_b200.rasterizer import GaussianRasterizationSettings
    gs, act, cam = small_scene(2500, 32, 96, 1)
    mesh = MeshBinding.from_scene(gs)
    raw = {k: gs[k].cuda().clone().requires_grad_(True) for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs")}
    settings = GaussianRasterizationSettings(96, 96, cam["tanfovx"], cam["tanfovy"], cam["bg"], 1.0, cam["viewmatrix"],
                                             cam["projmatrix"], 3, cam["campos"], False, False)
    color, radii = rasterize_bound(raw["delta"], raw["log_scales"], raw["raw_quats"], raw["opacity_logits"], raw["shs"],
                                   mesh, settings)
    eager = ours_forward(act, cam, 3)
    # fp32 activations in-kernel vs torch's eager ones may differ in the last bit -> allow a few threshold flips
    d = (color - eager[1]).abs()
    assert (d > IMG_TOL).float().mean().item() <= 1e-4 and d.max().item() <= 1e-2
    assert (radii != eager[2]).float().mean().item() <= 1e-3

    g = torch.Generator().manual_seed(1)
    dL = torch.randn(3, 96, 96, generator=g) / (3 * 96 * 96)
    (color * dL.cuda()).sum().backward()

    gs64 = {k: (v.double() if (isinstance(v, torch.Tensor) and v.is_floating_point()) else v) for k, v in gs.items()}
    leaves = {k: gs64[k].clone().requires_grad_(True) for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs")}
    gs64.update(leaves)
    bound = oracle.bind(gs64)
    bound["shs"] = leaves["shs"]
    geo = export_geom(2500, 96, 96, eager)
    out = oracle.rasterize(bound, to_dev(cam, "cpu"), 3,
                           xy_radii_override=(geo["means2D"].cpu().numpy(), eager[2].cpu().numpy()),
                           depth_override=geo["depths"].cpu().numpy())
    (out["color"] * dL.double()).sum().backward()
    for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs"):
        e = rel_l2(raw[k].grad, leaves[k].grad)
        assert e <= 3 * GRAD_TOL, "%s rel-L2 vs oracle %g" % (k, e)


def test_accumulate_mode_sums_views():
    from youreditableavatar_b200.parallel import GradBucket, render_batch_fwd_bwd
    from youreditableavatar_b200 import scene
    _, inp, _ = small_scene(3000, 32, 128, 0)
    cams = [to_dev(scene.orbit_camera(v, 3, 128, 128), "cuda") for v in range(3)]
    g = torch.Generator().manual_seed(2)
    dLs = [(torch.randn(3, 128, 128, generator=g) / (3 * 128 * 128)).cuda() for _ in cams]
    bucket = GradBucket(3000, 16, "cuda")
    render_batch_fwd_bwd(inp, cams, 3, lambda i, c, d, a: (dLs[i], None, None), bucket)
    want = None
    for cam, dL in zip(cams, dLs):
        fo = ours_forward(inp, cam, 3)
        go = ours_backward(inp, cam, 3, fo, dL)
        want = [x.clone() for x in go] if want is None else [w + x for w, x in zip(want, go)]
    for name, a, b in zip(["m2D", "col", "op", "m3D", "cov", "sh", "sc", "rot"], bucket.views, want):
        assert rel_l2(a, b) <= 1e-5, name
    # training subset: skipped gradients are simply not produced
    t = GradBucket(3000, 16, "cuda", names=GradBucket.TRAINING)
    render_batch_fwd_bwd(inp, cams, 3, lambda i, c, d, a: (dLs[i], None, None), t)
    assert t.views[0] is None and rel_l2(t.views[5], want[5]) <= 1e-5 and rel_l2(t.views[3], want[3]) <= 1e-5


def test_work_queue_covers_every_tile_many_sizes():
    """Regression for the SM-affine tile queues: every tile (also empty ones) must be rendered exactly once."""
    from youreditableavatar_b200 import scene
    _, inp, _ = small_scene(3000, 32, 64, 0)
    for (w, h) in [(16, 16), (33, 17), (640, 48), (1000, 1000), (2048, 96)]:
        cam = to_dev(scene.orbit_camera(1, 4, h, w), "cuda")
        for _ in range(3):
            out = ours_forward(inp, cam, 3, extras=True)
            assert torch.isfinite(out[1]).all()
            # background is white: an unrendered (garbage / zero) tile shows up as alpha+colour inconsistency
            bgmask = out[7][0] == 0
            assert torch.allclose(out[1][:, bgmask], torch.ones_like(out[1][:, bgmask]))


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams", [0, 1, 3])
@pytest.mark.parametrize("V", [1, 3, 11])
def test_multiview_batch_matches_per_view_calls(V, n_streams):
    """tgr_*_batch: per-view forward results are bit-identical to single-view calls; the summed gradients agree
    with per-view calls + accumulate to fp32 summation order (V = 11 spans two chunks of TGR_MAX_BATCH)."""
    from youreditableavatar_b200 import multiview as mv
    from youreditableavatar_b200.parallel import settings_from_cam
    _, inp, _ = small_scene(2500, 32, 112, 0)
    P = inp["means3D"].shape[0]
    cams = [to_dev(scene.orbit_camera(v, V, 112, 112, device="cpu"), "cuda") for v in range(V)]
    g = torch.Generator().manual_seed(5)
    dLc = (torch.randn(V, 3, 112, 112, generator=g) / (3 * 112 * 112)).cuda()
    dLd = (torch.randn(V, 1, 112, 112, generator=g) / (112 * 112)).cuda()
    dLa = (torch.randn(V, 1, 112, 112, generator=g) / (112 * 112)).cuda()
    e = torch.Tensor([])
    res = mv.c_rasterize_views([settings_from_cam(c, 3) for c in cams], inp["means3D"], e, inp["opacities"], inp["scales"],
                               inp["rotations"], e, inp["shs"], extras=True, n_streams=n_streams)
    state, color, radii, depth, alpha = res
    grads = mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa)
    ref_sum = None
    for v in range(V):
        fo = ours_forward(inp, cams[v], 3, extras=True)
        assert fo[0] == state.counts[v]
        assert torch.equal(fo[1], color[v]) and torch.equal(fo[2], radii[v])
        assert torch.equal(fo[6], depth[v]) and torch.equal(fo[7], alpha[v])
        go = ours_backward(inp, cams[v], 3, fo, dLc[v], dLd[v], dLa[v])
        ref_sum = [x.clone() for x in go] if ref_sum is None else [a + b for a, b in zip(ref_sum, go)]
    for name, a, b in zip(["m2D", "col", "op", "m3D", "cov", "sh", "sc", "rot"], grads, ref_sum):
        assert rel_l2(a, b) <= 2e-6, (name, rel_l2(a, b))
    # the Gaussian-range variant (ranges of the final kernel, used to overlap the all-reduce) gives the same sums
    # (not the same bits: the blend stage's floating-point reductions are re-run in a different order)
    seen = []
    grads_r = mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, chunks=3,
                                            on_chunk=lambda first, count: seen.append((first, count)))
    assert sum(c for _, c in seen) == P and all(f % 256 == 0 for f, _ in seen) and len(seen) == 3
    for a, b in zip(grads_r, grads):
        assert rel_l2(a, b) <= 2e-6
    # accumulate=True adds a second batch on top
    grads2 = mv.c_rasterize_views_backward(state, dLc, dL_dout_depth=dLd, dL_dout_alpha=dLa, accumulate_into=grads)
    for a, b in zip(grads2, ref_sum):
        assert rel_l2(a, 2 * b) <= 2e-6


@pytest.mark.gpu
def test_multiview_module_autograd_matches_single_view_modules():
    from youreditableavatar_b200.multiview import MultiViewRasterizer
    from youreditableavatar_b200.parallel import settings_from_cam
    from diff_gaussian_rasterization import GaussianRasterizer
    _, inp, _ = small_scene(1500, 32, 96, 0)
    V = 4
    cams = [to_dev(scene.orbit_camera(v, V, 96, 96, device="cpu"), "cuda") for v in range(V)]
    settings = [settings_from_cam(c, 3) for c in cams]
    leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth, alpha = MultiViewRasterizer(settings, extra_outputs=True, n_streams=2)(
        means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"])
    w = torch.linspace(0.5, 1.5, V, device="cuda").view(V, 1, 1, 1)
    ((color * w).sum() + 0.3 * depth.sum() - 0.2 * alpha.sum()).backward()
    leaves1 = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    loss = 0
    for v in range(V):
        c1, r1, d1, a1 = GaussianRasterizer(settings[v], extra_outputs=True)(
            means3D=leaves1["means3D"], means2D=torch.zeros_like(m2, requires_grad=True), opacities=leaves1["opacities"],
            shs=leaves1["shs"], scales=leaves1["scales"], rotations=leaves1["rotations"])
        assert torch.equal(c1, color[v]) and torch.equal(r1, radii[v])
        loss = loss + (c1 * w[v]).sum() + 0.3 * d1.sum() - 0.2 * a1.sum()
    loss.backward()
    for k in leaves:
        assert rel_l2(leaves[k].grad, leaves1[k].grad) <= 2e-6, k


@pytest.mark.gpu
def test_multiview_capacity_hints_sync_free_path_and_overflow_fallback():
    """Second and later batches of a shape launch binning / blending from a capacity hint without waiting for the
    instance counts: same bits as the synchronous first batch; a hint that is too small (device-side overflow
    detection) falls back to the synchronous path with identical results."""
    from youreditableavatar_b200 import multiview as mv
    from youreditableavatar_b200.parallel import settings_from_cam
    _, inp, _ = small_scene(3000, 32, 128, 0)
    V = 3
    cams = [to_dev(scene.orbit_camera(v, V, 128, 128, device="cpu"), "cuda") for v in range(V)]
    sets = [settings_from_cam(c, 3) for c in cams]
    g = torch.Generator().manual_seed(7)
    dLc = (torch.randn(V, 3, 128, 128, generator=g) / (3 * 128 * 128)).cuda()
    e = torch.Tensor([])
    args = (sets, inp["means3D"], e, inp["opacities"], inp["scales"], inp["rotations"], e, inp["shs"])

    mv.set_capacity_hints(True)
    try:
        st0, color0, radii0, depth0, alpha0 = mv.c_rasterize_views(*args, extras=True)
        assert st0.caps == st0.counts                                  # first batch of this shape: exact sizes
        g0 = mv.c_rasterize_views_backward(st0, dLc)
        st1, color1, radii1, depth1, alpha1 = mv.c_rasterize_views(*args, extras=True)
        assert st1.counts == st0.counts and min(st1.caps) > max(st1.counts)   # launched from the hint
        assert torch.equal(color1, color0) and torch.equal(radii1, radii0)
        assert torch.equal(depth1, depth0) and torch.equal(alpha1, alpha0)
        g1 = mv.c_rasterize_views_backward(st1, dLc)
        for a, b in zip(g1, g0):
            assert rel_l2(a, b) <= 2e-6
        # a hint far too small: the kernels flag the overflow on the device, the batch is redone synchronously
        for k in list(mv._capacity["hints"]):
            mv._capacity["hints"][k] = (1, mv._capacity["hints"][k][1])
        margin = mv._capacity["margin"]
        mv._capacity["margin"] = 0
        try:
            st2, color2, radii2, _, _ = mv.c_rasterize_views(*args, extras=True)
        finally:
            mv._capacity["margin"] = margin
        assert st2.caps == st2.counts == st0.counts
        assert torch.equal(color2, color0) and torch.equal(radii2, radii0)
        g2 = mv.c_rasterize_views_backward(st2, dLc)
        for a, b in zip(g2, g0):
            assert rel_l2(a, b) <= 2e-6
        # switched off: always the synchronous path
        mv.set_capacity_hints(False)
        st3 = mv.c_rasterize_views(*args, extras=True)[0]
        st4 = mv.c_rasterize_views(*args, extras=True)[0]
        assert st3.caps == st3.counts and st4.caps == st4.counts
    finally:
        mv.set_capacity_hints(True)


@pytest.mark.gpu
def test_depth_sort_key_bit_hints_and_their_fallback():
    """The depth sort only looks at the low bits in which the visible depth keys differ (OR / AND of the keys in the
    header): calls that follow a first one of the same scene sort hint + 1 bits.  A hint that is too small gives a
    wrong order on the device; the host notices it from the header mirror and renders again — same bits as a full
    32-bit sort, single-view operator and multi-view batch."""
    from youreditableavatar_b200 import multiview as mv, rasterizer as rz
    from youreditableavatar_b200.parallel import settings_from_cam
    _, inp, cam = small_scene(20000, 48, 256, 1)
    P = inp["means3D"].shape[0]
    dev = torch.cuda.current_device()
    rz._depth_bits_hint.pop((dev, P), None)
    full = ours_forward(inp, cam, 3)                      # no hint yet: all 32 bits
    need = rz._depth_bits_hint[(dev, P)]
    assert 9 <= need <= 30                                # an avatar at distance 3: exponent and sign are shared
    hinted = ours_forward(inp, cam, 3)                    # need + 1 bits
    k_full, i_full, r_full = export_binning(P, 256, 256, full)
    k_hint, i_hint, r_hint = export_binning(P, 256, 256, hinted)
    assert torch.equal(k_full, k_hint) and torch.equal(i_full, i_hint) and torch.equal(r_full, r_hint)
    assert torch.equal(full[1], hinted[1])
    rz._depth_bits_hint[(dev, P)] = 4                     # far too few bits: must be noticed and redone
    redo = ours_forward(inp, cam, 3)
    k_redo, i_redo, _ = export_binning(P, 256, 256, redo)
    assert torch.equal(k_full, k_redo) and torch.equal(i_full, i_redo) and torch.equal(full[1], redo[1])
    assert rz._depth_bits_hint[(dev, P)] == need
    # batch path
    V = 3
    cams = [to_dev(scene.orbit_camera(v, V, 256, 256, device="cpu"), "cuda") for v in range(V)]
    sets = [settings_from_cam(c, 3) for c in cams]
    e = torch.Tensor([])
    args = (sets, inp["means3D"], e, inp["opacities"], inp["scales"], inp["rotations"], e, inp["shs"])
    mv.set_capacity_hints(True)
    try:
        st0, color0, radii0 = mv.c_rasterize_views(*args)
        for k in list(mv._capacity["hints"]):
            mv._capacity["hints"][k] = (mv._capacity["hints"][k][0], 3)
        st1, color1, radii1 = mv.c_rasterize_views(*args)
        assert st1.caps == st1.counts == st0.counts        # redone the synchronous way
        assert torch.equal(color1, color0) and torch.equal(radii1, radii0)
        st2, color2, _ = mv.c_rasterize_views(*args)       # and the hint is healthy again
        assert min(st2.caps) > max(st2.counts) and torch.equal(color2, color0)
    finally:
        mv.set_capacity_hints(True)


@pytest.mark.gpu
def test_single_view_capacity_hints_launch_ahead_and_overflow_fallback():
    """The drop-in single-view operator: the first call of a shape waits for num_rendered before it sizes the binning
    buffer (as the reference does, rasterizer_impl.cu:277-281); later calls launch binning / sort / blending from the
    previous count x slack and read the count afterwards.  Same bits either way, the backward recovers the buffer layout
    from its size, and a view that outgrows its hint (device-side check) is redone synchronously."""
    from youreditableavatar_b200 import _lib, rasterizer as rz
    L = _lib.lib()
    _, inp, cam = small_scene(20000, 48, 256, 1)
    P = inp["means3D"].shape[0]
    g = torch.Generator().manual_seed(3)
    dL = (torch.randn(3, 256, 256, generator=g) / (3 * 256 * 256)).cuda()
    rz.set_capacity_hints(True)
    try:
        a = ours_forward(inp, cam, 3, extras=True)                     # no hint yet: exact size
        cap_a = L.tgr_binning_capacity(P, a[4].numel(), 256, 256)    # every layout component is monotone in the capacity:
        assert a[0] <= cap_a < a[0] + 32 and L.tgr_binning_bytes(P, cap_a, 256, 256) == a[4].numel()   # same size = same layout
        ga = ours_backward(inp, cam, 3, a, dL)
        b = ours_forward(inp, cam, 3, extras=True)                     # launched ahead of the count
        assert b[0] == a[0] and b[4].numel() > a[4].numel()
        for x, y in zip(a[1:3] + a[6:8], b[1:3] + b[6:8]):
            assert torch.equal(x, y)
        ka, ia, ra = export_binning(P, 256, 256, a)
        kb, ib, rb = export_binning(P, 256, 256, b)
        assert torch.equal(ka, kb) and torch.equal(ia, ib) and torch.equal(ra, rb)
        gb = ours_backward(inp, cam, 3, b, dL)
        for x, y in zip(ga, gb):
            assert rel_l2(y, x) <= 2e-6
        # a hint that is far too small: the device renders nothing, the host notices and redoes the call
        for k in list(rz._capacity["hints"]):
            rz._capacity["hints"][k] = 1
        margin = rz._capacity["margin"]
        rz._capacity["margin"] = 0
        try:
            c = ours_forward(inp, cam, 3, extras=True)
        finally:
            rz._capacity["margin"] = margin
        assert c[0] == a[0] and torch.equal(c[1], a[1]) and c[4].numel() == a[4].numel()
        # the cheaper reading of the reference's signature still holds: R and the buffers are all the backward needs
        with pytest.raises(RuntimeError, match="too small"):
            ours_backward(inp, cam, 3, (a[0],) + a[1:4] + (a[4][:1024],) + a[5:], dL)
    finally:
        rz.set_capacity_hints(True)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["edit3d", "edit2d"])
def test_fused_direct_binding_of_the_edit_models(variant):
    """DirectBinding through the BOUND kernels: the keep + edit concatenation of the edit stages rendered straight from
    the raw tensors.  edit3d: points = cat(keep_xyz, ori_edit + n * delta) (tetgs_edit_3d.py:272-280); edit2d: fixed points,
    flat Gaussians in the triangle frame with scales (1e-8, d, d) (tetgs_edit_2d.py:172-208, 279-282).  Image = the eager
    concatenation fed to the plain operator; gradients of the edit part vs the float64 restatement (oracle.bind_direct),
    gradient rows of the frozen keep part exactly zero."""
    from oracle import oracle
    from youreditableavatar_b200 import formats
    from youreditableavatar_b200.binding import DirectBinding, rasterize_bound
    from youreditableavatar_b200.rasterizer import GaussianRasterizationSettings
    gs, act, cam = small_scene(2400, 32, 96, 1)
    P = 2400
    K = 1500                                              # keep part: the first K Gaussians, taken over activated
    keep_xyz = act["means3D"][:K].clone()
    fi = gs["face_index"].long()
    tri = gs["verts"][gs["faces"].long()[fi]]
    w = gs["bary"][..., None]
    ori = (tri * w).sum(1).cuda()
    nrm = (gs["vert_normals"][gs["faces"].long()[fi]] * w).sum(1).cuda()
    raw = {k: gs[k].cuda().clone() for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs")}
    if variant == "edit2d":
        # fresh flat Gaussians on the edit faces: quaternion of the triangle frame, scales (1e-8, d, d)
        flat = formats.bind_edit_gaussians(gs["verts"], gs["faces"].long(), sh_coeffs=16)
        sel = torch.arange(K, P) % flat["log_scales"].shape[0]
        raw["log_scales"][K:] = flat["log_scales"][sel].cuda()
        raw["raw_quats"][K:] = flat["raw_quats"][sel].cuda()
        binding = DirectBinding.from_keep_edit(keep_xyz, ori[K:], None)
        delta_in = None
    else:
        binding = DirectBinding.from_keep_edit(keep_xyz, ori[K:], nrm[K:])
        delta_in = raw["delta"].reshape(-1, 1).clone().requires_grad_(True)      # [P,1] like `_edit_points`
    leaves = {k: raw[k].clone().requires_grad_(True) for k in ("log_scales", "raw_quats", "shs")}
    leaves["opacity_logits"] = raw["opacity_logits"].reshape(-1, 1).clone().requires_grad_(True)   # [P,1] like all_densities
    settings = GaussianRasterizationSettings(96, 96, cam["tanfovx"], cam["tanfovy"], cam["bg"], 1.0, cam["viewmatrix"],
                                             cam["projmatrix"], 3, cam["campos"], False, False)
    color, radii = rasterize_bound(delta_in, leaves["log_scales"], leaves["raw_quats"], leaves["opacity_logits"],
                                   leaves["shs"], binding, settings)
    # eager concatenation -> plain operator
    b64 = oracle.bind_direct(binding.origins.cpu(), None if binding.normals is None else binding.normals.cpu(),
                             None if delta_in is None else raw["delta"].cpu(), raw["log_scales"].cpu(), raw["raw_quats"].cpu(),
                             raw["opacity_logits"].cpu(), n_keep=K)
    eager_in = {k: v.float().cuda().contiguous() for k, v in b64.items()}
    eager_in["shs"] = raw["shs"]
    eager = ours_forward(eager_in, cam, 3)
    d = (color - eager[1]).abs()
    assert (d > IMG_TOL).float().mean().item() <= 1e-4 and d.max().item() <= 1e-2
    assert (radii != eager[2]).float().mean().item() <= 1e-3

    g = torch.Generator().manual_seed(1)
    dL = torch.randn(3, 96, 96, generator=g) / (3 * 96 * 96)
    (color * dL.cuda()).sum().backward()
    # float64 restatement with autograd; the keep rows enter detached
    o_leaves = {k: raw[k].double().cpu().clone().requires_grad_(True) for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs")}
    bound = oracle.bind_direct(binding.origins.cpu(), None if binding.normals is None else binding.normals.cpu(),
                               None if delta_in is None else o_leaves["delta"], o_leaves["log_scales"], o_leaves["raw_quats"],
                               o_leaves["opacity_logits"], n_keep=K)
    bound["shs"] = torch.cat([o_leaves["shs"][:K].detach(), o_leaves["shs"][K:]])
    geo = export_geom(P, 96, 96, eager)
    out = oracle.rasterize(bound, to_dev(cam, "cpu"), 3,
                           xy_radii_override=(geo["means2D"].cpu().numpy(), eager[2].cpu().numpy()),
                           depth_override=geo["depths"].cpu().numpy())
    (out["color"] * dL.double()).sum().backward()
    mine = {"log_scales": leaves["log_scales"].grad, "raw_quats": leaves["raw_quats"].grad,
            "opacity_logits": leaves["opacity_logits"].grad.reshape(-1), "shs": leaves["shs"].grad}
    if delta_in is not None:
        assert delta_in.grad.shape == (P, 1)
        mine["delta"] = delta_in.grad.reshape(-1)
    assert leaves["opacity_logits"].grad.shape == (P, 1)
    for k, gm in mine.items():
        assert float(gm[:K].abs().max()) == 0.0, "%s: frozen keep rows must get zero gradient" % k
        e = rel_l2(gm[K:], o_leaves[k].grad[K:])
        assert e <= 3 * GRAD_TOL, "%s rel-L2 vs oracle %g" % (k, e)

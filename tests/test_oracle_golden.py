"""CPU: pins oracle/oracle.py against outputs of the reference's own CUDA rasterizer (tests/golden/*.npz,
generated on a B200 by oracle/make_golden.py).  Integer artefacts bit-exact; image 1e-4; gradients 3e-3."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("train_", "marching_")))   # those: tests/test_train_oracle.py, test_marching_tets.py


def _load(path):
    z = np.load(path)
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    cam = {"image_width": int(z["W"]), "image_height": int(z["H"]), "tanfovx": float(z["cam_tanfovx"]),
           "tanfovy": float(z["cam_tanfovy"]), "viewmatrix": torch.from_numpy(z["cam_viewmatrix"]),
           "projmatrix": torch.from_numpy(z["cam_projmatrix"]), "campos": torch.from_numpy(z["cam_campos"]),
           "bg": torch.from_numpy(z["cam_bg"]), "scale_modifier": 1.0}
    return z, inp, cam


def test_golden_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_integer_pipeline_bit_exact(path):
    """keys / stable sort / ranges / tiles_touched from the reference's own fp32 records: bit-exact."""
    z, inp, cam = _load(path)
    keys, ids, ranges, cnt = oracle.build_keys(z["geom_means2D"], z["radii"], z["geom_depths"], cam["image_width"],
                                               cam["image_height"])
    assert int(cnt.sum()) == int(z["num_rendered"])
    assert np.array_equal(cnt.astype(np.uint32), z["geom_tiles_touched"].astype(np.uint32))
    assert np.array_equal(keys, z["keys"])
    assert np.array_equal(ids, z["ids"])
    assert np.array_equal(ranges, z["ranges"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_preprocess_matches_reference(path):
    z, inp, cam = _load(path)
    W, H = cam["image_width"], cam["image_height"]
    rec = oracle.preprocess(inp["means3D"], inp["opacities"], cam["viewmatrix"], cam["projmatrix"], cam["campos"], W, H,
                            cam["tanfovx"], cam["tanfovy"], scales=inp.get("scales"), rotations=inp.get("rotations"),
                            shs=inp.get("shs"), colors_precomp=inp.get("colors_precomp"), degree=int(z["degree"]))
    vis = z["radii"] > 0
    radius = torch.where(rec["valid"], rec["radius"], torch.zeros_like(rec["radius"])).numpy()
    # float64 vs fp32 may disagree only where 3*sqrt(lambda) sits within rounding of an integer
    assert (radius[vis] != z["radii"][vis]).mean() <= 2e-3
    np.testing.assert_allclose(rec["depth"].numpy()[vis], z["geom_depths"][vis], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(rec["xy"].numpy()[vis], z["geom_means2D"][vis], rtol=0, atol=2e-3)
    co = z["geom_conic_opacity"][vis]
    np.testing.assert_allclose(rec["conic"].numpy()[vis], co[:, :3], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(rec["rgb"].numpy()[vis], z["geom_rgb"][vis] if "shs" in inp else inp["colors_precomp"].numpy()[vis],
                               rtol=0, atol=2e-6)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_image_and_gradients_match_reference(path):
    z, inp, cam = _load(path)
    W, H = cam["image_width"], cam["image_height"]
    oin = {k: v.double().requires_grad_(True) for k, v in inp.items()}
    out = oracle.rasterize(oin, cam, int(z["degree"]), xy_radii_override=(z["geom_means2D"], z["radii"]),
                           depth_override=z["geom_depths"])
    assert out["num_rendered"] == int(z["num_rendered"])
    d = (out["color"].detach().numpy() - z["color"])
    assert np.mean(np.abs(d) > 1e-4) <= 1e-4 and np.abs(d).max() <= 1e-2, "image max-abs %g" % np.abs(d).max()
    nc = out["n_contrib"].numpy().reshape(-1)
    assert np.mean(nc != z["n_contrib"].astype(np.int64)) <= 1e-3
    np.testing.assert_allclose(out["final_T"].numpy().reshape(-1), z["final_T"], rtol=0, atol=2e-4)
    (out["color"] * torch.from_numpy(z["dL_dcolor"]).double()).sum().backward()

    def rel(a, b):
        a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
        n = np.linalg.norm(b)
        return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a - b)

    assert rel(oin["means3D"].grad, z["dL_dmeans3D"]) <= 3e-3
    assert rel(oin["opacities"].grad, z["dL_dopacity"]) <= 3e-3
    assert rel(oin["scales"].grad, z["dL_dscales"]) <= 3e-3
    assert rel(oin["rotations"].grad, z["dL_drotations"]) <= 3e-3
    if "shs" in inp:
        assert rel(oin["shs"].grad, z["dL_dsh"]) <= 3e-3
    else:
        assert rel(oin["colors_precomp"].grad, z["dL_dcolors"]) <= 3e-3


def test_get_higher_msb_matches_reference_table():
    # 32 + bit = 41/43/45/47 sorted bits at 256^2/512^2/1024^2/2048^2 (rasterizer_impl.cu:35-50,300)
    for res, bits in ((256, 41), (512, 43), (1024, 45), (2048, 47)):
        assert 32 + oracle.get_higher_msb((res // 16) ** 2) == bits


def test_analytic_single_gaussian():
    """Known answer: one isotropic Gaussian on the optical axis, alpha = min(0.99, o*exp(-r^2/(2 s^2)))."""
    from youreditableavatar_b200 import scene
    cam = scene.orbit_camera(0, 1, 64, 64, radius=3.0)
    centre = torch.tensor([[0.0, 0.0, -0.05]])
    s = 0.05
    inp = {"means3D": centre, "scales": torch.full((1, 3), s), "rotations": torch.tensor([[1.0, 0, 0, 0]]),
           "opacities": torch.tensor([[0.7]]), "colors_precomp": torch.tensor([[0.2, 0.5, 0.9]])}
    out = oracle.rasterize({k: v.double() for k, v in inp.items()}, cam, 0)
    rec = out["rec"]
    x, y = rec["xy"][0]
    focal = 64 / (2 * cam["tanfovx"])
    sig2 = (focal * s / float(rec["depth"][0])) ** 2 + 0.3
    px, py = int(round(float(x))), int(round(float(y)))
    r2 = (px - float(x)) ** 2 + (py - float(y)) ** 2
    alpha = min(0.99, 0.7 * np.exp(-0.5 * r2 / sig2))
    want = alpha * np.array([0.2, 0.5, 0.9]) + (1 - alpha) * 1.0
    np.testing.assert_allclose(out["color"][:, py, px].numpy(), want, rtol=1e-6)
    assert int(out["radii"][0]) == int(np.ceil(3 * np.sqrt(sig2)))

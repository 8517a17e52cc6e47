"""CPU: on-disk formats, binding rule and keep / edit inheritance — youreditableavatar_b200/formats.py
(device-agnostic torch) against oracle/format_oracle.py (numpy restatement of the reference) and the reference's own
triangle_area / calculate_distances outputs (tests/golden/train_binding.npz).  Index work is compared bit-exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import format_oracle as FO
from youreditableavatar_b200 import formats, scene

G = os.path.join(os.path.dirname(__file__), "golden")


def _mesh(g=12):
    verts, faces, f2t = scene.avatar_mesh(g, "cpu")
    return verts, faces.long(), f2t.long()


def test_oracle_primitives_match_reference_outputs():
    z = np.load(os.path.join(G, "train_binding.npz"))
    A, B, C, pts = (z[k].astype(np.float64) for k in ("A", "B", "C", "points"))
    np.testing.assert_allclose(FO.triangle_area(A, B, C), z["area"].reshape(-1), rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(FO.min_vertex_distance(pts, A, B, C), z["distances"].reshape(-1), rtol=2e-6)
    # the product's primitives (scene.face_areas; the distance inside bind_edit_gaussians) on the same triangles
    verts = torch.from_numpy(np.concatenate([z["A"], z["B"], z["C"]]))
    n = len(z["A"])
    faces = torch.stack([torch.arange(n), torch.arange(n) + n, torch.arange(n) + 2 * n], 1)
    np.testing.assert_allclose(scene.face_areas(verts, faces).numpy(), z["area"].reshape(-1), rtol=2e-5, atol=1e-9)


def test_binding_rule_matches_oracle_bit_exact():
    verts, faces, _ = _mesh()
    fi, bary = scene.bind_faces(verts, faces)
    fo, bo = FO.bind_faces(verts.numpy(), faces.numpy())
    # float32 vs float64 areas may disagree only for a face whose area sits within rounding of the mean
    assert fi.numel() == len(fo) or abs(fi.numel() - len(fo)) <= 4
    if fi.numel() == len(fo):
        assert np.array_equal(fi.numpy(), fo)
        np.testing.assert_allclose(bary.numpy(), bo, rtol=0, atol=1e-7)
    # structure: faces-with-one first (ascending), then faces-with-three, each three times in a row
    area = scene.face_areas(verts, faces)
    n1 = int((area < area.mean()).sum())
    assert torch.equal(fi[:n1], torch.where(area < area.mean())[0])
    rest = fi[n1:].reshape(-1, 3)
    assert torch.equal(rest[:, 0], rest[:, 1]) and torch.equal(rest[:, 0], rest[:, 2])
    assert torch.allclose(bary.sum(-1), torch.ones(fi.numel()))


def test_surface_mesh_npy_round_trip_is_readable_like_the_reference_reads_it(tmp_path):
    verts, faces, f2t = _mesh()
    p = str(tmp_path / "init_mesh.npy")
    formats.save_surface_mesh(p, verts, faces, f2t)
    raw = np.load(p, allow_pickle=True).item()               # tetgs_texture/refine.py:166
    assert set(raw) == {"vertices", "faces", "face_to_global_tet_idx"}
    assert raw["vertices"].shape == (verts.shape[0], 3) and raw["faces"].shape == (faces.shape[0], 3)
    m = formats.load_surface_mesh(p)
    assert torch.equal(m["faces"], faces) and torch.equal(m["face_to_global_tet_idx"], f2t)
    assert torch.equal(m["vertices"], verts) and "keep_faces_num" not in m
    # edit_mesh.npy: mesh_exporter_part.py:174-181
    p2 = str(tmp_path / "edit_mesh.npy")
    kv, kf = verts.shape[0] // 2, faces.shape[0] // 3
    mask = (torch.arange(verts.shape[0]) >= kv).int()
    formats.save_surface_mesh(p2, verts, faces, f2t, keep_vertices_num=kv, keep_faces_num=kf, editing_mask=mask)
    raw = np.load(p2, allow_pickle=True).item()
    assert set(raw) == {"vertices", "faces", "face_to_global_tet_idx", "keep_vertices_num", "keep_faces_num", "editing_mask"}
    m2 = formats.load_surface_mesh(p2)
    assert m2["keep_vertices_num"] == kv and m2["keep_faces_num"] == kf and torch.equal(m2["editing_mask"], mask)
    # the exporter's un-flattened {"mesh": {...}} is tolerated; a file without tet ids is refused
    with open(str(tmp_path / "nested.npy"), "wb") as fh:
        np.save(fh, {"mesh": raw}, allow_pickle=True)
    assert formats.load_surface_mesh(str(tmp_path / "nested.npy"))["keep_faces_num"] == kf
    with open(str(tmp_path / "bad.npy"), "wb") as fh:
        np.save(fh, {"vertices": raw["vertices"], "faces": raw["faces"]}, allow_pickle=True)
    with pytest.raises(KeyError, match="face_to_global_tet_idx"):
        formats.load_surface_mesh(str(tmp_path / "bad.npy"))


def test_split_edit_mesh_matches_oracle():
    verts, faces, f2t = _mesh()
    kv, kf = 40, faces.shape[0] // 2
    faces = faces.clone()
    faces[kf:] = faces[kf:].clamp_min(kv)                    # edit faces only reference edit vertices
    ev, ef = formats.split_edit_mesh({"vertices": verts, "faces": faces, "keep_vertices_num": kv, "keep_faces_num": kf})
    ov, of = FO.split_edit(verts.numpy(), faces.numpy(), kv, kf)
    assert np.array_equal(ev.numpy(), ov) and np.array_equal(ef.numpy(), of) and int(ef.min()) >= 0


def test_checkpoint_round_trip_uses_the_reference_key_names(tmp_path):
    verts, faces, f2t = _mesh()
    fi, _ = scene.bind_faces(verts, faces)
    gs = scene.make_gaussians(verts, faces, fi.numel(), seed=3)         # exactly the rule's Gaussian count
    p = str(tmp_path / "last.pt")
    formats.save_checkpoint(p, gs, f2t, iteration=7000, train_losses=[0.5, 0.4])
    raw = torch.load(p, weights_only=False)
    sd = raw["state_dict"]
    # exactly the parameters a surface-bound reference model registers (tetgs_model.py:100-242): its loader is a
    # strict load_state_dict (:674), so a missing or an extra key fails there
    assert set(sd) == set(formats.REFERENCE_STATE_KEYS)
    assert sd["surface_mesh_thickness"].shape == () and float(sd["surface_mesh_thickness"]) == pytest.approx(1e-6)
    P = fi.numel()
    assert sd["_points"].shape == (P, 1) and sd["all_densities"].shape == (P, 1) and sd["_sh_coordinates_dc"].shape == (P, 1, 3)
    assert sd["_sh_coordinates_rest"].shape == (P, 15, 3) and raw["iteration"] == 7000
    # `points` of the reference = ori_points + normals * _points (tetgs_model.py:252-258) == our activated means
    act = scene.activate(gs)
    assert torch.allclose(sd["ori_points"] + sd["normals"] * sd["_points"], act["means3D"], atol=1e-6)
    gs2, extras = formats.load_checkpoint(p)
    assert extras["train_losses"] == [0.5, 0.4]
    for k in ("delta", "log_scales", "raw_quats", "opacity_logits", "shs", "verts", "faces", "face_index", "bary"):
        assert torch.equal(gs2[k], gs[k]), k
    assert torch.allclose(gs2["vert_normals"], gs["vert_normals"], atol=1e-6)
    assert torch.equal(gs2["face_to_global_tet_idx"], f2t)
    assert gs2["surface_mesh_thickness"] == pytest.approx(1e-6)
    # Gaussians that are not bound by the reference's rule (subsampled here) are refused when SAVING
    sub = {k: (v[::2].contiguous() if k in ("face_index", "bary", "delta", "log_scales", "raw_quats", "opacity_logits", "shs")
               else v) for k, v in gs.items()}
    with pytest.raises(ValueError, match="1-or-3-per-face"):
        formats.save_checkpoint(str(tmp_path / "bad.pt"), sub, f2t)
    # a checkpoint whose Gaussian count contradicts its mesh is refused (the reference would fail in load_state_dict)
    bad = dict(sd)
    bad["_scales"] = sd["_scales"][:-1]
    with pytest.raises(ValueError, match="binds"):
        formats.gaussians_from_state_dict(bad)


def test_keep_inheritance_matches_numpy_restatement_and_concat_feeds_the_rasterizer():
    verts, faces, f2t = _mesh()
    fi, _ = scene.bind_faces(verts, faces)
    gs = scene.make_gaussians(verts, faces, fi.numel(), seed=5)
    g = torch.Generator().manual_seed(1)
    # the edited mesh keeps ~60 % of the old tetrahedra and brings new ones (ids beyond the old range)
    surviving = f2t[torch.rand(f2t.numel(), generator=g) < 0.6]
    edit_f2t = torch.cat([surviving, f2t.max() + 1 + torch.arange(50)])[torch.randperm(surviving.numel() + 50, generator=g)]
    keep = formats.inherit_keep_gaussians(gs, f2t, edit_f2t)
    idx_o = FO.inherit_keep(gs["face_index"].numpy().astype(np.int64), f2t.numpy(), edit_f2t.numpy())
    assert np.array_equal(keep["keep_indices"].numpy(), idx_o) and 0 < len(idx_o) < fi.numel()
    act = scene.activate(gs)
    assert torch.equal(keep["keep_xyz"], act["means3D"][idx_o])
    assert torch.allclose(keep["keep_scales"], gs["log_scales"][idx_o], atol=1e-5)         # log(exp(s)) = s
    assert torch.allclose(keep["keep_rots"].norm(dim=-1), torch.ones(len(idx_o)), atol=1e-6)
    assert keep["keep_face_indices"].dtype == torch.float32 and keep["keep_face_indices"].shape == (len(idx_o), 1)
    assert keep["sh_level"] == 4 and keep["keep_sh_coordinates_rest"].shape == (len(idx_o), 15, 3)
    # fresh Gaussians on an edit sub-mesh + concatenation, keep first (tetgs_edit_2d.py:284-330)
    ev, ef = verts, faces[: faces.shape[0] // 4]
    edit = formats.bind_edit_gaussians(ev, ef)
    fo, _ = FO.bind_faces(ev.numpy(), ef.numpy())
    assert edit["face_index"].numel() == len(fo)
    tri = ev[ef[edit["face_index"].long()]]
    ori = (tri * edit["bary"][..., None]).sum(1)
    d = FO.min_vertex_distance(ori.numpy().astype(np.float64), *(tri[:, k].numpy().astype(np.float64) for k in range(3)))
    np.testing.assert_allclose(torch.exp(edit["log_scales"][:, 1]).numpy(), d, rtol=1e-4)
    np.testing.assert_allclose(torch.exp(edit["log_scales"][:, 0]).numpy(), 1e-8, rtol=1e-5)
    # quaternion frame: first column of R(q) is the face normal
    q = edit["raw_quats"]
    r, x, y, zq = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    col0 = torch.stack([1 - 2 * (y * y + zq * zq), 2 * (x * y + r * zq), 2 * (x * zq - r * y)], -1)
    n = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert torch.allclose(col0, torch.nn.functional.normalize(n, dim=-1), atol=2e-3)
    full = formats.concat_keep_edit(keep, edit)
    K, E = len(idx_o), edit["face_index"].numel()
    assert full["means3D"].shape == (K + E, 3) and full["shs"].shape == (K + E, 16, 3) and full["opacities"].shape == (K + E, 1)
    assert torch.allclose(full["scales"][:K], act["scales"][idx_o], rtol=1e-5)
    assert torch.allclose(full["opacities"][:K], act["opacities"][idx_o], atol=1e-6)
    assert float(full["opacities"][K:].min()) > 0.999


def test_edit_pipeline_end_to_end_renders_through_the_oracle(tmp_path):
    """init_mesh.npy -> bind -> checkpoint -> geometry edit (edit_mesh.npy with kept faces first) -> inherit the kept
    Gaussians by tet id -> bind fresh ones on the edit sub-mesh -> the concatenation is what the rasterizer is fed:
    rendered here with the float64 oracle (CPU).  Kept Gaussians alone reproduce the part of the original image
    they rendered as part of the original model; the edited model covers at least as much of the image."""
    from oracle import oracle
    verts, faces, f2t = _mesh(10)
    formats.save_surface_mesh(str(tmp_path / "init_mesh.npy"), verts, faces, f2t)
    m = formats.load_surface_mesh(str(tmp_path / "init_mesh.npy"))
    fi, _ = scene.bind_faces(m["vertices"], m["faces"])
    gs = scene.make_gaussians(m["vertices"], m["faces"], fi.numel(), seed=2)
    formats.save_checkpoint(str(tmp_path / "last.pt"), gs, m["face_to_global_tet_idx"])
    gs, _ = formats.load_checkpoint(str(tmp_path / "last.pt"))
    # the "edit": faces in the upper part of the body are regenerated (same geometry here, new tet ids), the rest kept
    centre_z = m["vertices"][m["faces"]].mean(1)[:, 2]
    keep_f = centre_z <= 0.2
    kf = int(keep_f.sum())
    order = torch.cat([torch.where(keep_f)[0], torch.where(~keep_f)[0]])
    faces2 = m["faces"][order]
    f2t2 = m["face_to_global_tet_idx"][order].clone()
    f2t2[kf:] += int(f2t2.max()) + 1                        # regenerated region: tetrahedra that did not exist before
    # edit sub-mesh gets its own vertex block, as mesh_exporter_part.py:150-158 concatenates keep + edit vertices
    kv = m["vertices"].shape[0]
    verts2 = torch.cat([m["vertices"], m["vertices"]])
    faces2 = faces2.clone()
    faces2[kf:] += kv
    formats.save_surface_mesh(str(tmp_path / "edit_mesh.npy"), verts2, faces2, f2t2, keep_vertices_num=kv, keep_faces_num=kf)
    em = formats.load_surface_mesh(str(tmp_path / "edit_mesh.npy"))
    keep = formats.inherit_keep_gaussians(gs, gs["face_to_global_tet_idx"], em["face_to_global_tet_idx"])
    n_keep = keep["keep_xyz"].shape[0]
    kept_faces = set(torch.where(keep_f)[0].tolist())
    assert set(keep["keep_face_indices"].long().reshape(-1).tolist()) <= kept_faces and n_keep > 0
    ev, ef = formats.split_edit_mesh(em)
    edit = formats.bind_edit_gaussians(ev, ef)
    full = formats.concat_keep_edit(keep, edit)
    cam = scene.orbit_camera(0, 4, 48, 48, device="cpu")
    out_full = oracle.rasterize({k: v.double() for k, v in full.items()}, cam, degree=3)
    keep_only = {k: v[:n_keep].double() for k, v in full.items()}
    out_keep = oracle.rasterize(keep_only, cam, degree=3)
    act = scene.activate(gs)
    subset = {k: v[keep["keep_indices"]].double() for k, v in act.items()}
    out_subset = oracle.rasterize(subset, cam, degree=3)    # the same Gaussians taken straight from the original model
    assert torch.isfinite(out_full["color"]).all() and out_full["num_rendered"] > out_keep["num_rendered"] > 0
    assert float(out_full["alpha"].sum()) >= float(out_keep["alpha"].sum()) - 1e-9
    # inheritance + concatenation (log/exp, logit/sigmoid, normalised quaternions, fp32 storage) change nothing visible
    assert out_keep["num_rendered"] == out_subset["num_rendered"]
    assert float((out_keep["color"] - out_subset["color"]).abs().max()) < 1e-5

"""On-disk formats of the reference and (re)binding after a geometry edit.

Lets the B200 path consume the artefacts the reference's stages exchange, and re-bind Gaussians after an edit without
leaving the device:

  * surface-mesh `.npy` files — `init_mesh.npy` (Edit_core/tetgs_spatial/models/exporters/mesh_exporter_init.py:56-60)
    and `edit_mesh.npy` (mesh_exporter_part.py:174-181): a pickled dict with `vertices`, `faces`,
    `face_to_global_tet_idx` and, for an edit, `keep_vertices_num`, `keep_faces_num`, `editing_mask`; read the way
    the texture stages do (`np.load(path, allow_pickle=True).item()`, tetgs_texture/refine.py:166);
  * checkpoints — `torch.save({'state_dict': ..., **extras})` (tetgs_scene/tetgs_model.py:635-640) with the
    reference's parameter names (`_points`, `_scales`, `_quaternions`, `all_densities`, `_sh_coordinates_dc`,
    `_sh_coordinates_rest`, `_surface_mesh_faces`, `_verts_points`, `face_to_global_tet_idx`, `ori_points`,
    `normals`), mapped to / from the raw-parameter dict the fused binding kernels take (binding.py, scene.py);
    loading re-derives the binding from the mesh in the checkpoint, as `load_init_model` does (:643-675);
  * keep / edit inheritance by tetrahedron id (`convert_refined_tetgs_into_masked_gaussians`, :680-726): Gaussians on
    faces whose tetrahedron survives the edit are kept with their attributes, done here with device-side `isin`
    instead of the reference's numpy round trip;
  * the edit sub-mesh and its fresh Gaussians (tetgs_scene/tetgs_edit_2d.py:82-136) and the keep + edit
    concatenation the rasterizer is fed (:284-330).

Plain torch / numpy: this is init-time data plumbing (once per stage), not the per-step hot path.
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import scene

__all__ = ["load_surface_mesh", "save_surface_mesh", "split_edit_mesh", "gaussians_to_state_dict",
           "gaussians_from_state_dict", "save_checkpoint", "load_checkpoint", "inherit_keep_gaussians",
           "bind_edit_gaussians", "concat_keep_edit"]

MESH_KEYS = ("vertices", "faces", "face_to_global_tet_idx")
EDIT_KEYS = ("keep_vertices_num", "keep_faces_num", "editing_mask")


# ---------------------------------------------------------------------------------------------- mesh .npy
def load_surface_mesh(path: str, device="cpu") -> Dict[str, object]:
    """`np.load(path, allow_pickle=True).item()` (refine.py:166) -> tensors on `device`: vertices [Nv,3] f32,
    faces [Nf,3] i64, face_to_global_tet_idx [Nf] i64 and, when present, the three edit fields."""
    data = np.load(path, allow_pickle=True).item()
    if "mesh" in data and "vertices" not in data:          # tolerate the exporter's un-flattened params dict
        data = data["mesh"]
    for k in MESH_KEYS:
        if k not in data:
            raise KeyError("surface mesh file %s has no %r" % (path, k))
    as_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    out = {"vertices": torch.from_numpy(as_np(data["vertices"]).astype(np.float32)).to(device),
           "faces": torch.from_numpy(as_np(data["faces"]).astype(np.int64)).to(device),
           "face_to_global_tet_idx": torch.from_numpy(as_np(data["face_to_global_tet_idx"]).astype(np.int64)).reshape(-1).to(device)}
    if out["faces"].shape[0] != out["face_to_global_tet_idx"].shape[0]:
        raise ValueError("face_to_global_tet_idx must hold one tetrahedron id per face")
    if "keep_faces_num" in data:
        out["keep_vertices_num"] = int(data["keep_vertices_num"])
        out["keep_faces_num"] = int(data["keep_faces_num"])
        if data.get("editing_mask") is not None:
            out["editing_mask"] = torch.from_numpy(as_np(data["editing_mask"]).astype(np.int32)).to(device)
    return out


def save_surface_mesh(path: str, vertices, faces, face_to_global_tet_idx, keep_vertices_num: Optional[int] = None,
                      keep_faces_num: Optional[int] = None, editing_mask=None) -> None:
    """Writes what the reference's exporters write (`np.save(path, dict)`, tetgs_spatial/utils/saving.py:554-560):
    numpy vertices / faces / tet ids, plus the edit fields for an `edit_mesh.npy`."""
    as_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    d = {"vertices": as_np(vertices).astype(np.float64), "faces": as_np(faces).astype(np.int64),
         "face_to_global_tet_idx": as_np(face_to_global_tet_idx).astype(np.int64)}
    if keep_faces_num is not None:
        d["keep_vertices_num"], d["keep_faces_num"] = int(keep_vertices_num), int(keep_faces_num)
        d["editing_mask"] = None if editing_mask is None else torch.as_tensor(as_np(editing_mask)).int()
    with open(path, "wb") as fh:                             # np.save would append ".npy" to other suffixes
        np.save(fh, d, allow_pickle=True)


def split_edit_mesh(mesh: Dict[str, object]) -> Tuple[torch.Tensor, torch.Tensor]:
    """tetgs_edit_2d.py:82-99 — the edit sub-mesh of an `edit_mesh.npy`: vertices[keep_vertices_num:] and
    faces[keep_faces_num:] re-indexed from 0."""
    kv, kf = int(mesh["keep_vertices_num"]), int(mesh["keep_faces_num"])
    return mesh["vertices"][kv:].contiguous(), (mesh["faces"][kf:] - kv).contiguous()


# ---------------------------------------------------------------------------------------------- checkpoints
# every key of a surface-bound reference model's state_dict (the nn.Parameters registered at tetgs_model.py:100-242);
# `face_to_global_tet_idx` only exists when the model was built with it (:116-119), `_sh_coordinates_rest` when sh_levels > 1
REFERENCE_STATE_KEYS = ("_surface_mesh_faces", "surface_mesh_thickness", "_verts_points", "face_to_global_tet_idx",
                        "_points", "ori_points", "normals", "all_densities", "_scales", "_quaternions",
                        "_sh_coordinates_dc", "_sh_coordinates_rest")


def check_reference_binding(gs: Dict[str, torch.Tensor]) -> None:
    """The reference re-derives the binding from the mesh when it loads a checkpoint (`load_init_model`,
    tetgs_model.py:643-675: 1 Gaussian per small face, 3 per large one, in face order).  A parameter set whose binding
    is anything else (subsampled, padded, re-ordered) cannot be expressed in that format: refuse to SAVE it rather than
    write a file that only fails when it is loaded."""
    fi, bary = scene.bind_faces(gs["verts"].float(), gs["faces"].long())
    have = gs["face_index"].long().reshape(-1)
    if fi.numel() != have.numel() or not torch.equal(fi.to(have.device).long(), have) or \
            not torch.allclose(bary.to(gs["bary"].device), gs["bary"].float(), atol=1e-6):
        raise ValueError("these Gaussians are not bound by the reference's 1-or-3-per-face rule (%d Gaussians, the rule "
                         "binds %d to this mesh): the reference checkpoint format cannot hold them"
                         % (have.numel(), fi.numel()))


def gaussians_to_state_dict(gs: Dict[str, torch.Tensor], face_to_global_tet_idx: Optional[torch.Tensor] = None,
                            surface_mesh_thickness: Optional[float] = None, spatial_extent: float = 1.0
                            ) -> Dict[str, torch.Tensor]:
    """Raw-parameter dict (scene.make_gaussians / binding.MeshBinding layout) -> the reference's state-dict names and
    shapes (tetgs_model.py:100-242): delta [P,1] as `_points`, log-scales, un-normalised quaternions, opacity logits
    [P,1], SH split into dc [P,1,3] / rest [P,M-1,3], the mesh, the derived `ori_points` / `normals`, and the scalar
    `surface_mesh_thickness` (default: spatial extent / 1e6, tetgs_model.py:105-106) — `load_init_model` loads the
    dict with a strict `load_state_dict` (:674), so every registered parameter has to be there."""
    faces = gs["faces"].long()
    fi = gs["face_index"].long()
    w = gs["bary"][..., None]
    if surface_mesh_thickness is None:
        surface_mesh_thickness = float(spatial_extent) / 1_000_000
    sd = {
        "surface_mesh_thickness": torch.tensor(float(surface_mesh_thickness), device=gs["verts"].device),
        "_points": gs["delta"].reshape(-1, 1).clone(),
        "_scales": gs["log_scales"].clone(),
        "_quaternions": gs["raw_quats"].clone(),
        "all_densities": gs["opacity_logits"].reshape(-1, 1).clone(),
        "_sh_coordinates_dc": gs["shs"][:, :1].clone(),
        "_surface_mesh_faces": faces.clone(),
        "_verts_points": gs["verts"].clone(),
        "ori_points": (gs["verts"][faces[fi]] * w).sum(1),
        "normals": (gs["vert_normals"][faces[fi]] * w).sum(1),
    }
    if gs["shs"].shape[1] > 1:
        sd["_sh_coordinates_rest"] = gs["shs"][:, 1:].clone()
    if face_to_global_tet_idx is not None:
        sd["face_to_global_tet_idx"] = face_to_global_tet_idx.clone()
    return sd


def gaussians_from_state_dict(sd: Dict[str, torch.Tensor], device=None) -> Dict[str, torch.Tensor]:
    """Inverse of `gaussians_to_state_dict`, the way `load_init_model` rebuilds the model (tetgs_model.py:643-675):
    the binding (face per Gaussian, barycentric weights, vertex normals) is re-derived from the mesh in the checkpoint
    with the reference's rule, then the learned tensors are adopted.  Raises if the checkpoint's Gaussian count does
    not match what the rule yields for its mesh."""
    dev = device if device is not None else sd["_verts_points"].device
    t = lambda k: sd[k].to(dev)
    verts = t("_verts_points").float()
    if verts.shape[1] != 3:                                   # tetgs_model.py:653-656
        verts = verts.repeat(1, 3)
    faces = t("_surface_mesh_faces").long()
    face_index, bary = scene.bind_faces(verts, faces)
    P = sd["_scales"].shape[0]
    if face_index.numel() != P:
        raise ValueError("checkpoint holds %d Gaussians but its mesh binds %d" % (P, face_index.numel()))
    shs = t("_sh_coordinates_dc").float()
    if "_sh_coordinates_rest" in sd:
        shs = torch.cat([shs, t("_sh_coordinates_rest").float()], dim=1)
    gs = {
        "verts": verts.contiguous(), "faces": faces.to(torch.int32).contiguous(),
        "vert_normals": scene.vertex_normals(verts, faces).contiguous(),
        "face_index": face_index.to(torch.int32).contiguous(), "bary": bary.contiguous(),
        "delta": t("_points").float().reshape(-1).contiguous(), "log_scales": t("_scales").float().contiguous(),
        "raw_quats": t("_quaternions").float().contiguous(),
        "opacity_logits": t("all_densities").float().reshape(-1).contiguous(), "shs": shs.contiguous(),
    }
    if "face_to_global_tet_idx" in sd:
        gs["face_to_global_tet_idx"] = t("face_to_global_tet_idx").long().reshape(-1)
    if "surface_mesh_thickness" in sd:
        gs["surface_mesh_thickness"] = float(sd["surface_mesh_thickness"])
    return gs


def save_checkpoint(path: str, gs: Dict[str, torch.Tensor], face_to_global_tet_idx: Optional[torch.Tensor] = None,
                    surface_mesh_thickness: Optional[float] = None, spatial_extent: float = 1.0, **extras) -> None:
    """tetgs_model.py:635-640 `save_model`: {'state_dict': ..., **kwargs} (the loops pass train_losses, epoch,
    iteration, optimizer_state_dict — refine.py:345-373).  Raises ValueError when `gs` is not bound by the reference's
    rule (see `check_reference_binding`)."""
    check_reference_binding(gs)
    ckpt = {"state_dict": gaussians_to_state_dict(gs, face_to_global_tet_idx, surface_mesh_thickness, spatial_extent)}
    ckpt.update(extras)
    torch.save(ckpt, path)


def load_checkpoint(path: str, device="cpu") -> Tuple[Dict[str, torch.Tensor], Dict[str, object]]:
    """-> (raw-parameter dict, the checkpoint's other entries)."""
    ckpt = torch.load(path, map_location=device, weights_only=False)
    gs = gaussians_from_state_dict(ckpt["state_dict"], device)
    return gs, {k: v for k, v in ckpt.items() if k != "state_dict"}


# ---------------------------------------------------------------------------------------------- keep / edit
def inherit_keep_gaussians(gs: Dict[str, torch.Tensor], face_to_global_tet_idx: torch.Tensor,
                           edit_face_to_global_tet_idx: torch.Tensor) -> Dict[str, object]:
    """`convert_refined_tetgs_into_masked_gaussians` (tetgs_model.py:680-726) on the device: a face is inherited when
    its tetrahedron id also occurs in the edited mesh, a Gaussian when its face is; returns the reference's
    `keep_*` dict (world positions, opacity logits, log-scales = log(exp(_scales)), NORMALISED quaternions, SH dc /
    rest, face indices as float [K,1], sh_level) plus `keep_indices` (which Gaussians were kept)."""
    act = scene.activate(gs)
    dev = act["means3D"].device
    f2t = face_to_global_tet_idx.to(dev).long().reshape(-1)
    edit = edit_face_to_global_tet_idx.to(dev).long().reshape(-1)
    face_mask = torch.isin(f2t, edit)                                         # np.isin(face_to_global_tet_idx, edit…)
    fi = gs["face_index"].to(dev).long().reshape(-1)
    idx = torch.nonzero(face_mask[fi]).reshape(-1)                            # np.isin(face_indices, where(face_mask))
    M = gs["shs"].shape[1]
    sh_level = int(round(M ** 0.5))
    keep = {
        "keep_xyz": act["means3D"][idx].float(),
        "keep_opacities": gs["opacity_logits"].reshape(-1, 1)[idx].float(),
        "keep_scales": torch.log(act["scales"][idx]).float(),                 # scale_inverse_activation(scaling)
        "keep_rots": act["rotations"][idx].float(),                           # the `quaternions` property normalises
        "keep_sh_coordinates_dc": gs["shs"][idx, :1].float(),
        "keep_face_indices": fi[idx].reshape(-1, 1).float(),
        "sh_level": sh_level,
        "keep_indices": idx,
    }
    if sh_level > 1:
        keep["keep_sh_coordinates_rest"] = gs["shs"][idx, 1:].float()
    return keep


def bind_edit_gaussians(edit_vertices: torch.Tensor, edit_faces: torch.Tensor, sh_coeffs: int = 16,
                        opacity: float = 0.9999) -> Dict[str, torch.Tensor]:
    """Fresh Gaussians on the edit sub-mesh (tetgs_edit_2d.py:100-221): the 1-or-3-per-face rule, flat Gaussians in the
    triangle frame — quaternion from (normal, first edge, their cross product), scales (1e-8, d, d) with d the distance
    to the nearest triangle vertex (:172-208) — grey colour 0.5 (:104-107), opacity inverse_sigmoid(0.9999) (the
    bound models' default, tetgs_model.py:197-199), zero offset.  Returns the raw-parameter dict."""
    verts, faces = edit_vertices.float(), edit_faces.long()
    face_index, bary = scene.bind_faces(verts, faces)
    P = face_index.numel()
    dev = verts.device
    tri = verts[faces[face_index]]
    ori = (tri * bary[..., None]).sum(1)
    eps = 1e-8
    n = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    v0 = n / (n.norm(dim=-1, keepdim=True) + eps)
    v1 = tri[:, 1] - tri[:, 0]
    v1 = v1 / (v1.norm(dim=-1, keepdim=True) + eps)
    v2 = torch.linalg.cross(v0, v1)
    v2 = v2 / (v2.norm(dim=-1, keepdim=True) + eps)
    quats = scene._matrix_to_quaternion(torch.stack([v0, v1, v2], dim=-1))
    d = (tri - ori[:, None, :]).norm(dim=-1).min(dim=-1)[0].clamp_min(1e-7)
    log_scales = torch.stack([torch.full_like(d, 1e-8).log(), d.log(), d.log()], -1)
    shs = torch.zeros(P, sh_coeffs, 3, device=dev)
    shs[:, 0] = (0.5 - 0.5) / scene.C0                                        # RGB2SH(grey 0.5)
    logit = float(np.log(opacity / (1.0 - opacity)))
    return {"verts": verts.contiguous(), "faces": faces.to(torch.int32).contiguous(),
            "vert_normals": scene.vertex_normals(verts, faces).contiguous(),
            "face_index": face_index.to(torch.int32).contiguous(), "bary": bary.contiguous(),
            "delta": torch.zeros(P, device=dev), "log_scales": log_scales.contiguous(), "raw_quats": quats.contiguous(),
            "opacity_logits": torch.full((P,), logit, device=dev), "shs": shs}


def concat_keep_edit(keep: Dict[str, object], edit_gs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """What the edit models hand the rasterizer (tetgs_edit_2d.py:284-330): keep Gaussians first, then the edit
    Gaussians, activated — means3D, scales = exp, rotations = normalize, opacities = sigmoid, SH [P, M, 3]."""
    e = scene.activate(edit_gs)
    M = e["shs"].shape[1]
    ksh = keep["keep_sh_coordinates_dc"]
    if "keep_sh_coordinates_rest" in keep:
        ksh = torch.cat([ksh, keep["keep_sh_coordinates_rest"]], dim=1)
    if ksh.shape[1] < M:                                                      # lower SH level kept: pad with zeros
        ksh = torch.cat([ksh, ksh.new_zeros(ksh.shape[0], M - ksh.shape[1], 3)], dim=1)
    return {
        "means3D": torch.cat([keep["keep_xyz"], e["means3D"]]).contiguous(),
        "scales": torch.cat([torch.exp(keep["keep_scales"]), e["scales"]]).contiguous(),
        "rotations": torch.cat([torch.nn.functional.normalize(keep["keep_rots"], dim=-1), e["rotations"]]).contiguous(),
        "opacities": torch.cat([torch.sigmoid(keep["keep_opacities"]), e["opacities"]]).contiguous(),
        "shs": torch.cat([ksh[:, :M], e["shs"]]).contiguous(),
    }

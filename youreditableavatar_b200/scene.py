"""Deterministic synthetic TetGS scenes: tet grid -> avatar shell by marching tetrahedra ->
mesh-bound Gaussians -> orbit cameras.  Used by tests, bench.py and smoke(); device-agnostic torch code.

What is mirrored from the reference (behaviour only):
  * tet-grid npz schema `vertices [Nv,3] f32`, `indices [Nt,4] i64` (tetgs_spatial/models/isosurface.py:61-71);
  * marching tetrahedra rule incl. the face -> tet map (isosurface.py:112-184);
  * binding: 1 Gaussian at (1/3,1/3,1/3) if face area < mean area else 3 at permutations of (2/3,1/6,1/6),
    faces-with-1 first (tetgs_scene/tetgs_model.py:328-377); mean = ori + normal*delta (:252-258);
  * flat-Gaussian frame quaternion / (eps,d,d)-style scales (tetgs_scene/tetgs_edit_2d.py:172-208);
  * orbit cameras (tetgs_scene/cameras.py:281-345) and the transposed view / projection matrices fed to the
    rasterizer (tetgs_model.py:479-521, utils/graphics_utils.py:39-86).
"""
import math
from typing import Dict, Optional, Tuple

import torch

C0 = 0.28209479177387814

# ---------------------------------------------------------------------------------------------
# tet grid
# ---------------------------------------------------------------------------------------------
_KUHN = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def _corner_offsets_of_kuhn_tets() -> torch.Tensor:
    """[6,4,3] integer corner offsets of the 6 Kuhn tetrahedra that share the cube diagonal 000-111."""
    tets = []
    for perm in _KUHN:
        c = [0, 0, 0]
        path = [tuple(c)]
        for ax in perm:
            c[ax] = 1
            path.append(tuple(c))
        tets.append(path)
    return torch.tensor(tets, dtype=torch.int64)


def make_tet_grid(g: int, device="cpu") -> Dict[str, torch.Tensor]:
    """Full (g+1)^3 lattice over [-1,1]^3, 6 Kuhn tets per cube — the `{res}_tets.npz` schema."""
    lin = torch.linspace(-1.0, 1.0, g + 1, device=device, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(lin, lin, lin, indexing="ij")
    vertices = torch.stack([X, Y, Z], -1).reshape(-1, 3)
    r = torch.arange(g, device=device)
    cx, cy, cz = torch.meshgrid(r, r, r, indexing="ij")
    cubes = torch.stack([cx, cy, cz], -1).reshape(-1, 3)
    indices = _tets_of_cubes(cubes, g)
    return {"vertices": vertices, "indices": indices}


def _tets_of_cubes(cubes: torch.Tensor, g: int) -> torch.Tensor:
    off = _corner_offsets_of_kuhn_tets().to(cubes.device)  # [6,4,3]
    c = cubes[:, None, None, :] + off[None]                # [Nc,6,4,3]
    n = g + 1
    return ((c[..., 0] * n + c[..., 1]) * n + c[..., 2]).reshape(-1, 4)


# ---------------------------------------------------------------------------------------------
# avatar shell: inside-positive field, 1-Lipschitz (max of r - distance-to-segment)
# ---------------------------------------------------------------------------------------------
_CAPSULES = [
    # (ax, ay, az, bx, by, bz, radius)
    (0.0, 0.0, 0.66, 0.0, 0.0, 0.70, 0.125),      # head
    (0.0, 0.0, 0.50, 0.0, 0.0, 0.58, 0.055),      # neck
    (0.0, 0.0, 0.05, 0.0, 0.0, 0.38, 0.150),      # torso
    (-0.10, 0.0, 0.42, 0.10, 0.0, 0.42, 0.085),   # shoulders
    (-0.06, 0.0, 0.00, 0.06, 0.0, 0.00, 0.120),   # hips
    (-0.085, 0.0, -0.05, -0.11, 0.0, -0.78, 0.070),  # legs
    (0.085, 0.0, -0.05, 0.11, 0.0, -0.78, 0.070),
    (-0.19, 0.0, 0.42, -0.36, 0.02, 0.02, 0.048),    # arms
    (0.19, 0.0, 0.42, 0.36, 0.02, 0.02, 0.048),
]


def avatar_field(p: torch.Tensor) -> torch.Tensor:
    """Occupancy-style level function: > 0 inside the synthetic avatar (isosurface.py:114 uses sdf > 0)."""
    best = torch.full(p.shape[:-1], -1e9, dtype=p.dtype, device=p.device)
    for ax, ay, az, bx, by, bz, r in _CAPSULES:
        a = torch.tensor([ax, ay, az], dtype=p.dtype, device=p.device)
        b = torch.tensor([bx, by, bz], dtype=p.dtype, device=p.device)
        ab = b - a
        t = ((p - a) @ ab / (ab @ ab)).clamp(0.0, 1.0)
        d = (p - (a + t[..., None] * ab)).norm(dim=-1)
        best = torch.maximum(best, r - d)
    return best


def surface_cubes(g: int, device="cpu", coarse: int = 8) -> torch.Tensor:
    """Integer coords [Nc,3] of the fine cubes that can contain the level set (two-level search)."""
    if g <= 64:
        r = torch.arange(g, device=device)
        cx, cy, cz = torch.meshgrid(r, r, r, indexing="ij")
        return torch.stack([cx, cy, cz], -1).reshape(-1, 3)
    gc = (g + coarse - 1) // coarse
    h = 2.0 / g
    r = torch.arange(gc, device=device)
    cx, cy, cz = torch.meshgrid(r, r, r, indexing="ij")
    cc = torch.stack([cx, cy, cz], -1).reshape(-1, 3)
    centre = -1.0 + (cc.to(torch.float32) + 0.5) * (coarse * h)
    f = avatar_field(centre)
    keep = f.abs() <= (coarse * h) * (math.sqrt(3.0) / 2.0) * 1.001
    cc = cc[keep]
    r8 = torch.arange(coarse, device=device)
    ox, oy, oz = torch.meshgrid(r8, r8, r8, indexing="ij")
    o = torch.stack([ox, oy, oz], -1).reshape(-1, 3)
    fine = (cc[:, None, :] * coarse + o[None]).reshape(-1, 3)
    fine = fine[(fine < g).all(-1)]
    # second filter on fine cube centres
    centre = -1.0 + (fine.to(torch.float32) + 0.5) * h
    f = avatar_field(centre)
    return fine[f.abs() <= h * (math.sqrt(3.0) / 2.0) * 1.001]


# ---------------------------------------------------------------------------------------------
# marching tetrahedra (isosurface.py:112-184)
# ---------------------------------------------------------------------------------------------
_TRI_TABLE = torch.tensor(
    [[-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
     [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
     [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
     [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]],
    dtype=torch.int64)
_NUM_TRI = torch.tensor([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=torch.int64)
_TET_EDGES = torch.tensor([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=torch.int64)


def marching_tets(pos: torch.Tensor, level: torch.Tensor, tets: torch.Tensor):
    """pos [Nv,3], level [Nv] (>0 inside), tets [Nt,4] -> verts [V,3], faces [F,3], face_to_tet [F]."""
    dev = pos.device
    occ = level > 0
    occ4 = occ[tets.reshape(-1)].reshape(-1, 4)
    s = occ4.sum(-1)
    valid = (s > 0) & (s < 4)
    vt = tets[valid]
    edges = vt[:, _TET_EDGES.to(dev)].reshape(-1, 2)
    edges = torch.sort(edges, dim=1)[0]
    uniq, inv = torch.unique(edges, dim=0, return_inverse=True)
    crossing = occ[uniq.reshape(-1)].reshape(-1, 2).sum(-1) == 1
    mapping = torch.full((uniq.shape[0],), -1, dtype=torch.int64, device=dev)
    mapping[crossing] = torch.arange(int(crossing.sum()), device=dev)
    idx_map = mapping[inv].reshape(-1, 6)
    ev = uniq[crossing]
    p2 = pos[ev.reshape(-1)].reshape(-1, 2, 3)
    l2 = level[ev.reshape(-1)].reshape(-1, 2, 1).clone()
    l2[:, -1] *= -1
    w = torch.flip(l2, [1]) / l2.sum(1, keepdim=True)
    verts = (p2 * w).sum(1)
    code = (occ4[valid].to(torch.int64) * torch.tensor([1, 2, 4, 8], device=dev)).sum(-1)
    ntri = _NUM_TRI.to(dev)[code]
    table = _TRI_TABLE.to(dev)
    one, two = ntri == 1, ntri == 2
    faces = torch.cat([
        torch.gather(idx_map[one], 1, table[code[one]][:, :3]).reshape(-1, 3),
        torch.gather(idx_map[two], 1, table[code[two]][:, :6]).reshape(-1, 3),
    ], 0)
    vidx = torch.where(valid)[0]
    face_to_tet = torch.cat([vidx[one], vidx[two].repeat_interleave(2)])
    return verts, faces, face_to_tet


def avatar_mesh(g: int, device="cpu"):
    """Marching-tets surface of the synthetic avatar on the (g+1)^3 Kuhn grid; only surface cubes are built."""
    cubes = surface_cubes(g, device)
    tets_global = _tets_of_cubes(cubes, g)
    ids, inv = torch.unique(tets_global.reshape(-1), return_inverse=True)
    n = g + 1
    iz = ids % n
    iy = (ids // n) % n
    ix = ids // (n * n)
    pos = torch.stack([ix, iy, iz], -1).to(torch.float32) * (2.0 / g) - 1.0
    level = avatar_field(pos)
    tets = inv.reshape(-1, 4)
    verts, faces, f2t_local = marching_tets(pos, level, tets)
    cube_lin = (cubes[:, 0] * g + cubes[:, 1]) * g + cubes[:, 2]
    tet_global_id = (cube_lin[:, None] * 6 + torch.arange(6, device=device)[None]).reshape(-1)
    return verts, faces, tet_global_id[f2t_local]


# ---------------------------------------------------------------------------------------------
# binding (tetgs_model.py:328-377) and per-Gaussian parameters
# ---------------------------------------------------------------------------------------------
def face_areas(verts, faces):
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    return 0.5 * torch.linalg.cross(b - a, c - a).norm(dim=-1)


def vertex_normals(verts, faces):
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    fn = torch.linalg.cross(b - a, c - a)  # area-weighted
    vn = torch.zeros_like(verts)
    for k in range(3):
        vn.index_add_(0, faces[:, k], fn)
    return torch.nn.functional.normalize(vn, dim=-1, eps=1e-6)


def bind_faces(verts, faces):
    """Reference rule: returns face_index [P0] and bary [P0,3], faces-with-1 first then faces-with-3."""
    area = face_areas(verts, faces)
    one = area < area.mean()
    dev = verts.device
    f1 = torch.where(one)[0]
    f3 = torch.where(~one)[0]
    b1 = torch.tensor([[1 / 3, 1 / 3, 1 / 3]], dtype=torch.float32, device=dev).expand(f1.numel(), 3)
    b3 = torch.tensor([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]], dtype=torch.float32,
                      device=dev).repeat(f3.numel(), 1)
    face_index = torch.cat([f1, f3.repeat_interleave(3)])
    bary = torch.cat([b1, b3], 0)
    return face_index, bary


def _matrix_to_quaternion(R: torch.Tensor) -> torch.Tensor:
    """Rotation matrices [N,3,3] -> quaternions (r,x,y,z), real part non-negative."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    q_abs = torch.sqrt(torch.clamp(torch.stack([
        1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1), min=0.0))
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    q = cand[torch.arange(R.shape[0], device=R.device), best]
    q = torch.where(q[:, :1] < 0, -q, q)
    return torch.nn.functional.normalize(q, dim=-1)


def make_gaussians(verts, faces, P: int, seed: int = 0, sh_coeffs: int = 16) -> Dict[str, torch.Tensor]:
    """Binds exactly P Gaussians to the mesh and draws their raw (pre-activation) parameters."""
    dev = verts.device
    gen = torch.Generator(device="cpu").manual_seed(seed)
    face_index, bary = bind_faces(verts, faces)
    P0 = face_index.numel()
    if P0 >= P:
        keep = torch.randperm(P0, generator=gen)[:P].sort()[0].to(dev)
    else:
        extra = torch.randint(0, P0, (P - P0,), generator=gen).to(dev)
        keep = torch.cat([torch.arange(P0, device=dev), extra])
    face_index, bary = face_index[keep].contiguous(), bary[keep].contiguous()

    tri = verts[faces[face_index]]                      # [P,3,3]
    vn = vertex_normals(verts, faces)
    ori = (tri * bary[..., None]).sum(1)
    # triangle frame: normal, first edge, their cross product (tetgs_edit_2d.py:172-197)
    eps = 1e-8
    n = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    v0 = n / (n.norm(dim=-1, keepdim=True) + eps)
    v1 = tri[:, 1] - tri[:, 0]
    v1 = v1 / (v1.norm(dim=-1, keepdim=True) + eps)
    v2 = torch.linalg.cross(v0, v1)
    v2 = v2 / (v2.norm(dim=-1, keepdim=True) + eps)
    R = torch.stack([v0, v1, v2], dim=-1)               # columns
    quats = _matrix_to_quaternion(R)
    # tangential scale = distance to the nearest triangle vertex, jittered; flat along the normal
    d = (tri - ori[:, None, :]).norm(dim=-1).min(dim=-1)[0].clamp_min(1e-7)
    jitter = torch.randn(P, generator=gen).to(dev) * 0.3
    log_t = torch.log(d) + jitter
    log_scales = torch.stack([log_t + math.log(0.1), log_t, log_t], -1)
    opacity_logits = (2.0 + 1.5 * torch.randn(P, generator=gen)).to(dev)
    rgb = torch.rand(P, 3, generator=gen).to(dev)
    sh = torch.zeros(P, sh_coeffs, 3, device=dev)
    sh[:, 0] = (rgb - 0.5) / C0
    if sh_coeffs > 1:
        sh[:, 1:] = (0.05 * torch.randn(P, sh_coeffs - 1, 3, generator=gen)).to(dev)
    delta = (1e-3 * torch.randn(P, generator=gen)).to(dev)
    # raw quaternions are un-normalised parameters: scale them a little so normalisation matters
    raw_quats = quats * (1.0 + 0.1 * torch.rand(P, 1, generator=gen).to(dev))
    return {
        "verts": verts.contiguous(), "faces": faces.to(torch.int32).contiguous(), "vert_normals": vn.contiguous(),
        "face_index": face_index.to(torch.int32).contiguous(), "bary": bary, "delta": delta,
        "log_scales": log_scales.contiguous(), "raw_quats": raw_quats.contiguous(),
        "opacity_logits": opacity_logits, "shs": sh.contiguous(),
    }


def activate(gs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Eager restatement of the binding + activations (tetgs_model.py:252-286): what the drop-in API is fed."""
    verts, faces = gs["verts"], gs["faces"].long()
    fi = gs["face_index"].long()
    tri = verts[faces[fi]]
    nrm = gs["vert_normals"][faces[fi]]
    w = gs["bary"][..., None]
    ori = (tri * w).sum(1)
    n = (nrm * w).sum(1)
    means = ori + n * gs["delta"][:, None]
    return {
        "means3D": means.contiguous(),
        "scales": torch.exp(gs["log_scales"]),
        "rotations": torch.nn.functional.normalize(gs["raw_quats"], dim=-1),
        "opacities": torch.sigmoid(gs["opacity_logits"])[:, None].contiguous(),
        "shs": gs["shs"],
    }


# ---------------------------------------------------------------------------------------------
# cameras (cameras.py:281-345; tetgs_model.py:479-521; graphics_utils.py:39-86)
# ---------------------------------------------------------------------------------------------
def projection_matrix(znear, zfar, fovx, fovy) -> torch.Tensor:
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4, dtype=torch.float64)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def orbit_camera(k: int, V: int, H: int, W: int, radius: float = 3.0, fovy_deg: float = 45.0,
                 focal_scale: float = 1.4, device="cpu") -> Dict[str, object]:
    """View k of a V-view orbit: azimuth 360k/V, elevation cycling [5,-15,25] (paint_2dgs.py:161-166)."""
    az = math.radians(360.0 * k / max(V, 1))
    el = math.radians([5.0, -15.0, 25.0][k % 3])
    cam = torch.tensor([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az),
                        radius * math.sin(el)], dtype=torch.float64)
    centre = torch.tensor([0.0, 0.0, -0.05], dtype=torch.float64)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    look = torch.nn.functional.normalize(centre - cam, dim=0)
    right = torch.nn.functional.normalize(torch.linalg.cross(look, up), dim=0)
    upv = torch.nn.functional.normalize(torch.linalg.cross(right, look), dim=0)
    c2w = torch.eye(4, dtype=torch.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, upv, -look, cam
    c2w[:3, 1:3] *= -1            # OpenGL (y up, z back) -> COLMAP (y down, z forward)
    w2c = torch.linalg.inv(c2w)
    focal = focal_scale * 0.5 * H / math.tan(0.5 * math.radians(fovy_deg))
    fovy = 2 * math.atan(H / (2 * focal))
    fovx = 2 * math.atan(W / (2 * focal))
    view = w2c.transpose(0, 1)
    proj = projection_matrix(1e-4, 100.0, fovx, fovy).transpose(0, 1)
    full = view @ proj
    f32 = dict(dtype=torch.float32, device=device)
    return {
        "image_height": H, "image_width": W, "tanfovx": math.tan(fovx * 0.5), "tanfovy": math.tan(fovy * 0.5),
        "viewmatrix": view.to(**f32).contiguous(), "projmatrix": full.to(**f32).contiguous(),
        "campos": cam.to(**f32).contiguous(), "bg": torch.ones(3, **f32), "scale_modifier": 1.0,
    }


# ---------------------------------------------------------------------------------------------
# named configs (BASELINE.json `configs`)
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (P, H=W, views, tet-grid resolution g)
    "C1": (10_000, 256, 1, 32),
    "C2": (300_000, 512, 4, 208),
    "C3": (1_000_000, 1024, 1, 376),
    "C4": (1_000_000, 1024, 64, 376),
    "C5": (4_000_000, 2048, 8, 752),
}

_scene_cache: Dict[Tuple, Dict[str, torch.Tensor]] = {}


def make_scene(name_or_P, res: Optional[int] = None, g: Optional[int] = None, seed: int = 0, device="cpu",
               sh_coeffs: int = 16) -> Dict[str, torch.Tensor]:
    """Raw mesh-bound Gaussian parameters for a named config ("C1".."C5") or an explicit (P, g)."""
    if isinstance(name_or_P, str):
        P, res, _views, g0 = CONFIGS[name_or_P]
        g = g or g0
    else:
        P = int(name_or_P)
        assert g is not None
    key = (P, g, seed, str(device), sh_coeffs)
    if key not in _scene_cache:
        verts, faces, f2t = avatar_mesh(g, device)
        gs = make_gaussians(verts, faces, P, seed, sh_coeffs)
        gs["face_to_global_tet_idx"] = f2t
        gs["grid_res"] = g
        _scene_cache[key] = gs
    return _scene_cache[key]

"""Generates tests/golden/marching_tets.npz by running the reference's OWN marching tetrahedra —
`MarchingTetrahedraHelper._forward`, /root/reference/Edit_core/tetgs_spatial/models/isosurface.py:112-184 — imported from
where it lies, on CPU, on two small tet grids (a sphere and the synthetic avatar field with noise).

isosurface.py imports its package (`tetgs_spatial`, `.models.mesh`, `.utils.typing`), which pulls in libraries that are
absent here and that `_forward` never touches: those three modules are stubbed (typing names only).
Run in the build container (needs /root/reference):  python oracle/make_golden_mt.py
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = "/root/reference/Edit_core/tetgs_spatial/models/isosurface.py"


class _Sub:
    def __getitem__(self, item):
        return object


def load_reference_class():
    import typing
    pkg = types.ModuleType("tetgs_spatial"); pkg.__path__ = []
    models = types.ModuleType("tetgs_spatial.models"); models.__path__ = []
    mesh = types.ModuleType("tetgs_spatial.models.mesh"); mesh.Mesh = object
    utils = types.ModuleType("tetgs_spatial.utils"); utils.__path__ = []
    typ = types.ModuleType("tetgs_spatial.utils.typing")
    for n in ("Tuple", "Optional", "Union", "List", "Dict", "Any", "Callable"):
        setattr(typ, n, getattr(typing, n))
    for n in ("Float", "Integer", "Int", "Bool", "Num"):
        setattr(typ, n, _Sub())
    typ.Tensor = torch.Tensor
    typ.__all__ = [n for n in dir(typ) if not n.startswith("_")]
    for name, m in [("tetgs_spatial", pkg), ("tetgs_spatial.models", models), ("tetgs_spatial.models.mesh", mesh),
                    ("tetgs_spatial.utils", utils), ("tetgs_spatial.utils.typing", typ)]:
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_isosurface", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MarchingTetrahedraHelper


def main():
    from youreditableavatar_b200 import scene
    Helper = load_reference_class()
    out = {}
    g = torch.Generator().manual_seed(0)
    for name, res in (("sphere", 10), ("avatar", 24)):
        grid = scene.make_tet_grid(res)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "tets.npz")
            np.savez(path, vertices=grid["vertices"].numpy(), indices=grid["indices"].numpy())
            helper = Helper(res, path)
        pos = helper.grid_vertices
        if name == "sphere":
            sdf = 0.62 - pos.norm(dim=-1) + 0.03 * torch.randn(pos.shape[0], generator=g)
        else:
            sdf = scene.avatar_field(pos * torch.tensor([1.0, 1.0, 1.0])) + 0.004 * torch.randn(pos.shape[0], generator=g)
        r = helper._forward(pos, sdf.clone(), helper.indices)
        out[name + "_pos"] = pos.numpy()
        out[name + "_sdf"] = sdf.numpy()
        out[name + "_tets"] = helper.indices.numpy()
        for k in ("verts", "faces", "face_to_tet_idx", "valid_tets", "interp_v"):
            out[name + "_" + k] = r[k].numpy()
        print(name, "verts", tuple(r["verts"].shape), "faces", tuple(r["faces"].shape))
    out["cases"] = np.array(["sphere", "avatar"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "marching_tets.npz"), **out)


if __name__ == "__main__":
    main()

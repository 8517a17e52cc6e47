"""CPU: synthetic scene generator (determinism, reference binding rule, camera conventions) and the
data-parallel host logic (view sharding, flat gradient bucket, gloo all-reduce with world_size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from youreditableavatar_b200 import scene
from youreditableavatar_b200.parallel import GRAD_NAMES, GradBucket, grad_shapes, shard_views


def test_tet_grid_schema_and_volume():
    g = scene.make_tet_grid(4)
    v, t = g["vertices"], g["indices"]
    assert v.shape == (125, 3) and v.dtype == torch.float32 and t.shape == (6 * 64, 4) and t.dtype == torch.int64
    p = v[t].double()
    vol = torch.linalg.det(p[:, 1:] - p[:, :1]).abs() / 6
    assert torch.allclose(vol.sum(), torch.tensor(8.0, dtype=torch.float64))   # tets tile [-1,1]^3 exactly


def test_marching_tets_closed_surface_and_face_to_tet():
    verts, faces, f2t = scene.avatar_mesh(32)
    assert faces.shape[0] == f2t.shape[0] and faces.max() < verts.shape[0]
    # watertight: every undirected edge is shared by exactly two faces
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).sort(dim=1)[0]
    _, counts = torch.unique(e, dim=0, return_counts=True)
    assert bool((counts == 2).all())
    assert abs(scene.avatar_field(verts).abs().max().item()) < 0.05          # vertices sit near the level set
    assert int(f2t.min()) >= 0 and int(f2t.max()) < 6 * 32 ** 3


def test_binding_rule_one_or_three_per_face():
    verts, faces, _ = scene.avatar_mesh(32)
    fi, bary = scene.bind_faces(verts, faces)
    area = scene.face_areas(verts, faces)
    small = area < area.mean()
    n1, n3 = int(small.sum()), int((~small).sum())
    assert fi.numel() == n1 + 3 * n3
    assert torch.equal(fi[:n1], torch.where(small)[0])                         # faces-with-1 first
    assert torch.allclose(bary.sum(-1), torch.ones(fi.numel()))
    assert torch.allclose(bary[:n1], torch.full((n1, 3), 1 / 3))
    assert torch.allclose(bary[n1:n1 + 3], torch.tensor([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]]))


def test_scene_is_deterministic_and_sized():
    a = scene.make_gaussians(*scene.avatar_mesh(32)[:2], 5000, seed=0)
    b = scene.make_gaussians(*scene.avatar_mesh(32)[:2], 5000, seed=0)
    for k in ("face_index", "bary", "delta", "log_scales", "raw_quats", "opacity_logits", "shs"):
        assert torch.equal(a[k], b[k]), k
        assert a[k].shape[0] == 5000
    act = scene.activate(a)
    assert torch.allclose(act["rotations"].norm(dim=-1), torch.ones(5000), atol=1e-5)
    assert float(act["opacities"].min()) > 0 and float(act["opacities"].max()) < 1
    # flat Gaussians: the normal-direction scale is a tenth of the tangential one (tetgs_edit_2d.py:199-208 style)
    assert torch.allclose(act["scales"][:, 0] * 10, act["scales"][:, 1], rtol=1e-4)


def test_camera_conventions():
    cam = scene.orbit_camera(3, 8, 128, 128)
    V = cam["viewmatrix"].double()      # transposed world->view: flat memory column-major
    w2c = V.t()
    R = w2c[:3, :3]
    assert torch.allclose(R @ R.t(), torch.eye(3, dtype=torch.float64), atol=1e-6)
    c = -R.t() @ w2c[:3, 3]
    assert torch.allclose(c.float(), cam["campos"], atol=1e-5)
    centre_view = (torch.tensor([0.0, 0.0, -0.05, 1.0], dtype=torch.float64) @ V)
    assert abs(float(centre_view[0])) < 1e-6 and abs(float(centre_view[1])) < 1e-6 and float(centre_view[2]) > 2.5
    assert torch.allclose(cam["projmatrix"].double()[:, 3], V[:, 2], atol=1e-6)   # w_clip = z_view
    assert abs(cam["tanfovy"] - np.tan(np.radians(22.5)) / 1.4) < 1e-9


def test_shard_views_partition():
    for n, w in ((64, 8), (7, 2), (3, 4)):
        seen = sorted(v for r in range(w) for v in shard_views(n, r, w))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_views(4, 4, 4)


def test_grad_bucket_layout():
    b = GradBucket(10, 16, "cpu")
    assert [tuple(v.shape) for v in b.views] == list(grad_shapes(10, 16))
    for v in b.views:
        assert v.data_ptr() % 16 == 0
    b.views[5].fill_(2.0)
    assert float(b.flat.sum()) == 2.0 * 10 * 16 * 3
    # SH rows sit last in memory: everything else is one contiguous slice in front of them
    assert b.sh_offset == b.flat.numel() - 10 * 16 * 3 and b.views[5].data_ptr() == b.flat[b.sh_offset:].data_ptr()
    t = GradBucket(10, 16, "cpu", names=GradBucket.TRAINING)
    assert [n for n, v in zip(GRAD_NAMES, t.views) if v is not None] == list(GradBucket.TRAINING)
    assert t.flat.numel() < b.flat.numel()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, M, n_views = 37, 16, 6
    bucket = GradBucket(P, M, "cpu", names=GradBucket.TRAINING)
    mine = shard_views(n_views, rank, world)
    for i, v in enumerate(mine):                     # stand-in for the per-view backward: overwrite, then add
        for k, t in enumerate(bucket.views):
            if t is None:
                continue
            contrib = torch.full_like(t, float(v + 1) * (k + 1))
            if i == 0:
                t.copy_(contrib)
            else:
                t.add_(contrib)
    ref = bucket.flat.clone()
    bucket.all_reduce()
    want = sum(v + 1 for v in range(n_views))
    ok = all(torch.allclose(t, torch.full_like(t, float(want) * (k + 1))) for k, t in enumerate(bucket.views) if t is not None)
    # overlapped variant: SH rows in ranges (as the chunked backward hands them over), then the rest
    summed = bucket.flat.clone()
    bucket.flat.copy_(ref)
    for first in range(0, P, 16):
        bucket.all_reduce_sh_rows_async(first, min(16, P - first))
    bucket.all_reduce_rest_and_wait()
    ok = ok and torch.equal(bucket.flat, summed) and bucket._pending == []
    # all rows of a range in one (coalesced where the backend can) launch
    bucket.flat.copy_(ref)
    for first in range(0, P, 16):
        bucket.all_reduce_rows_async(first, min(16, P - first))
    bucket.wait()
    ok = ok and torch.equal(bucket.flat, summed) and bucket._pending == []
    q.put((rank, ok, mine))
    dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3, 5]

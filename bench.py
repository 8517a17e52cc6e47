#!/usr/bin/env python
"""bench.py — fwd+bwd rasterize views/s on the BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W [--impl reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A *step* = every rank renders V (default 8) different orbit views of the C3 scene — 1 M mesh-bound Gaussians,
1024x1024, SH degree 3, forward + backward with the depth/alpha extras — summing the per-Gaussian gradients of
its views in one flat buffer, then (N > 1) all-reduces that buffer once.  Ours renders the V views as ONE
multi-view batch (youreditableavatar_b200.multiview: one preprocess / backward-preprocess launch per batch, the
per-view stages on `--streams` CUDA streams); the reference arm has no such call and loops over views.
Prints ONE JSON line on rank 0.

  value   views/s over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e     same metric as a training step through the public operator API with HOST buffers: every view's camera
          and uint8 target image come from pinned host memory, the image loss and its upstream gradients are
          formed on the device from the rendered outputs (ours: colour MSE by this library's image-loss kernels;
          reference arm: the same loss with torch ops), and the step's loss is read back, all inside the timed
          region (Gaussian parameters are model state and stay resident, as in the reference's training loops)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md.
`--impl reference` times the UNMODIFIED reference CUDA rasterizer (oracle/_ref, built by oracle/build_ref.py)
on the same scene through its own `_C` entry points; if that build is absent it falls back to the CPU oracle.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+bwd rasterize views/s (1M Gaussians, 1024^2)"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--views-per-step", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=0,
                    help="0 (default): fused schedule, every per-view stage is one launch for the whole batch; "
                         "n >= 1: per-view launches round-robin on n CUDA streams")
    ap.add_argument("--comm", default="auto", choices=["auto", "nvls", "nvls_sh", "nccl", "rows", "sh"],
                    help="N > 1, gradient exchange: nvls = own in-switch all-reduce kernel on a symmetric-memory buffer, "
                         "nvls_sh = the same kernel per range of Gaussians (SH rows) on a side stream, overlapped with the "
                         "backward's per-Gaussian kernel, nccl = one NCCL all-reduce, auto (default) = nccl at N = 2 and "
                         "nvls_sh from N = 4, rows / sh = NCCL per Gaussian range overlapped with the backward")
    ap.add_argument("--legs", default="C1,C2,C5",
                    help="other BASELINE.json configs measured compactly after the headline config (N = 1: all listed; "
                         "N > 1: only C5, the config BASELINE.json quotes at 8 GPUs); empty string = none")
    ap.add_argument("--comm-chunks", type=int, default=2, help="Gaussian ranges for --comm rows / sh")
    ap.add_argument("--no-train-step", action="store_true",
                    help="skip the train_step leg (render + image loss + backward + Adam, SURVEY §8 f1-f3)")
    ap.add_argument("--per-view-api", action="store_true",
                    help="ours: loop over the single-view drop-in calls instead of the multi-view batch")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # NCCL kernels on a high-priority stream: the range-wise all-reduce has to get SMs while the backward's
        # per-Gaussian kernel (a full-grid, compute-heavy launch) is still running, or nothing overlaps
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ----------------------------------------------------------------------------------------------------
def build_workload(cfg, V, rank, world):
    from youreditableavatar_b200 import scene
    P, res, _, g = scene.CONFIGS[cfg]
    gs = scene.make_scene(cfg, device="cuda")
    act = scene.activate(gs)
    n_views = V * world
    cams = [scene.orbit_camera(v, n_views, res, res, device="cuda") for v in range(rank, n_views, world)]
    gen = torch.Generator().manual_seed(1 + rank)
    N = res * res
    up_host = []
    for _ in range(V):
        dc = (torch.randn(3, res, res, generator=gen) / (3 * N)).pin_memory()
        dd = (torch.randn(1, res, res, generator=gen) / N).pin_memory()
        da = (torch.randn(1, res, res, generator=gen) / N).pin_memory()
        up_host.append((dc, dd, da))
    up_dev = [tuple(t.cuda() for t in u) for u in up_host]
    # e2e leg: the step's host-side inputs are uint8 target images (what a dataset holds), one per view
    targets_host = torch.randint(0, 256, (V, 3, res, res), generator=gen, dtype=torch.uint8).pin_memory()
    return P, res, act, cams, up_host, up_dev, targets_host


def image_loss(color, depth, alpha, target_u8):
    """Loss of the e2e leg, formed on the device from the rendered outputs (plain torch elementwise ops — the
    reference's L1+SSIM loss is out of scope): mean squared colour error against the uint8 target,
    plus small depth / coverage terms when the extras are rendered.  Returns (dL_dcolor, dL_ddepth, dL_dalpha, loss);
    works on one view [3,H,W] or a batch [V,3,H,W]."""
    diff = color - target_u8.to(torch.float32).mul_(1.0 / 255.0)
    loss = (diff * diff).mean()
    dLc = diff * (2.0 / diff.numel())
    dLd = dLa = None
    if depth is not None:
        cov = alpha - 1.0
        loss = loss + 5e-4 * (depth * depth).mean() + 0.5 * (cov * cov).mean()
        dLd = depth * (1e-3 / depth.numel())
        dLa = cov * (1.0 / alpha.numel())
    return dLc, dLd, dLa, loss


class ResultReader:
    """Device->host read of the step's result: an asynchronous copy into pinned memory every step; the host
    consumes the value one step later (so it never stalls the GPU) and waits for the last one in drain()."""

    def __init__(self):
        self.pinned = torch.zeros(2, 1).pin_memory()
        self.events = [None, None]
        self.k = 0

    def push(self, result):
        slot = self.k & 1
        self.pinned[slot].copy_(result.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream())
        self.events[slot] = ev
        self.k += 1
        prev = self.events[slot ^ 1]
        if prev is not None:
            prev.synchronize()
            return float(self.pinned[slot ^ 1])
        return None

    def drain(self):
        for ev in self.events:
            if ev is not None:
                ev.synchronize()
        return float(self.pinned[(self.k - 1) & 1]) if self.k else None


def cam_to_host(cam):
    return {k: (v.cpu().pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in cam.items()}


def cam_to_dev(cam):
    return {k: (v.cuda(non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in cam.items()}


class HostFeeder:
    """Per-view host->device traffic of the e2e leg for the arms that render one view per call: the camera and
    the uint8 target image of view i+1 are copied from pinned memory on a side stream while view i renders.  The
    same feeder (and the same device-side loss) serves the reference arm and ours-per-view."""

    def __init__(self, host_cams, targets_host):
        self.cams, self.targets = host_cams, targets_host
        self.s_in = torch.cuda.Stream()
        self.keep = []
        self.slot = {}
        self.reader = ResultReader()

    def begin_step(self):
        self.s_in.wait_stream(torch.cuda.current_stream())   # recycled input buffers are no longer in use
        self.keep.clear()
        self.slot.clear()
        self._prefetch(0)

    def _prefetch(self, i):
        if i >= len(self.cams):
            return
        with torch.cuda.stream(self.s_in):
            cam = {k: (v.cuda(non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in self.cams[i].items()}
            tgt = self.targets[i].cuda(non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        self.slot[i] = (cam, tgt, ev)

    def view(self, i):
        cam, tgt, ev = self.slot[i]
        torch.cuda.current_stream().wait_event(ev)
        self.keep.append((cam, tgt))
        self._prefetch(i + 1)
        return cam, tgt

    def end_step(self, result):
        return self.reader.push(result)

    def drain(self):
        return self.reader.drain()


class BatchFeeder:
    """Host->device traffic of the e2e leg for the multi-view batch, double-buffered: while step k runs its backward,
    the cameras (one packed block) and the uint8 target images of step k+1 are copied from pinned host memory on a
    side stream; every step consumes its own fresh copy.  (The copy is issued behind the forward, not at the start of the
    step: concurrent with the forward's HBM-bound per-Gaussian and sort kernels it cost the step ~0.15 ms at N = 1 and
    0.3-0.5 ms at N = 8, where eight ranks pull from host memory at once; the blend backward is not HBM-bound.)"""

    def __init__(self, host_cams, targets_host):
        self.cams, self.targets = host_cams, targets_host
        self.s_in = torch.cuda.Stream()
        self.pending = None
        self.keep = []
        # all cameras of the batch travel as ONE pinned block [V, 38] = view 16 | proj 16 | campos 3 | bg 3
        self.cam_pack = torch.stack([torch.cat([c["viewmatrix"].flatten(), c["projmatrix"].flatten(),
                                                c["campos"].flatten(), c["bg"].flatten()]) for c in host_cams]).pin_memory()
        self.reader = ResultReader()

    def _issue(self):
        with torch.cuda.stream(self.s_in):
            pack = self.cam_pack.cuda(non_blocking=True)
            cams = []
            for v, c in enumerate(self.cams):
                d = dict(c)
                d["viewmatrix"], d["projmatrix"] = pack[v, 0:16].view(4, 4), pack[v, 16:32].view(4, 4)
                d["campos"], d["bg"] = pack[v, 32:35], pack[v, 35:38]
                cams.append(d)
            tgt = self.targets.cuda(non_blocking=True)
            ev = torch.cuda.Event(); ev.record(self.s_in)
        return cams, tgt, ev

    def begin_step(self):
        if self.pending is None:          # first step (or a caller that never prefetched)
            self.pending = self._issue()
        self.cur = self.pending
        self.pending = None
        torch.cuda.current_stream().wait_event(self.cur[2])
        self.keep = [self.cur]
        return self.cur[0]

    def targets_dev(self):
        return self.cur[1]

    def prefetch_next(self):
        """Called once forward and loss are queued: inputs of the NEXT step start streaming behind them (under the blend
        backward, which is not HBM-bound); their buffers must not be recycled under this step's kernels."""
        if self.pending is None:
            self.s_in.wait_stream(torch.cuda.current_stream())
            self.pending = self._issue()

    def end_step(self, result):
        return self.reader.push(result)

    def drain(self):
        return self.reader.drain()


class OursRunner:
    """Multi-view batch through youreditableavatar_b200.parallel.render_views_fwd_bwd."""
    name = "ours"
    n_up = 3

    def __init__(self, P, res, act, extras=True, n_streams=0, comm="nvls", comm_chunks=4, bucket=None):
        from youreditableavatar_b200.parallel import GradBucket, SymmGradBucket
        self.act, self.extras, self.n_streams, self.comm, self.comm_chunks = act, extras, n_streams, comm, comm_chunks
        # "nvls": the gradient buffer lives in symmetric memory and is summed by this library's own in-switch
        # all-reduce kernel (falls back to NCCL without multicast support); "nccl": one NCCL all-reduce;
        # "rows" / "sh": NCCL all-reduces of Gaussian ranges overlapped with the backward
        self.bucket = bucket if bucket is not None else \
            (SymmGradBucket if comm in ("nvls", "nvls_sh") else GradBucket)(P, 16, "cuda", names=GradBucket.TRAINING)
        self.loss_ws = None   # workspace of the image-loss kernels (e2e leg), allocated on first use

    # TGR_STEP_DIAG=1: CUDA events at step start / forward done / upstream done / step done, printed per rank at exit
    diag = os.environ.get("TGR_STEP_DIAG") == "1"
    diag_events = []

    @classmethod
    def diag_report(cls, rank):
        torch.cuda.synchronize()
        for leg in ("value", "e2e"):
            ev = [e[1] for e in cls.diag_events if e[0] == leg]
            ev = ev[len(ev) // 2:]                                # second half: the timed part of the leg
            if not ev:
                continue
            seg = [sum(a.elapsed_time(b) for a, b in zip([e[i] for e in ev], [e[i + 1] for e in ev])) / len(ev) for i in range(3)]
            gap = sum(ev[k][3].elapsed_time(ev[k + 1][0]) for k in range(len(ev) - 1)) / max(len(ev) - 1, 1)
            print("[diag rank %d %5s] forward %.3f  upstream %.3f  backward+exchange %.3f  gap to next step %.3f ms (%d steps)" % (
                rank, leg, seg[0], seg[1], seg[2], gap, len(ev)), file=sys.stderr)

    def _mark(self, evs):
        if self.diag:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)

    def step(self, cams, ups, world, feeder=None):
        from youreditableavatar_b200.parallel import render_views_fwd_bwd
        evs = []
        self._mark(evs)
        if feeder is not None:
            cams = feeder.begin_step()

        box = {}

        def upstream(color, depth, alpha):
            self._mark(evs)
            try:
                return upstream_(color, depth, alpha)
            finally:
                self._mark(evs)

        def upstream_(color, depth, alpha):
            if feeder is None:
                return ups if self.extras else (ups[0], None, None)
            # same loss as image_loss() below (what the reference arm evaluates with torch ops): the colour MSE against
            # the uint8 targets and its gradient come from this library's image-loss kernels (tgr_image_loss, L2
            # weights: two streaming launches for the whole batch), the small depth / coverage terms stay torch ops
            from youreditableavatar_b200 import loss_utils
            if self.loss_ws is None:
                from youreditableavatar_b200 import _lib
                V_, _, H_, W_ = color.shape
                self.loss_ws = torch.empty(_lib.lib().tgr_image_loss_bytes(V_, W_, H_), dtype=torch.uint8, device=color.device)
            out, dLc = loss_utils.image_loss_and_grad(color, feeder.targets_dev(), 0.0, 1.0, 0.0, workspace=self.loss_ws)
            loss, dLd, dLa = out[0], None, None
            if depth is not None:
                cov = alpha - 1.0
                loss = loss + 5e-4 * (depth * depth).mean() + 0.5 * (cov * cov).mean()
                dLd = depth * (1e-3 / depth.numel())
                dLa = cov * (1.0 / alpha.numel())
            box["loss"] = loss
            feeder.prefetch_next()
            return dLc, dLd, dLa

        if self.comm in ("rows", "sh", "nvls_sh"):   # all-reduces issued range by range from inside the backward
            render_views_fwd_bwd(self.act, cams, 3, upstream, self.bucket, extras=self.extras, n_streams=self.n_streams,
                                 all_reduce=True, comm_chunks=self.comm_chunks,
                                 comm_mode="sh" if self.comm == "nvls_sh" else self.comm)
        else:
            render_views_fwd_bwd(self.act, cams, 3, upstream, self.bucket, extras=self.extras, n_streams=self.n_streams)
            self.bucket.all_reduce()
        self._mark(evs)
        if self.diag:
            OursRunner.diag_events.append(("value" if feeder is None else "e2e", evs))
        if feeder is not None:
            return feeder.end_step(box["loss"])   # D2H read of the step's loss
        return None


class OursPerViewRunner:
    """The single-view drop-in calls in a Python loop with in-kernel gradient accumulation (what a caller that
    keeps the reference's one-view-per-call structure gets)."""
    name = "ours-per-view"

    def __init__(self, P, res, act, extras=True, bucket=None):
        from youreditableavatar_b200.parallel import GradBucket
        self.act, self.extras = act, extras
        self.bucket = bucket if bucket is not None else GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)

    def step(self, cams, ups, world, feeder=None):
        from youreditableavatar_b200 import rasterizer as rz
        e = torch.Tensor([])
        act = self.act
        loss = None
        if feeder is not None:
            feeder.begin_step()
        for i in range(len(cams)):
            cam, tgt = (cams[i], None) if feeder is None else feeder.view(i)
            fwd = rz.c_rasterize_gaussians(cam["bg"], act["means3D"], e, act["opacities"], act["scales"], act["rotations"],
                                           1.0, e, cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
                                           cam["image_height"], cam["image_width"], act["shs"], 3, cam["campos"], False,
                                           False, extras=self.extras)
            R, color, radii, geom, binning, img = fwd[:6]
            if feeder is None:
                up = ups[i] if self.extras else (ups[i][0], None, None)
            else:
                dLc, dLd, dLa, l = image_loss(color, fwd[6] if self.extras else None, fwd[7] if self.extras else None, tgt)
                up = (dLc, dLd, dLa)
                loss = l if loss is None else loss + l
            kw = dict(accumulate_into=self.bucket.views) if i > 0 else dict(out=self.bucket.views)
            rz.c_rasterize_gaussians_backward(cam["bg"], act["means3D"], radii, e, act["scales"], act["rotations"], 1.0, e,
                                              cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], up[0],
                                              act["shs"], 3, cam["campos"], geom, R, binning, img, False,
                                              dL_dout_depth=up[1], dL_dout_alpha=up[2], **kw)
        self.bucket.all_reduce()
        if feeder is not None:
            return feeder.end_step(loss)           # D2H read of the step's loss
        return None


class RefRunner:
    """The reference's own CUDA rasterizer (unmodified sources compiled into oracle/_ref)."""
    name = "reference"

    def __init__(self, P, res, act):
        from oracle import ref_cuda
        self.ref, self.act = ref_cuda, act
        self.acc = None

    def step(self, cams, ups, world, feeder=None):
        loss = None
        if feeder is not None:
            feeder.begin_step()
        for i in range(len(cams)):
            cam, tgt = (cams[i], None) if feeder is None else feeder.view(i)
            fwd = self.ref.forward(self.act, cam, 3)
            if feeder is None:
                dLc = ups[i][0]
            else:
                dLc, _, _, l = image_loss(fwd[1], None, None, tgt)   # it has no depth/alpha outputs
                loss = l if loss is None else loss + l
            grads = self.ref.backward(self.act, cam, 3, fwd, dLc)
            need = (2, 3, 5, 6, 7)  # opacity, means3D, sh, scales, rotations — what ours accumulates too
            if i == 0:
                self.acc = [grads[k] for k in need]   # first view: adopt the freshly zero-filled tensors
            else:
                for a, k in zip(self.acc, need):        # later views: what a user of the reference has to do
                    a.add_(grads[k])
        if world > 1:
            import torch.distributed as dist
            for a in self.acc:
                dist.all_reduce(a)
        if feeder is not None:
            return feeder.end_step(loss)           # D2H read of the step's loss
        return None


def timed(runner, cams, ups, world, steps, warmup, feeder=None):
    for _ in range(warmup):
        runner.step(cams, ups, world, feeder)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    from youreditableavatar_b200 import multiview as _mv
    t0, w0 = time.perf_counter(), _mv.host_wait_seconds
    for _ in range(steps):
        runner.step(cams, ups, world, feeder)
    # host time to issue a step, and the part of it that was CPU work (the rest: blocked on the instance counts of the
    # step just issued, i.e. waiting for the GPU to catch up)
    timed.host_ms_per_step = (time.perf_counter() - t0) * 1e3 / steps
    timed.host_cpu_ms_per_step = timed.host_ms_per_step - (_mv.host_wait_seconds - w0) * 1e3 / steps
    if feeder is not None:
        feeder.drain()                       # the last step's result has reached the host
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)  # ms


# ----------------------------------------------------------------------------------------------------
def cpu_oracle_baseline(cfg, budget_tiles=420):
    """float64 oracle (oracle/oracle.py) on a bounded sample of the same workload: the full per-Gaussian
    preprocess + fwd+bwd blending of `budget_tiles` non-empty tiles, extrapolated by list entries."""
    import numpy as np
    from oracle import oracle
    from youreditableavatar_b200 import scene
    P, res, _, g = scene.CONFIGS[cfg]
    torch.set_num_threads(os.cpu_count() or 1)
    gs = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in scene.make_scene(cfg, device="cuda").items()}
    act = {k: v.cpu() for k, v in scene.activate(gs).items()}
    cam = scene.orbit_camera(0, 8, res, res, device="cpu")
    t0 = time.time()
    inp = {k: v.double().requires_grad_(True) for k, v in act.items()}
    rec = oracle.preprocess(inp["means3D"], inp["opacities"], cam["viewmatrix"], cam["projmatrix"], cam["campos"], res, res,
                            cam["tanfovx"], cam["tanfovy"], scales=inp["scales"], rotations=inp["rotations"],
                            shs=inp["shs"], degree=3)
    xy32 = rec["xy"].detach().numpy().astype(np.float32)
    radii = torch.where(rec["valid"], rec["radius"].detach(), torch.zeros_like(rec["radius"])).numpy().astype(np.int32)
    keys, ids, ranges, cnt = oracle.build_keys(xy32, radii, rec["depth"].detach().numpy().astype(np.float32), res, res)
    t_pre = time.time() - t0
    lens = (ranges[:, 1].astype(np.int64) - ranges[:, 0].astype(np.int64))
    nonempty = np.flatnonzero(lens > 0)
    rng = np.random.RandomState(0)
    pick = rng.choice(nonempty, size=min(budget_tiles, len(nonempty)), replace=False)
    sub = np.zeros_like(ranges)
    sub[pick] = ranges[pick]
    t1 = time.time()
    color, depth, alpha, fT, nc = oracle.blend(rec, ids, sub, res, res, cam["bg"])
    (color.sum() + depth.sum() + alpha.sum()).backward()
    t_blend = time.time() - t1
    frac = lens[pick].sum() / max(lens.sum(), 1)
    per_view = t_pre + t_blend / max(frac, 1e-9)
    return {"value": 1.0 / per_view, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/oracle.py float64: full preprocess+keys+sort of %d Gaussians (%.1f s) + fwd+bwd blend of "
                      "%d of %d non-empty tiles (%.1f s, %.3f of the list entries), extrapolated to one view"
                      % (P, t_pre, len(pick), len(nonempty), t_blend, frac)}


# ----------------------------------------------------------------------------------------------------
# Training step through the rows either side of the rasterizer: render V views, the
# reference's image loss 0.8 L1 + 0.2 (1 - SSIM) against uint8 targets from pinned host memory, backward, Adam.
# Both arms run the same semantics (one optimizer step per V-view batch, gradients summed over the views); the
# learning rates are the reference's defaults x 1e-3 so that the synthetic scene — trained against random targets
# here — does not drift during the timed region (the kernels' cost does not depend on the rates).
def _train_opt():
    from youreditableavatar_b200.optimizer import OptimizationParams
    o = OptimizationParams()
    k = 1e-3
    return OptimizationParams(position_lr_init=o.position_lr_init * k, position_lr_final=o.position_lr_final * k,
                              feature_lr=o.feature_lr * k, opacity_lr=o.opacity_lr * k, scaling_lr=o.scaling_lr * k,
                              rotation_lr=o.rotation_lr * k)


def _time_events(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def train_step_ours(P, res, act, cams, targets_host, V, steps, warmup, peak):
    from youreditableavatar_b200 import _lib, loss_utils
    from youreditableavatar_b200.optimizer import TetGSOptimizer
    from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd
    L = _lib.lib()
    params = {k: act[k].clone() for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    bucket = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
    gv = bucket.named()
    opt = TetGSOptimizer({"points": params["means3D"], "sh": params["shs"], "all_densities": params["opacities"],
                          "scales": params["scales"], "quaternions": params["rotations"]}, _train_opt(), 1.0,
                         grads={"points": gv["dL_dmeans3D"], "sh": gv["dL_dsh"], "all_densities": gv["dL_dopacity"],
                                "scales": gv["dL_dscales"], "quaternions": gv["dL_drotations"]})
    feeder = BatchFeeder([cam_to_host(c) for c in cams], targets_host)
    ws = torch.empty(L.tgr_image_loss_bytes(V, res, res), dtype=torch.uint8, device="cuda")
    box = {}

    def upstream(color, depth, alpha):
        out, grad = loss_utils.image_loss_and_grad(color, feeder.targets_dev(), 0.8, 0.0, 0.2, workspace=ws)
        box["loss"] = out[0]
        feeder.prefetch_next()
        return grad, None, None

    def step():
        cs = feeder.begin_step()
        opt.update_learning_rate()
        render_views_fwd_bwd(params, cs, 3, upstream, bucket, extras=False)
        opt.step()
        return feeder.end_step(box["loss"])

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    n0 = L.tgr_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    last_loss = feeder.drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (L.tgr_kernel_launches() - n0) // steps
    # the two new kernels alone (inputs exceed the L2: 126 MB of images + 302 MB of maps; 944 MB of optimizer state)
    color = torch.rand(V, 3, res, res, device="cuda")
    tgt = targets_host.cuda()
    ms_loss = _time_events(lambda: loss_utils.image_loss_and_grad(color, tgt, 0.8, 0.0, 0.2, workspace=ws), 10)
    ms_adam = _time_events(opt.step, 10)
    n_px = V * 3 * res * res
    loss_bytes = n_px * (4 + 1 + 12 + 12 + 4 + 1 + 4)           # fwd 17 B + bwd 21 B per pixel-channel (u8 target)
    n_par = sum(g["params"][0].numel() for g in opt.optimizer.param_groups)
    adam_bytes = n_par * 28                                       # p, g, m, v read; p, m, v written
    return {"value": V * 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "views_per_step": V, "kernel_launches_per_step": int(launches),
            "loss": last_loss, "h2d_bytes_per_step": int(targets_host.numel() + V * 38 * 4), "d2h_bytes_per_step": 4,
            "what": "BatchFeeder (uint8 targets + cameras from pinned host memory) -> multi-view render -> fused "
                    "0.8 L1 + 0.2 (1 - SSIM) loss and gradient (tgr_image_loss, 3 launches) -> backward into the flat "
                    "bucket -> one-launch Adam over 5 parameter groups (tgr_adam_step) -> async loss read-back",
            "image_loss": {"ms": ms_loss, "algorithmic_bytes": loss_bytes, "achieved_gbs": loss_bytes / ms_loss / 1e6,
                           "frac_of_hbm_peak": loss_bytes / ms_loss / 1e6 / peak, "launches": 3},
            "adam": {"ms": ms_adam, "parameters": n_par, "algorithmic_bytes": adam_bytes,
                     "achieved_gbs": adam_bytes / ms_adam / 1e6, "frac_of_hbm_peak": adam_bytes / ms_adam / 1e6 / peak,
                     "launches": 1}}


def train_step_reference(P, res, act, cams, targets_host, V, steps, warmup):
    """The reference's own pieces: its CUDA rasterizer (oracle/_ref) one view per call, its loss as the ATen op
    sequence of utils/loss_utils.py in fp32 (oracle/train_oracle.image_loss(dtype=float32)) with autograd,
    torch.optim.Adam(eps=1e-15) over its six parameter groups (tetgs_optimizer.py:66-92)."""
    from oracle import ref_cuda, train_oracle
    o = _train_opt()
    leaf = lambda t: t.clone().requires_grad_(True)
    pts, opa, sca, rot = leaf(act["means3D"]), leaf(act["opacities"]), leaf(act["scales"]), leaf(act["rotations"])
    dc, rest = leaf(act["shs"][:, :1]), leaf(act["shs"][:, 1:])
    ta = torch.optim.Adam([{"params": [pts], "lr": o.position_lr_init}, {"params": [dc], "lr": o.feature_lr},
                           {"params": [rest], "lr": o.feature_lr / 20.0}, {"params": [opa], "lr": o.opacity_lr},
                           {"params": [sca], "lr": o.scaling_lr}, {"params": [rot], "lr": o.rotation_lr}], lr=0.0, eps=1e-15)
    feeder = HostFeeder([cam_to_host(c) for c in cams], targets_host)

    def step():
        feeder.begin_step()
        with torch.no_grad():
            cur = {"means3D": pts, "opacities": opa, "scales": sca, "rotations": rot,
                   "shs": torch.cat([dc, rest], dim=1)}                    # tetgs_model.py sh_coordinates: cat(dc, rest)
        acc, loss = None, None
        for i in range(V):
            cam, tgt = feeder.view(i)
            with torch.no_grad():
                fwd = ref_cuda.forward(cur, cam, 3)
            p = fwd[1].detach().requires_grad_(True)
            l, _ = train_oracle.image_loss(p[None], (tgt.float() / 255.0)[None], 0.8, 0.0, 0.2, dtype=torch.float32)
            (l / V).backward()
            loss = l.detach() / V if loss is None else loss + l.detach() / V
            with torch.no_grad():
                g = ref_cuda.backward(cur, cam, 3, fwd, p.grad)
                need = (2, 3, 5, 6, 7)
                if acc is None:
                    acc = [g[k] for k in need]
                else:
                    for a, k in zip(acc, need):
                        a.add_(g[k])
        opa.grad, pts.grad, sca.grad, rot.grad = acc[0], acc[1], acc[3], acc[4]
        dc.grad, rest.grad = acc[2][:, :1].contiguous(), acc[2][:, 1:].contiguous()   # what cat's backward produces
        ta.step()
        return feeder.end_step(loss)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    last_loss = feeder.drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    color = torch.rand(1, 3, res, res, device="cuda")
    tgt = torch.rand(1, 3, res, res, device="cuda")

    def loss_once():
        p = color.clone().requires_grad_(True)
        train_oracle.image_loss(p, tgt, 0.8, 0.0, 0.2, dtype=torch.float32)[0].backward()

    ms_loss = _time_events(loss_once, 10) * V
    ms_adam = _time_events(ta.step, 10)
    return {"value": V * 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "views_per_step": V, "loss": last_loss,
            "what": "HostFeeder -> reference CUDA rasterizer one view per call -> loss_utils.py op sequence in fp32 "
                    "(ATen conv2d + elementwise, autograd) -> gradient accumulation over the views -> torch.optim.Adam "
                    "over 6 groups -> async loss read-back",
            "image_loss": {"ms": ms_loss, "note": "V single-view fwd+bwd evaluations"},
            "adam": {"ms": ms_adam}}


def bound_path_leg(cfg, reps=10):
    """SURVEY §8 a1 on a measured path: one view, forward + backward from the RAW scene parameters (offsets along the
    normals, log-scales, un-normalised quaternions, opacity logits, SH) — (a) the way the reference's scene models do it:
    eager torch binding + activations with autograd around the rasterizer (tetgs_model.py:252-286), here around this
    library's drop-in operator; (b) `rasterize_bound`: binding, activations and their chain rule inside the preprocess
    kernels (csrc/preprocess.cu / preprocess_bwd.cu, BOUND variants)."""
    from youreditableavatar_b200 import scene
    from youreditableavatar_b200.binding import MeshBinding, rasterize_bound
    from youreditableavatar_b200.rasterizer import GaussianRasterizer
    from youreditableavatar_b200.parallel import settings_from_cam
    P, res, _, _ = scene.CONFIGS[cfg]
    gs = scene.make_scene(cfg, device="cuda")
    cam = scene.orbit_camera(0, 8, res, res, device="cuda")
    st = settings_from_cam(cam, 3)
    mesh = MeshBinding.from_scene(gs)
    names = ("delta", "log_scales", "raw_quats", "opacity_logits", "shs")
    raw = {k: gs[k].cuda().clone().requires_grad_(True) for k in names}
    dL = torch.randn(3, res, res, device="cuda") / (3 * res * res)
    rast = GaussianRasterizer(st)

    def eager():
        g = dict(gs)
        g.update(raw)
        act = scene.activate(g)
        m2 = torch.zeros_like(act["means3D"], requires_grad=True)
        color, _ = rast(means3D=act["means3D"], means2D=m2, opacities=act["opacities"], shs=act["shs"], scales=act["scales"],
                        rotations=act["rotations"])
        (color * dL).sum().backward()
        for v in raw.values():
            v.grad = None

    def fused():
        color, _ = rasterize_bound(raw["delta"], raw["log_scales"], raw["raw_quats"], raw["opacity_logits"], raw["shs"], mesh, st)
        (color * dL).sum().backward()
        for v in raw.values():
            v.grad = None

    ms_e, ms_f = _time_events(eager, reps), _time_events(fused, reps)
    return {"config": cfg, "what": "one view fwd+bwd from the raw parameters through autograd", "eager_binding_ms": ms_e,
            "fused_binding_ms": ms_f, "speedup": ms_e / ms_f}


_result_fd = None


def emit_result(obj):
    """The ONE JSON line of the run, written to the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _result_fd is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_result_fd, line)


def workload_string(cfg, V):
    from youreditableavatar_b200 import scene
    P, res, _, g = scene.CONFIGS[cfg]
    pretty = "%dM" % (P // 1_000_000) if P % 1_000_000 == 0 else "%dk" % (P // 1000)
    return ("%s: %s mesh/tet-bound Gaussians (synthetic avatar shell, marching tets on a %d^3 Kuhn grid), %dx%d, "
            "SH degree 3, fwd+bwd with depth/alpha outputs, %d views per step and GPU" % (cfg, pretty, g, res, res, V))


def pick_comm(args, world):
    """--comm auto: NCCL at N = 2 (two ranks gain nothing from the switch: measured 0.47 ms vs 0.63 ms for the in-switch
    kernel on 236 MB), this library's in-switch all-reduce kernel from N = 4."""
    if args.comm != "auto":
        return args.comm
    return "nvls_sh" if world >= 4 else "nccl"


def allreduce_check(bucket, world):
    """Outside the timed region: the library's in-switch all-reduce against NCCL on the same data, bit for bit."""
    if world == 1:
        return None
    import torch.distributed as dist
    if not getattr(bucket, "nvls", False):
        return "nccl (torch.distributed) is the exchange: nothing to cross-check"
    g = torch.Generator(device="cuda").manual_seed(1234 + dist.get_rank())
    bucket.flat.copy_(torch.randn(bucket.flat.numel(), device="cuda", generator=g) * 1e-3)
    want = bucket.flat.clone()
    dist.all_reduce(want)
    bucket.all_reduce()
    torch.cuda.synchronize()
    same = torch.equal(bucket.flat, want)
    worst = float((bucket.flat - want).abs().max())
    rel = float((bucket.flat.double() - want.double()).norm() / want.double().norm())
    flag = torch.tensor([0.0 if same else 1.0, worst, rel], device="cuda", dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if flag[0].item() == 0.0:
        return "bit-identical to NCCL all_reduce (%d floats, every rank)" % bucket.flat.numel()
    # fp32 addition is not associative: NCCL picks its algorithm (ring / tree / NVLS) per size and rank count, the
    # in-switch reduction adds in the switch's order — the sums may differ in the last bit
    verdict = "agrees with" if flag[2].item() <= 1e-6 else "MISMATCH vs"
    return "%s NCCL all_reduce to summation order: max-abs %.3g, rel-L2 %.3g (%d floats, max over ranks)" % (
        verdict, flag[1].item(), flag[2].item(), bucket.flat.numel())


def exchange_check(runner, P, act, cams, ups, world):
    """Outside the timed region: one whole step through the runner's exchange (also the range-wise, overlapped one)
    against the same step's local gradients summed with one NCCL all_reduce.  Two backward runs differ in the order of
    their floating-point reductions, so this is a relative-L2 figure, not a bit comparison."""
    import torch.distributed as dist
    from youreditableavatar_b200.parallel import GradBucket, render_views_fwd_bwd
    ref = GradBucket(P, 16, "cuda", names=GradBucket.TRAINING)
    render_views_fwd_bwd(act, cams, 3, lambda c, d, a: ups, ref, extras=True)
    dist.all_reduce(ref.flat)
    runner.step(cams, ups, world)
    torch.cuda.synchronize()
    rel = float((runner.bucket.flat.double() - ref.flat.double()).norm() / ref.flat.double().norm())
    t = torch.tensor([rel], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return "whole step vs local gradients + NCCL all_reduce: rel-L2 %.2e (max over ranks)" % float(t.item())


def measure(cfg, V, args, world, rank, local, steps, warmup, full):
    """One config through one arm: device-resident throughput (`value`), the e2e leg with host buffers, and (ours) the
    same workload through the single-view drop-in calls.  full=True adds clock sampling and per-stage event timing."""
    from youreditableavatar_b200 import scene
    ours = args.impl == "ours"
    P, res, act, cams, up_host, up_dev, targets_host = build_workload(cfg, V, rank, world)
    batched = ours and not args.per_view_api
    L = None
    if ours:                                   # the reference arm never loads this repo's library
        from youreditableavatar_b200 import _lib
        L = _lib.lib()
    comm = pick_comm(args, world)
    if batched:
        up_stack_dev = tuple(torch.stack([u[k] for u in up_host]).cuda() for k in range(3))
        runner = OursRunner(P, res, act, n_streams=args.streams, comm=comm, comm_chunks=args.comm_chunks)
        ups = up_stack_dev
    else:
        runner = OursPerViewRunner(P, res, act) if ours else RefRunner(P, res, act)
        ups = up_dev
    for _ in range(max(warmup, 3)):
        runner.step(cams, ups, world)
    sampler = ClockSampler(local) if full else None
    launches0 = L.tgr_kernel_launches() if ours else 0
    if ours and full:
        L.tgr_profile_enable(1)
    if sampler:
        sampler.start()
    ms = timed(runner, cams, ups, world, steps, 0)
    clocks = sampler.stop() if sampler else None
    launches = (L.tgr_kernel_launches() - launches0) if ours else None
    stage = {}
    if ours and full:
        from youreditableavatar_b200 import _lib
        sums = (C.c_float * _lib.NUM_STAGES)()
        cnts = (C.c_int32 * _lib.NUM_STAGES)()
        L.tgr_profile_collect(sums, cnts)
        stage = {n: {"ms_avg": (sums[i] / cnts[i]) if cnts[i] else None, "launches": int(cnts[i])}
                 for i, n in enumerate(_lib.STAGE_NAMES)}
        L.tgr_profile_enable(0)
    views = V * world * steps
    host_ms = getattr(timed, "host_ms_per_step", None)
    out = {"P": P, "res": res, "views_per_step_per_gpu": V, "value": views / (ms / 1000.0), "ms_per_step": ms / steps,
           "host_issue_ms_per_step": host_ms, "host_cpu_ms_per_step": getattr(timed, "host_cpu_ms_per_step", None),
           "gpu_launches": None if launches is None else int(launches), "clocks": clocks, "stages": stage}

    host_cams = [cam_to_host(c) for c in cams]
    feeder = BatchFeeder(host_cams, targets_host) if batched else HostFeeder(host_cams, targets_host)
    ms_e2e = timed(runner, cams, ups, world, steps, max(warmup, 3), feeder)
    cam_bytes = sum(v.numel() * 4 for v in host_cams[0].values() if isinstance(v, torch.Tensor))
    out["e2e"] = {"value": views / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": V * cam_bytes + targets_host.numel(),
                  "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps,
                  "host_issue_ms_per_step": getattr(timed, "host_ms_per_step", None),
                  "host_cpu_ms_per_step": getattr(timed, "host_cpu_ms_per_step", None)}
    if batched:
        pv = OursPerViewRunner(P, res, act, bucket=runner.bucket)
        k = max(2, steps // 4)
        ms_pv = timed(pv, cams, up_dev, world, k, 2)
        out["per_view_api"] = {"value": V * world * k / (ms_pv / 1000.0), "unit": UNIT,
                               "note": "same workload through GaussianRasterizer-style single-view calls in a loop, one stream"}
        if world > 1 and full:
            out["allreduce_check"] = allreduce_check(runner.bucket, world)
            out["exchange_check"] = exchange_check(runner, P, act, cams, ups, world)
        out["comm"] = ("in-switch NVLS all-reduce kernel of this library" if getattr(runner.bucket, "nvls", False) else
                       "NCCL" if comm in ("nvls", "nvls_sh", "nccl") else "NCCL, %d ranges overlapped" % args.comm_chunks) if world > 1 else None
        if out["comm"] and comm == "nvls_sh" and getattr(runner.bucket, "nvls", False):
            out["comm"] += ", SH rows in %d ranges on a side stream overlapped with the backward's per-Gaussian kernel" % args.comm_chunks
    out["_keep"] = (P, res, act, cams, targets_host, runner)
    return out


def stage_table(P, res, V, Rs, stage_ms, peak_gbs, sm_max_mhz):
    """Per-stage roofline table of one fused batch: CUDA-event time, algorithmic bytes against the
    measured HBM peak, and — from the committed ncu capture — warp instructions against the SMs' issue rate."""
    N, T, R = res * res, ((res + 15) // 16) ** 2, sum(Rs)
    tb = max(1, (T - 1).bit_length())
    depth_passes = tile_passes = None
    try:
        meta = json.load(open(os.path.join(ROOT, "profiles", "stage_metrics_C3_batch8.json")))
    except Exception:
        meta = {"stages": {}}
    npass = lambda name, default: len([k for k in meta["stages"].get(name, {}).get("kernels", []) if "radix_pass" in k["name"]]) or default
    depth_passes, tile_passes = npass("depth_sort", 4), npass("tile_sort", (tb + 7) // 8)
    alg = {
        "preprocess": P * 236 + V * P * 66,
        "depth_sort": V * P * (4 + depth_passes * 16),
        "emit": V * P * 24 + R * 8,
        "tile_sort": R * (4 + tile_passes * 16),
        "ranges": R * 4 + V * T * 16,
        "blend_fwd": V * N * 28 + R * 52,
        "blend_bwd": V * (N * 20 + P * 48) + R * 40,
        "preprocess_bwd": V * P * 53 + P * 236 * 2,
    }
    issue_peak = 148 * 4 * sm_max_mhz * 1e6          # warp instructions per second, all schedulers issuing every cycle
    same_workload = (P, res, V) == (1_000_000, 1024, 8)
    table = {}
    for name, a in alg.items():
        ms = (stage_ms.get(name) or {}).get("ms_avg")
        if not ms:
            continue
        e = {"ms": ms, "algorithmic_bytes": int(a), "achieved_gbs": a / ms / 1e6, "hbm_frac": a / ms / 1e6 / peak_gbs}
        m = meta["stages"].get(name)
        if m and same_workload:
            e["warp_instructions"] = int(m["warp_instructions"])
            e["issue_frac"] = m["warp_instructions"] / issue_peak / (ms * 1e-3)
            e["dram_bytes_ncu"] = int(m["dram_bytes"])
        table[name] = e
    return table, meta


def main():
    global _result_fd
    args = parse()
    # stdout carries exactly one JSON line: libraries write to fd 1 directly (NCCL prints its version banner there
    # when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the run and the result goes to the saved descriptor
    sys.stdout.flush()
    _result_fd = os.dup(1)
    os.dup2(2, 1)
    world, rank, local = dist_setup(args)
    if args.gpus != world and rank == 0 and world > 1:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)
    V = args.views_per_step
    cfg = args.config
    ours = args.impl == "ours"

    if not ours:
        from oracle import ref_cuda
        if not ref_cuda.available():
            if rank == 0:
                cb = cpu_oracle_baseline(cfg)
                emit_result(({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
                              "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": 1000.0 / cb["value"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic", "config": {"workload": workload_string(cfg, V)}, "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0},
                              "note": "reference CUDA build (oracle/_ref) absent: CPU oracle port timed instead"}))
            return

    m = measure(cfg, V, args, world, rank, local, args.steps, args.warmup, full=True)
    if OursRunner.diag:
        OursRunner.diag_report(rank)
    P, res, act, cams, targets_host, runner = m.pop("_keep")
    batched = ours and not args.per_view_api

    # ---- the other BASELINE.json configs, compact: same arm, same timing rules, fewer steps ----------------------
    legs = {}
    leg_cfgs = [c for c in args.legs.split(",") if c and c != cfg]
    if world > 1:
        leg_cfgs = [c for c in leg_cfgs if c == "C5"]       # C5 is the config BASELINE.json quotes at 8 GPUs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_max = float((m["clocks"] or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(cfg, V), "views_per_step_per_gpu": V, "global_views_per_step": V * world,
                       "parallelism": "dp%d over views, flat gradient buffer all-reduced once per step" % world,
                       "cache": "inputs (236 B of parameters per Gaussian + V different cameras) exceed the 126 MB L2 from "
                                "C2 upwards; no flush"},
            "api": (("multi-view batch (MultiViewRasterizer / tgr_*_batch), " +
                     ("fused per-stage launches" if args.streams == 0 else "%d streams" % args.streams)) if batched
                    else "single-view calls in a loop"),
            "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "clocks": m["clocks"],
            "host_issue_ms_per_step": m["host_issue_ms_per_step"], "host_cpu_ms_per_step": m["host_cpu_ms_per_step"],
        }
        out["e2e"]["note"] = ("both arms: cameras + uint8 targets from pinned host memory, colour MSE (+ depth / coverage terms "
                              "where rendered) formed on the device, loss read back every step.  Ours evaluates the MSE and "
                              "its gradient with this library's image-loss kernels, the reference arm with torch elementwise "
                              "ops, and the reference arm pays add_ accumulation of 5 gradient tensors per view (its API "
                              "renders one view per call): what a user of each pays for the same training step")
        if m.get("comm"):
            out["exchange"] = m["comm"]      # (top level: `config` describes the workload and is the same in both arms)
        if m.get("allreduce_check"):
            out["allreduce_check"] = m["allreduce_check"]
            out["exchange_check"] = m.get("exchange_check")
    if ours and rank == 0:
        from youreditableavatar_b200 import multiview as mv
        from youreditableavatar_b200.parallel import settings_from_cam
        e = torch.Tensor([])
        st = mv.c_rasterize_views([settings_from_cam(c, 3) for c in cams], act["means3D"], e, act["opacities"],
                                  act["scales"], act["rotations"], e, act["shs"], extras=True)[0]
        Rs = list(st.counts)
        del st
        if batched and args.streams == 0:
            table, meta = stage_table(P, res, V, Rs, m["stages"], peak, sm_max)
            dom = max(table, key=lambda k: table[k]["ms"]) if table else None
            if dom:
                d = table[dom]
                bound = "issue" if "issue_frac" in d and d["issue_frac"] > d["hbm_frac"] else "hbm"
                out["roofline"] = {
                    "bound": bound, "kernel": dom,
                    "achieved": (d["warp_instructions"] / (d["ms"] * 1e-3) / 1e9) if bound == "issue" else d["achieved_gbs"],
                    "peak": (148 * 4 * sm_max * 1e6 / 1e9) if bound == "issue" else peak,
                    "unit": "G warp-instructions/s" if bound == "issue" else "GB/s",
                    "frac": d["issue_frac"] if bound == "issue" else d["hbm_frac"],
                    "hbm_frac": d["hbm_frac"], "hbm_achieved_gbs": d["achieved_gbs"], "hbm_peak_gbs": peak,
                    "traffic": d.get("dram_bytes_ncu"), "algorithmic_bytes_per_launch": d["algorithmic_bytes"],
                    "ms_per_launch": d["ms"], "views_per_launch": V, "num_rendered_per_view": Rs,
                    "peak_source": ("MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s") +
                                   "; issue peak = 148 SMs x 4 schedulers x %.0f MHz" % sm_max,
                    "instruction_source": "profiles/stage_metrics_C3_batch8.json (%s)" % meta.get("note", "ncu --set full"),
                    "timing": "CUDA events around every launch of the stage inside the timed region, on the launching stream",
                    "note": "the blend kernels gather 16-byte records that mostly hit in L2 and evaluate exp / FMA chains per "
                            "(pixel, Gaussian) pair: they are bound by the SMs' instruction issue rate, not by HBM — `frac` is "
                            "issued warp instructions per second over the issue peak, `hbm_frac` the algorithmic-bytes figure"}
            out["kernels"] = table
        out["stages"] = m["stages"]
        if m.get("per_view_api"):
            out["per_view_api"] = m["per_view_api"]
    if ours and batched and world == 1 and not args.no_train_step:
        try:
            out["bound_path"] = bound_path_leg(cfg)
        except Exception as ex:
            out["bound_path"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        try:
            out["train_step"] = train_step_ours(P, res, act, cams, targets_host, V, max(4, args.steps // 2),
                                                max(args.warmup, 3), peak)
        except Exception as ex:   # the rasterize metric above stands on its own
            out["train_step"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    if not ours and world == 1 and not args.no_train_step:
        try:
            out["train_step"] = train_step_reference(P, res, act, cams, targets_host, V, max(4, args.steps // 2),
                                                     max(args.warmup, 3))
        except Exception as ex:
            out["train_step"] = {"error": "%s: %s" % (type(ex).__name__, ex)}

    # free the headline workload before the legs (C5 alone holds several GB of workspaces per view)
    del runner, act, cams, targets_host
    from youreditableavatar_b200 import scene
    scene._scene_cache.clear()
    torch.cuda.empty_cache()
    for c in leg_cfgs:
        Vc = {"C1": 1, "C2": 4}.get(c, V)           # BASELINE.json: C1 one view, C2 a batch of 4, C3 / C5 8 per step
        try:
            r = measure(c, Vc, args, world, rank, local, max(5, args.steps // 2), 3, full=False)
            r.pop("_keep")
            legs[c] = {"workload": workload_string(c, Vc), "value": r["value"], "ms_per_step": r["ms_per_step"],
                       "e2e": r["e2e"]["value"], "gpu_launches": r["gpu_launches"]}
            if r.get("per_view_api"):
                legs[c]["per_view_api"] = r["per_view_api"]["value"]
        except Exception as ex:
            legs[c] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        scene._scene_cache.clear()
        torch.cuda.empty_cache()
    if rank != 0:
        return
    if legs:
        out["configs"] = legs
    if ours:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_oracle_baseline(cfg)
    else:
        out["impl"] = "reference"
        out["reference_kind"] = "reference CUDA rasterizer, unmodified sources compiled for sm_100a into oracle/_ref"
        out["gpu_launches"] = None
        out["cpu_baseline"] = {"value": out["value"], "unit": UNIT, "cores": 0, "kind": "reference",
                               "sample": "not a CPU run: the reference ships no CPU rasterizer; this arm is its CUDA "
                                         "rasterizer on the same B200 (see ours-arm cpu_baseline for the CPU oracle)"}
    emit_result(out)


if __name__ == "__main__":
    main()

"""Adam over the Gaussian parameter groups in one launch.

Host-side mirror of Edit_core/tetgs_scene/tetgs_optimizer.py: `OptimizationParams` (:8-31, same defaults),
`TetGSOptimizer` (:47-126: `step`, `zero_grad`, `update_learning_rate`, `state_dict`, `load_state_dict`,
`add_param_group`) and the learning-rate schedule `get_expon_lr_func` (Edit_core/utils/general_utils.py:25-58).
Where the reference wraps `torch.optim.Adam(l, lr=0.0, eps=1e-15)` (:92) — a chain of multi-tensor kernels that
streams parameters, gradients and both moments several times per step — `FusedAdam.step()` is ONE kernel of this
library (`tgr_adam_step`, csrc/adam.cu) over all groups, reading the gradients where the rasterizer's backward (or
the all-reduce) left them: `.grad` of each parameter, or slices of a flat `GradBucket` (parallel.py).
No CPU path: CPU parameters raise.
"""
import ctypes as C
import math
from typing import Callable, Dict, Iterable, List, Optional

import torch

from . import _lib
from ._lib import TgrAdamGroup, check

__all__ = ["get_expon_lr_func", "OptimizationParams", "FusedAdam", "TetGSOptimizer"]


def get_expon_lr_func(lr_init: float, lr_final: float, lr_delay_steps: int = 0, lr_delay_mult: float = 1.0,
                      max_steps: int = 1000000) -> Callable[[int], float]:
    """general_utils.py:25-58 — log-linear interpolation from lr_init to lr_final over max_steps, optionally eased
    in over lr_delay_steps; 0 for negative steps or when both ends are 0."""

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(
                0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        else:
            delay_rate = 1.0
        t = min(max(step / max_steps, 0.0), 1.0)
        return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)

    return helper


class OptimizationParams:
    """tetgs_optimizer.py:8-31"""

    def __init__(self, iterations: int = 15_000, position_lr_init: float = 0.00016, position_lr_final: float = 0.0000016,
                 position_lr_delay_mult: float = 0.01, position_lr_max_steps: int = 30_000, feature_lr: float = 0.0025,
                 opacity_lr: float = 0.05, scaling_lr: float = 0.005, rotation_lr: float = 0.001):
        self.iterations = iterations
        self.position_lr_init = position_lr_init
        self.position_lr_final = position_lr_final
        self.position_lr_delay_mult = position_lr_delay_mult
        self.position_lr_max_steps = position_lr_max_steps
        self.feature_lr = feature_lr
        self.opacity_lr = opacity_lr
        self.scaling_lr = scaling_lr
        self.rotation_lr = rotation_lr


class FusedAdam:
    """torch.optim.Adam semantics (betas, eps, no weight decay, no amsgrad) with torch-style `param_groups`
    (dicts with "params": [one tensor], "lr", "name") and `state`, one kernel per step for all groups.
    Extra per-group keys:
      "lr_alt", "period", "split"   element i of the (flattened) parameter uses "lr" when i % period < split and
                                    "lr_alt" otherwise — SH stored as [P,M,3] rows with the reference's two rates
                                    (dc: feature_lr, rest: feature_lr / 20, tetgs_optimizer.py:77-85);
      "grad"                        tensor to read the gradient from instead of `param.grad` (a GradBucket view)."""

    def __init__(self, param_groups: Iterable[Dict], lr: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8):
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.param_groups: List[Dict] = []
        self.state: Dict[int, Dict] = {}
        self.grad_scale = 1.0
        for g in param_groups:
            self.add_param_group(g)

    def add_param_group(self, group: Dict) -> None:
        g = dict(group)
        params = g["params"]
        params = [params] if isinstance(params, torch.Tensor) else list(params)
        if len(params) != 1:
            raise ValueError("FusedAdam: one tensor per param group (as TetGSOptimizer builds them)")
        p = params[0]
        if not p.is_cuda:
            raise RuntimeError("FusedAdam: parameters must be CUDA tensors (there is no CPU path)")
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("FusedAdam: parameters must be contiguous float32")
        g["params"] = [p]
        g.setdefault("lr", self.defaults["lr"])
        g.setdefault("betas", self.defaults["betas"])
        g.setdefault("eps", self.defaults["eps"])
        if len(self.param_groups) >= _lib.ADAM_MAX_GROUPS:
            raise RuntimeError("FusedAdam: at most %d param groups" % _lib.ADAM_MAX_GROUPS)
        if self.param_groups and (tuple(g["betas"]) != tuple(self.param_groups[0]["betas"]) or
                                  g["eps"] != self.param_groups[0]["eps"]):
            raise RuntimeError("FusedAdam: betas and eps are shared by all groups (one launch)")
        self.param_groups.append(g)

    def _check_shared_hyperparameters(self) -> None:
        """One launch steps every group with the same betas and eps; groups that disagree (only possible through
        `load_state_dict` or by editing `param_groups`) must not be stepped silently with group 0's values."""
        for i, g in enumerate(self.param_groups[1:], 1):
            g0 = self.param_groups[0]
            if tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"]:
                raise RuntimeError("FusedAdam: group %r has betas / eps %r / %r but group %r has %r / %r — they are shared "
                                   "by all groups (one launch)" % (g.get("name", i), tuple(g["betas"]), g["eps"],
                                                                   g0.get("name", 0), tuple(g0["betas"]), g0["eps"]))

    def _state_of(self, idx: int, p: torch.Tensor) -> Dict:
        st = self.state.get(idx)
        if st is None:
            st = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            self.state[idx] = st
        return st

    @torch.no_grad()
    def step(self) -> None:
        # groups that share (step count, device) go into one launch — all of them in the reference's loops, where
        # every parameter receives a gradient every iteration; a group that skipped iterations (no gradient then,
        # torch.optim skips it too) has its own bias correction and therefore its own launch
        ready = []
        for idx, g in enumerate(self.param_groups):
            p = g["params"][0]
            grad = g.get("grad")
            if grad is None:
                grad = p.grad
            if grad is None:
                continue                       # torch.optim skips parameters without a gradient
            if grad.dtype != torch.float32 or grad.numel() != p.numel() or grad.device != p.device:
                raise RuntimeError("FusedAdam: gradient of group %r does not match its parameter" % g.get("name", idx))
            ready.append((idx, g, p, grad if grad.is_contiguous() else grad.contiguous()))
        self._check_shared_hyperparameters()
        batches: Dict[tuple, list] = {}
        for idx, g, p, grad in ready:          # nothing is counted or launched unless every group validated
            st = self._state_of(idx, p)
            batches.setdefault((st["step"] + 1, p.device), []).append((g, p, grad, st))
        b1, b2 = self.param_groups[0]["betas"] if self.param_groups else (0.9, 0.999)
        for (step, device), entries in batches.items():
            groups = (TgrAdamGroup * _lib.ADAM_MAX_GROUPS)()
            keep = []
            for n, (g, p, grad, st) in enumerate(entries):
                e = groups[n]
                e.param, e.grad = p.data_ptr(), grad.data_ptr()
                e.exp_avg, e.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                e.count = p.numel()
                e.lr = float(g["lr"])
                e.lr_alt = float(g.get("lr_alt", g["lr"]))
                e.period, e.split = int(g.get("period", 0)), int(g.get("split", 0))
                keep.append(grad)
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                check(_lib.lib().tgr_adam_step(groups, len(entries), step, float(b1), float(b2),
                                               float(self.param_groups[0]["eps"]), float(self.grad_scale), stream),
                      "tgr_adam_step")
            for _, _, _, st in entries:        # counted only once the launch of its batch has been accepted
                st["step"] = step
            del keep                           # the launch is stream-ordered after the producers of `grad`

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            p = g["params"][0]
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self) -> Dict:
        """Same schema as torch.optim.Optimizer.state_dict (the reference checkpoints it, refine.py:395-402)."""
        groups = []
        for i, g in enumerate(self.param_groups):
            d = {k: v for k, v in g.items() if k not in ("params", "grad", "_keepalive")}
            d["params"] = [i]
            groups.append(d)
        state = {i: {"step": torch.tensor(float(s["step"])), "exp_avg": s["exp_avg"], "exp_avg_sq": s["exp_avg_sq"]}
                 for i, s in self.state.items()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: Dict) -> None:
        if len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        merged = [dict(g, **{k: v for k, v in d.items() if k != "params"})
                  for g, d in zip(self.param_groups, sd["param_groups"])]
        live, self.param_groups = self.param_groups, merged
        try:                                   # a checkpoint with per-group betas / eps cannot be stepped by one launch:
            self._check_shared_hyperparameters()   # refused before anything of the live optimizer is touched
        except RuntimeError:
            self.param_groups = live
            raise
        self.state = {}
        for i, s in sd["state"].items():
            p = self.param_groups[int(i)]["params"][0]
            self.state[int(i)] = {"step": int(float(s["step"])),
                                  "exp_avg": s["exp_avg"].to(device=p.device, dtype=torch.float32).contiguous().clone(),
                                  "exp_avg_sq": s["exp_avg_sq"].to(device=p.device, dtype=torch.float32).contiguous().clone()}


class TetGSOptimizer:
    """tetgs_optimizer.py:47-126 without the model object: `params` maps the reference's group names to tensors —
    "points" (or the bound model's per-Gaussian offsets), "sh_coordinates_dc" + "sh_coordinates_rest" or a single
    "sh" [P,M,3] tensor in the rasterizer's row layout (then dc / rest learning rates are applied inside the row),
    "all_densities", "scales", "quaternions"; absent names are frozen, as the reference's learn_* flags do.
    `grads` optionally maps names to gradient tensors (GradBucket views) read instead of `.grad`."""

    def __init__(self, params: Dict[str, torch.Tensor], opt: Optional[OptimizationParams] = None,
                 spatial_lr_scale: float = 1.0, grads: Optional[Dict[str, torch.Tensor]] = None):
        if opt is None:
            opt = OptimizationParams()
        self.current_iteration = 0
        self.num_iterations = opt.iterations
        self.spatial_lr_scale = spatial_lr_scale
        grads = grads or {}
        l = []

        def add(name, lr, **kw):
            if name in params:
                g = {"params": [params[name]], "lr": lr, "name": name}
                if name in grads:
                    g["grad"] = grads[name]
                g.update(kw)
                l.append(g)

        add("points", opt.position_lr_init * spatial_lr_scale)
        add("sh_coordinates_dc", opt.feature_lr)
        add("sh_coordinates_rest", opt.feature_lr / 20.0)
        if "sh" in params:
            M = params["sh"].shape[1]
            add("sh", opt.feature_lr, lr_alt=opt.feature_lr / 20.0, period=3 * M, split=3)
        add("all_densities", opt.opacity_lr)
        add("scales", opt.scaling_lr)
        add("quaternions", opt.rotation_lr)
        self.optimizer = FusedAdam(l, lr=0.0, eps=1e-15)
        self.position_sheduler_func = get_expon_lr_func(
            lr_init=opt.position_lr_init * spatial_lr_scale, lr_final=opt.position_lr_final * spatial_lr_scale,
            lr_delay_mult=opt.position_lr_delay_mult, max_steps=opt.position_lr_max_steps)

    def step(self):
        self.optimizer.step()
        self.current_iteration += 1

    def zero_grad(self, set_to_none: bool = True):
        self.optimizer.zero_grad(set_to_none=set_to_none)

    def update_learning_rate(self, iteration: Optional[int] = None):
        if iteration is None:
            iteration = self.current_iteration
        lr = 0.
        for param_group in self.optimizer.param_groups:
            if param_group["name"] == "points":
                lr = self.position_sheduler_func(iteration)
                param_group["lr"] = lr
        return lr

    def add_param_group(self, new_param_group):
        self.optimizer.add_param_group(new_param_group)

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict)

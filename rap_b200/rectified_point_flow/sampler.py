"""`rectified_point_flow.sampler` on the fused sm_100a path (reference sampler.py:11-171).

`get_sampler(name)` returns `partial(flow_sampler, euler_step)` exactly like the reference; only 'euler'
is registered there (sampler.py:163-167) and unknown names raise the same ValueError.

`flow_sampler` has two routes, both on the GPU:
  * fused  -- when `flow_model_fn` is a closure over a rap_b200 `PointCloudDiT` (the closure
    `RectifiedPointFlow.sample_rectified_flow` builds, modeling.py:672-708, or any callable carrying a
    `rap_context` attribute), the whole S-step loop runs inside `rap_sample_euler` with x_t resident
    on the device and both trajectories written by kernel epilogues;
  * opaque -- any other callable is evaluated once per step and only the rigidity forcing uses the
    fused Procrustes kernel.
"""
from __future__ import annotations

from functools import partial
from typing import Callable

import torch

from . import procrustes
from .flow_model import PointCloudDiT
from .procrustes import rigidify_prediction_with_procrustes  # noqa: F401  (re-exported like the reference)


def _closure_vars(fn: Callable) -> dict:
    code, cells = getattr(fn, "__code__", None), getattr(fn, "__closure__", None)
    if code is None or not cells:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell
        except ValueError:  # pragma: no cover
            pass
    return out


def _fused_context(flow_model_fn: Callable):
    """Returns (model, ctx dict, cells) if the callable exposes everything the fused sampler needs."""
    ctx = getattr(flow_model_fn, "rap_context", None)
    if ctx is not None:
        return ctx["model"], ctx, {}
    cells = _closure_vars(flow_model_fn)
    need = ("self", "cond", "local_features", "scales", "anchor_indices", "cu_seqlens_batch", "cu_seqlens_part")
    if not all(k in cells for k in need):
        return None
    try:
        vals = {k: cells[k].cell_contents for k in need}
        latent = cells["latent_features"].cell_contents if "latent_features" in cells else None
    except ValueError:
        return None
    model = getattr(vals["self"], "flow_model", None)
    if not isinstance(model, PointCloudDiT) or latent is not None:
        return None
    ctx = dict(model=model, cond=vals["cond"], local_features=vals["local_features"], scales=vals["scales"],
               anchor_indices=vals["anchor_indices"], cu_seqlens_batch=vals["cu_seqlens_batch"],
               cu_seqlens_part=vals["cu_seqlens_part"])
    return model, ctx, cells


def flow_sampler(step_fn: Callable, flow_model_fn: Callable, x_1: torch.Tensor, x_0: torch.Tensor, anchor_indices: torch.Tensor,
                 num_steps: int = 20, points_per_part: torch.Tensor | None = None, cu_seqlens_batch: torch.Tensor | None = None,
                 condition: torch.Tensor | None = None, return_trajectory: bool = False, rigidity_forcing: bool = False,
                 return_end_point_trajectory: bool = True):
    """Same contract as the reference (sampler.py:11-74): returns `{'end_point_trajectory','trajectory'}` (each
    (S,*x_1.shape) fp32) when `return_trajectory`, else the final x_t."""
    fused = _fused_context(flow_model_fn) if step_fn is euler_step else None
    if fused is not None:
        model, ctx, cells = fused
        plan = model.plan(ctx["cu_seqlens_batch"], ctx["cu_seqlens_part"], points_per_part, max_timesteps=max(64, num_steps))
        plan.set_condition(ctx["cond"], ctx["local_features"], ctx["scales"], ctx["anchor_indices"])
        if rigidity_forcing and condition is not None and condition.data_ptr() != ctx["cond"].data_ptr():
            raise ValueError("fused sampler: `condition` must be the tensor the flow model is conditioned on")
        want_feats = bool(cells.get("return_transformer_features") and cells["return_transformer_features"].cell_contents)
        out = plan.sample(x_1, num_steps, rigidity_forcing, return_trajectory=return_trajectory, return_features=want_feats)
        if want_feats:   # reproduce the closure's side effects (modeling.py:678-695)
            cells["transformer_features"].cell_contents = out["transformer_features"]
            if "last_call_captured" in cells:
                cells["last_call_captured"].cell_contents[0] = True
            if "call_count" in cells:
                cells["call_count"].cell_contents[0] = max(num_steps - 1, 0)
        if return_trajectory:
            shape = (num_steps, *x_1.shape)
            return {"end_point_trajectory": out["end_point_trajectory"].view(shape), "trajectory": out["trajectory"].view(shape)}
        return out["x_final"].view(x_1.shape)

    # ---- opaque callable: per-step evaluation on the GPU
    dt = 1.0 / num_steps
    x_t = x_1.clone()
    trajectory = torch.empty((num_steps, *x_1.shape), device=x_1.device) if return_trajectory else None
    trajectory_x_t = torch.empty((num_steps, *x_1.shape), device=x_1.device) if return_trajectory else None
    for step in range(num_steps):
        t = 1 - step * dt
        x_t, x_0_hat = step_fn(x_t, t, dt, flow_model_fn, anchor_indices, x_0)
        if rigidity_forcing:
            x_t = procrustes.rigidify_and_interpolate(x_0_hat, condition, x_1, points_per_part, 1 - t + dt, t - dt)
        if return_trajectory:
            trajectory[step] = x_0_hat
            trajectory_x_t[step] = x_t
    if return_trajectory:
        return {"end_point_trajectory": trajectory, "trajectory": trajectory_x_t}
    return x_t.detach()


def euler_step(x_t: torch.Tensor, t: float, dt: float, flow_model_fn: Callable, anchor_indices: torch.Tensor, x_0: torch.Tensor):
    """Euler step for an opaque `flow_model_fn` (sampler.py:79-92): returns (x_t - dt v, x_t - v t)."""
    v = flow_model_fn(x_t, t)
    return x_t - dt * v, x_t - v * t


def get_sampler(sampler_name: str):
    step_fns = {"euler": euler_step}
    if sampler_name not in step_fns:
        raise ValueError(f"Unknown sampler: {sampler_name}. Available: {list(step_fns.keys())}")
    return partial(flow_sampler, step_fns[sampler_name])

#!/usr/bin/env python
"""Headline benchmark: registered pairs/s of the flow-matching registration sampler (BASELINE.json metric).

One "step" = one pass of the hot path over one batch: register `--pairs` synthetic two-view scenes
(4096 pts/view, 20 Euler flow steps of the rap_12 DiT, then the per-view Procrustes SE(3) fit) --
BASELINE.json configs[1] at N=1.  Scenes shard across ranks with no data-path collective (weak scaling);
each step ends with one NCCL all-gather of the recovered poses when N>1.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA path (this repo)
  python bench.py --impl reference [--gpus N] --steps K --warmup W  # reference algorithm on host cores (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "registered pairs/sec (N=4096 pts, 20 flow steps)"
VIEWS, PTS, FLOW_STEPS, LAYERS = 2, 4096, 20, 12
F_GEMM_PER_TOKEN = 2 * (179 * 512 + LAYERS * (2 * (512 * 1536 + 512 * 512) + 512 * 4096 + 2048 * 512) + 512 * 512 + 512 * 256 + 256 * 3)


def flops_per_pair() -> tuple[float, float]:
    """(GEMM, attention-core) algorithmic FLOPs per registered pair."""
    tokens = VIEWS * PTS
    f_attn_tok = LAYERS * 4 * 512 * (PTS + VIEWS * PTS)
    return tokens * FLOW_STEPS * F_GEMM_PER_TOKEN, tokens * FLOW_STEPS * f_attn_tok


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def env_int(k: str, d: int) -> int:
    return int(os.environ.get(k, d))


# --------------------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_sample(steps: int, warmup: int, budget_s: float = 150.0) -> dict:
    """Times the reference algorithm (oracle port, fp32, all host threads) on a bounded sample of the cfg2 workload:
    one DiT velocity evaluation (1 of the 20 flow steps) over `n` of the 12 layers, extrapolated linearly
    (layers are identical in cost), plus the final Procrustes fit.  Returns per-step seconds and the pairs/s estimate."""
    from oracle import rap_oracle as orc
    from rap_b200.synth import DiTConfig, make_batch, make_state_dict
    b = make_batch(1, VIEWS, PTS)
    ts = torch.full((1,), 0.5)

    def run(n_layers: int) -> float:
        sd = make_state_dict(DiTConfig(num_layers=n_layers))
        t0 = time.perf_counter()
        with torch.inference_mode():
            v = orc.dit_forward(sd, b.x_1, ts, b.cond, b.features, b.scales, b.anchor_indices, b.cu_seqlens_batch, b.cu_seqlens_part)
            orc.fit_transformations(b.cond, b.x_1 - v, b.points_per_part)
        return time.perf_counter() - t0

    # thread-count calibration on one layer (all hardware threads is often slower than one thread per physical core)
    ncpu = os.cpu_count() or 1
    best = None
    for threads in sorted({ncpu, max(ncpu // 2, 1), min(ncpu, 32)}, reverse=True):
        torch.set_num_threads(threads)
        run(1)
        t = run(1)
        if best is None or t < best[0]:
            best = (t, threads)
    t1, threads = best
    torch.set_num_threads(threads)
    n = 12
    for cand in (12, 6, 3, 1):
        n = cand
        if (steps + warmup) * t1 * cand <= budget_s:
            break
    for _ in range(warmup):
        run(n)
    times = [run(n) for _ in range(steps)]
    t_fixed = max(t1 - (run(2) - t1), 0.0) if n < 12 else 0.0     # embed + head + fit share, from the 1- vs 2-layer difference
    per_step = sum(times) / len(times)
    t_fwd12 = per_step if n == 12 else t_fixed + (per_step - t_fixed) * (12.0 / n)
    pairs_per_s = 1.0 / (FLOW_STEPS * t_fwd12)
    return {"value": pairs_per_s, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"1 of {FLOW_STEPS} flow steps (one DiT forward at 2x{PTS} pts, {n} of {LAYERS} layers, fp32 torch CPU, {threads} of {ncpu} host threads (fastest of a 3-way calibration)) "
                      f"+ Procrustes fit; {per_step:.2f} s/sample, extrapolated x{FLOW_STEPS} steps" + ("" if n == 12 else f" and x{12 // n} layers"),
            "seconds_per_sample": per_step, "steps": steps}


def run_reference(args) -> None:
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    r = cpu_sample(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "pairs/s", "n_gpus": max(world, args.gpus), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds_per_sample"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": f"pairwise registration, {VIEWS}x{PTS} pts, {FLOW_STEPS} flow steps, batch=1 (BASELINE configs[1])"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args) -> None:
    import torch.distributed as dist
    from rap_b200 import _lib
    from rap_b200.engine import Registrar
    from rap_b200.synth import DiTConfig, make_batch, make_state_dict
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B = args.steps, max(args.warmup, 3), args.pairs

    sd = make_state_dict(DiTConfig(num_layers=LAYERS), seed=0)
    reg = Registrar(sd, num_layers=LAYERS, device=dev, num_steps=FLOW_STEPS, rigidity_forcing=False)
    host = make_batch(B, VIEWS, PTS, first_scene=rank * B)            # each rank registers its own scenes (weak scaling)
    host = type(host)(**{k: v.pin_memory() for k, v in host.__dict__.items()})
    d = host.to(dev)
    plan = reg.plan_for(host.cu_seqlens_batch, host.points_per_part)
    from rap_b200.sharding import gather_poses

    def step_device():
        r = reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        if world > 1:   # the one collective of the path: NCCL all-gather of the recovered (R|t) of every rank
            gather_poses(r.rotations, r.translations)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(W):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    launches0 = _lib.lib.rap_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_device()
    e1.record()
    barrier()
    launches = _lib.lib.rap_launch_count() - launches0
    t_dev = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    clk = clocks.stop()
    assert _lib.device_error() == 0

    # end to end through the public API: pinned host inputs -> H2D -> sample + fit -> D2H of xyz and poses
    reg.register_host(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = reg.register_host(host)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    h2d = sum(int(getattr(host, k).numel() * getattr(host, k).element_size()) for k in ("cond", "features", "scales", "anchor_indices", "x_1"))
    d2h = int(sum(x.numel() * x.element_size() for x in (out.points, out.rotations, out.translations)))

    line = {"metric": METRIC, "value": world * B * K / t_dev, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"pairwise registration, {VIEWS}x{PTS} pts/view, {FLOW_STEPS} flow steps, rap_12 DiT (85.6M), batch={B} pair(s)/GPU "
                                   "(BASELINE configs[1]); random-init weights (seed 0)",
                       "parallelism": f"scene-sharded dp{world}", "pairs_per_gpu": B,
                       "l2": "no explicit flush: per-step working set ~290 MB (171 MB fp16 weights streamed every flow step + activations) exceeds the 126 MB L2"},
            "e2e": {"value": world * B * K / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clk}

    if rank == 0:
        pk = peaks()
        with _lib.profile() as prof:                         # separate pass of the same step with CUDA events around every launch
            reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        pr = prof.result
        f_gemm, f_attn = flops_per_pair()
        fm = pr.get("fmha_varlen", {"ms": float("nan"), "launches": 1})
        ach = f_attn * B / (fm["ms"] * 1e-3) / 1e12
        total_ms = sum(v["ms"] for v in pr.values())
        # dram__bytes_read+write per fmha launch from the committed ncu --set full capture of this workload (profiles/r01_v3_ncu_summary.md):
        # 25.4 MB read (= the 25.2 MB qkv buffer once), 0 written back (the 8.4 MB output stays in L2 for the out-projection GEMM).
        traffic = 25.42e6 if (B == 1 and world == 1) else None
        line["roofline"] = {"bound": "tensor", "kernel": "fmha_varlen (tcgen05 QK^T/PV, hd=64)", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                            "frac": ach / pk["tflops"], "traffic": traffic, "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                            "secondary_bound": "MUFU: one exp2 per score at 16/clk/SM = 1191 TFLOP/s-equivalent at 1965 MHz (tools/microbench/exp_rate.cu)",
                            "flops_per_launch": f_attn * B / max(fm["launches"], 1), "avg_launch_ms": fm["ms"] / max(fm["launches"], 1),
                            "share_of_step": fm["ms"] / total_ms if total_ms else None,
                            "path_frac": (f_gemm + f_attn) * world * B * K / t_dev / 1e12 / pk["tflops"]}
        # The profiling pass brackets every launch with an event pair and disables launch overlap; its per-class times exceed the timed
        # (unprofiled) step by a fixed cost per launch (tools/batch_breakdown.py: ~6.5 us).  Reported next to the raw figures, not instead.
        n_launch = sum(v["launches"] for v in pr.values())
        ev_us = max(0.0, (total_ms - t_dev / K * 1e3) / max(n_launch, 1) * 1e3)
        fm_net_ms = fm["ms"] - ev_us * 1e-3 * fm["launches"]
        line["roofline"]["event_overhead_us_per_launch"] = ev_us
        line["roofline"]["achieved_net_of_event_overhead"] = f_attn * B / (fm_net_ms * 1e-3) / 1e12 if fm_net_ms > 0 else None
        gemm_ms = sum(pr[k]["ms"] for k in pr if k.startswith("gemm"))
        line["kernel_breakdown_ms"] = {k: round(v["ms"], 3) for k, v in pr.items()}
        line["roofline"]["gemm_tflops"] = f_gemm * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        # once-per-scene pre-processing stage of the path: local features of both views, timed separately like the
        # reference does (it runs in demo.py before the sampler) and folded into an "including features" throughput
        try:
            from rap_b200.dataset_process import FeatureExtractor
            from rap_b200.synth import make_raw_cloud, make_spinnet_state_dict
            import numpy as np
            fx = FeatureExtractor(des_r=1.0, state_dict=make_spinnet_state_dict())
            clouds = [make_raw_cloud(10 + v, 60000, PTS) for v in range(VIEWS)]
            clouds = [(p.to(dev), k.to(dev), np.random.RandomState(v).choice(60000, 60000, replace=False)) for v, (p, k) in enumerate(clouds)]
            for p_, k_, perm in clouds:
                fx.extract_features(p_, k_, perm=perm)
            torch.cuda.synchronize()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for _ in range(3):
                for p_, k_, perm in clouds:
                    fx.extract_features(p_, k_, perm=perm)
            x1.record(); torch.cuda.synchronize()
            t_fx = x0.elapsed_time(x1) / 3e3
            line["extractor"] = {"ms_per_pair": t_fx * 1e3, "workload": f"{VIEWS} views x {PTS} keypoints, 60000-point raw clouds, des_r=1.0 m",
                                 "pairs_per_s_including_features": 1.0 / (t_dev / (K * B) + t_fx)}
        except Exception as e:  # pragma: no cover
            line["extractor"] = {"error": str(e)[:200]}
        # keypoint selection in front of the extractor, steady state, device-resident inputs
        try:
            import numpy as np
            from rap_b200.dataset_process.utils import dataset_utils as du
            from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
            g = torch.Generator(device=dev).manual_seed(3)
            raw = torch.randn(1_000_000, 3, device=dev, generator=g) * torch.tensor([40.0, 40.0, 6.0], device=dev)
            cloud = torch.randn(2, 100_000, 3, device=dev, generator=g) * torch.tensor([30.0, 30.0, 4.0], device=dev)
            lens, ks = torch.tensor([100_000, 100_000]), torch.tensor([5000, 5000])

            def timed(fn, n):
                fn(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / n
            line["keypoint_selection"] = {
                "voxel_downsample_ms": timed(lambda: du.voxel_down_sample_torch(raw, 0.5), 5), "voxel_workload": "1e6 points, 0.5 m voxels",
                "fps_ms": timed(lambda: sample_farthest_points(cloud, lengths=lens, K=ks, random_start_point=False), 3),
                "fps_workload": "2 clouds x 100000 points -> 5000 keypoints each (one 8-CTA cluster per cloud)"}
        except Exception as e:  # pragma: no cover
            line["keypoint_selection"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_sample(steps=1, warmup=0, budget_s=30.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rap_b200", choices=["rap_b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1, help="pairs per GPU per step (configs[1] is batch=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if env_int("WORLD_SIZE", 1) != args.gpus and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={env_int('WORLD_SIZE', 1)})")
        run_cuda(args)


if __name__ == "__main__":
    main()

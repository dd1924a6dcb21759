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


def workload_string(pairs_per_gpu: int) -> str:
    """One string for both arms (the driver compares `config.workload` of the CUDA and the reference line)."""
    return (f"pairwise registration, {VIEWS}x{PTS} pts/view, {FLOW_STEPS} flow steps, rap_12 DiT (85.6M), batch={pairs_per_gpu} pair(s)/GPU "
            f"(BASELINE configs[{1 if pairs_per_gpu == 1 else 3}] shape); random-init weights (seed 0)")


def default_pairs(world: int) -> int:
    """configs[1] (batch = 1) on one GPU; 8 pairs per GPU and step when the scenes of a batch are sharded over several GPUs
    (configs[3] is 32 per GPU; 8 keeps a step under a second so that K steps stay within the bench budget)."""
    return 1 if world == 1 else 8


# --------------------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_threads_calibrated(run1) -> int:
    """All hardware threads is often slower than one thread per physical core: pick the fastest of a 3-way calibration."""
    ncpu = os.cpu_count() or 1
    best = None
    for threads in sorted({ncpu, max(ncpu // 2, 1), min(ncpu, 32)}, reverse=True):
        torch.set_num_threads(threads)
        run1()
        t = run1()
        if best is None or t < best[0]:
            best = (t, threads)
    torch.set_num_threads(best[1])
    return best[1]


def cpu_sample(steps: int, warmup: int, budget_s: float = 60.0, full_pair: bool = False) -> dict:
    """Times the reference algorithm (oracle port, fp32, host threads) on the cfg2 workload.
    Bounded sample: one DiT velocity evaluation (1 of the 20 flow steps, all 12 layers) + the final Procrustes fit per step, as many
    steps as fit `budget_s`; pairs/s = 1 / (20 x that).  With `full_pair` (the `--impl reference` arm) ONE whole pair -- all 20 flow
    steps through the oracle's sampler and the fit, i.e. exactly the workload the CUDA arm times -- is run in addition and ITS wall
    time is the reported value, so the reference figure is measured, not extrapolated."""
    from oracle import rap_oracle as orc
    from rap_b200.synth import DiTConfig, make_batch, make_state_dict
    b = make_batch(1, VIEWS, PTS)
    ts = torch.full((1,), 0.5)
    sd = make_state_dict(DiTConfig(num_layers=LAYERS))
    sd1 = make_state_dict(DiTConfig(num_layers=1))

    def run(state) -> float:
        t0 = time.perf_counter()
        with torch.inference_mode():
            v = orc.dit_forward(state, b.x_1, ts, b.cond, b.features, b.scales, b.anchor_indices, b.cu_seqlens_batch, b.cu_seqlens_part)
            orc.fit_transformations(b.cond, b.x_1 - v, b.points_per_part)
        return time.perf_counter() - t0

    ncpu = os.cpu_count() or 1
    threads = cpu_threads_calibrated(lambda: run(sd1))
    t_first = run(sd)                                                     # also the warm-up
    n_steps = max(1, min(steps, int(budget_s / max(t_first, 1e-3))))
    for _ in range(min(warmup, 1)):
        run(sd)
    times = [run(sd) for _ in range(n_steps)]
    per_fwd = sum(times) / len(times)
    out = {"value": 1.0 / (FLOW_STEPS * per_fwd), "unit": "pairs/s", "cores": threads, "kind": "port", "seconds_per_sample": per_fwd, "steps": n_steps,
           "sample": f"{n_steps} x [1 of {FLOW_STEPS} flow steps: one DiT forward at {VIEWS}x{PTS} pts, all {LAYERS} layers, fp32 torch CPU + Procrustes fit], "
                     f"{threads} of {ncpu} host threads (fastest of a 3-way calibration); {per_fwd:.2f} s/sample, x{FLOW_STEPS} steps"}
    if full_pair:
        t0 = time.perf_counter()
        with torch.inference_mode():
            orc.register(sd, b.cond, b.features, b.scales, b.anchor_indices, b.points_per_part, b.cu_seqlens_batch, b.cu_seqlens_part, b.x_1,
                         FLOW_STEPS, False)
        t_pair = time.perf_counter() - t0
        out.update(value=1.0 / t_pair, seconds_per_pair=t_pair,
                   sample=f"ONE whole pair measured: all {FLOW_STEPS} flow steps of the oracle's sampler + Procrustes fit at {VIEWS}x{PTS} pts, fp32 torch CPU, "
                          f"{threads} of {ncpu} host threads: {t_pair:.1f} s (not extrapolated); per-forward samples beside it: {n_steps} x {per_fwd:.2f} s "
                          f"(x{FLOW_STEPS} = {FLOW_STEPS * per_fwd:.1f} s)")
    return out


def run_reference(args) -> None:
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    world = max(world, args.gpus)
    r = cpu_sample(args.steps, args.warmup, budget_s=60.0, full_pair=True)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds_per_pair"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_string(args.pairs or default_pairs(world))},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "a 'step' of this arm is one bounded CPU sample (see cpu_baseline.sample); ms_per_step is the measured whole pair"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- CUDA arm
def timed(fn, n: int, warm: int = 1) -> float:
    """Milliseconds per call: CUDA events on the current stream, `warm` untimed calls first, synchronised on both sides."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def own_kernel_comparators(dev) -> dict:
    """This library's attention / GEMM kernels timed in isolation on the shapes tools/comparators.py times the libraries on."""
    import ctypes as C
    from rap_b200 import _lib
    T = VIEWS * PTS
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(T, 3 * 512, device=dev, dtype=torch.float16, generator=g)
    out = torch.empty(T, 512, device=dev, dtype=torch.float16)
    res = {"attention": {}, "gemm": {}}
    for name, cu in (("per_view", [0, PTS, 2 * PTS]), ("per_scene", [0, T])):
        fl = sum(4.0 * 512 * (e - s) ** 2 for s, e in zip(cu[:-1], cu[1:]))
        row = {}
        # -1: the kernel the product launches (one item per CTA); 1 / 2: the opt-in persistent stream-K kernel, faster as an isolated launch
        # but not inside the sampler
        for key, variant in (("rap", -1), ("rap_persistent", 1), ("rap_persistent_contiguous", 2), ("rap_3tile_64key", 3)):
            ms = C.c_float(0)
            _lib.check(_lib.lib.rap_bench_fmha_variant(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, out.data_ptr(), 20, variant, C.byref(ms)), "rap_bench_fmha")
            row[key + "_ms"], row[key + "_tflops"] = ms.value, fl / ms.value / 1e9
        res["attention"][name] = row
    st = torch.cuda.current_stream().cuda_stream
    # (epilogue, tile N, K, N): the four DiT GEMM shapes with the epilogues the product fuses into them
    for name, (epi, bn, K, N) in {"qkv_512x1536 (+qk RMS-norm, fp16 out)": (2, 256, 512, 1536), "out_512x512 (+bias +fp32 residual)": (0, 256, 512, 512),
                                   "ff1_512x4096 (+bias, GEGLU, fp16 out)": (3, 256, 512, 4096), "ff2_2048x512 (+bias +fp32 residual)": (0, 256, 2048, 512)}.items():
        a = torch.randn(T, K, device=dev, dtype=torch.float16)
        w = torch.randn(N, K, device=dev, dtype=torch.float16) / K ** 0.5
        bias = torch.randn(N, device=dev)
        gamma = torch.ones(2 * 8 * 64, device=dev)
        h = torch.randn(T, N, device=dev) if epi == 0 else None
        o32 = torch.empty(T, N, device=dev) if epi == 0 else None
        o16 = torch.empty(T, N // 2 if epi == 3 else N, device=dev, dtype=torch.float16) if epi != 0 else None
        fn = lambda: _lib.check(_lib.lib.rap_test_gemm(epi, bn, T, N, K, a.data_ptr(), w.data_ptr(), bias.data_ptr() if epi != 2 else None,  # noqa: E731
                                                       _lib.ptr(h), _lib.ptr(o32), _lib.ptr(o16), gamma.data_ptr() if epi == 2 else None, None, st), "rap_test_gemm")
        ms = timed(fn, 30, 3)
        res["gemm"][name] = {"rap_ms": ms, "rap_tflops": 2.0 * T * N * K / ms / 1e9}
    return res


def comparators_block(dev, sd, d, plan, reg) -> dict:
    """Library kernels of the reference's GPU path on the exact shapes of the headline workload, next to this library's own kernels,
    same process, CUDA events."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import comparators as cmp
    T = VIEWS * PTS
    seg = {"per_view": [0, PTS, 2 * PTS], "per_scene": [0, T]}
    own = own_kernel_comparators(dev)
    att = cmp.attention_comparators(T, seg)
    if os.environ.get("RAP_BENCH_FLASHINFER", "0") == "1":   # JIT-compiles a CUTLASS FMHA with nvcc on first use (minutes on a fresh box): opt-in
        fi = cmp.flashinfer_comparator(T, seg, timeout_s=float(os.environ.get("RAP_BENCH_FLASHINFER_TIMEOUT", "150")))
        for k in seg:
            att[k].update(fi.get(k, {}) if "error" not in fi else {"flashinfer_error": fi["error"]})
    for k in seg:
        att[k].update(own["attention"][k])
        libs = {n: att[k][n + "_tflops"] for n in ("fa2", "cudnn", "torch_flash", "flashinfer") if n + "_tflops" in att[k]}
        if libs:
            best = max(libs, key=libs.get)
            att[k]["best_library"], att[k]["rap_vs_best_library"] = best, att[k]["rap_tflops"] / libs[best]
            if "fa2_tflops" in att[k]:
                att[k]["rap_vs_fa2"] = att[k]["rap_tflops"] / att[k]["fa2_tflops"]
    gm = cmp.gemm_comparators(T, {"qkv_512x1536": (512, 1536), "out_512x512": (512, 512), "ff1_512x4096": (512, 4096), "ff2_2048x512": (2048, 512)})
    for name, v in own["gemm"].items():
        key = name.split(" ")[0]
        gm[key].update({"rap_fused": name, **v, "rap_vs_cublas": v["rap_tflops"] / gm[key]["cublas_tflops"]})
    out = {"attention": att, "gemm": gm,
           "note": "cuBLAS rows time the bare fp16 GEMM + bias that the reference launches (its norm / GELU / residual kernels run separately); the rap rows "
                   "include the fused epilogue named in rap_fused.  Attention rows: identical (8192,3,8,64) fp16 qkv, segments = views / scene."}
    # flat copies of the headline ratios (the per-scene launch dominates the step)
    a = att["per_scene"]
    for n in ("fa2", "cudnn", "flashinfer", "torch_flash"):
        if n + "_tflops" in a:
            out[n + "_tflops"] = a[n + "_tflops"]
    out["rap_fmha_tflops"] = a["rap_tflops"]
    if "rap_vs_best_library" in a:
        out["fmha_vs_best_library"] = min(att[k]["rap_vs_best_library"] for k in seg if "rap_vs_best_library" in att[k])
    try:   # whole forward: the library-kernel stack the reference's GPU path amounts to (cuBLAS autocast + FA2 + eager elementwise) vs ours
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        lib_ms = cmp.library_forward_ms(sd_dev, d)
        ts = torch.full((1,), 0.5, device=dev)
        plan.set_condition(d.cond, d.features, d.scales, d.anchor_indices)
        own_ms = timed(lambda: plan.forward(d.x_1, ts), 10, 2)
        out["forward"] = {"library_stack_ms": lib_ms, "rap_ms": own_ms, "speedup": lib_ms / own_ms,
                          "library_stack": "torch fp16 autocast F.linear (cuBLAS) + flash_attn_varlen_qkvpacked_func (FA2) + eager LayerNorm/GELU/residual, TF32 head"}
    except Exception as e:  # pragma: no cover
        out["forward"] = {"error": repr(e)[:200]}
    return out


def extra_configs(dev, sd) -> dict:
    """One GPU's share of the other BASELINE.json configs through `Registrar` (device-resident inputs, CUDA events, one warm-up call)."""
    from rap_b200.engine import Registrar
    from rap_b200.synth import make_batch
    out = {}
    for key, (label, B, V, N, S, rigid, reps) in {
            "cfg1": ("configs[0]: 1 pair, 2 x 512 pts, 5 steps (launch-bound size: 690 kernels of a few microseconds)", 1, 2, 512, 5, False, 20),
            "cfg3": ("configs[2]: 1 scene, 4 views x 4096 pts, 20 steps, rigidity forcing on", 1, 4, 4096, 20, True, 2),
            "cfg4_share": ("configs[3] share of one GPU: 32 pairs x 2 x 4096 pts, 20 steps", 32, 2, 4096, 20, False, 1),
            "cfg5_share": ("configs[4] shape: 2 of one GPU's 8 pairs x 2 x 16384 pts, 50 steps", 2, 2, 16384, 50, False, 1)}.items():
        try:
            reg = Registrar(sd, num_layers=LAYERS, device=dev, num_steps=S, rigidity_forcing=rigid)
            d = make_batch(B, V, N).to(dev)
            plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
            ms = timed(lambda: reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1), reps, 3)   # 3 warm-ups: a plan captures its sampler graph on the third call
            flops = B * V * N * S * (F_GEMM_PER_TOKEN + LAYERS * 4 * 512 * (N + V * N))
            out[key] = {"workload": label, "ms_per_batch": ms, "scenes_per_s": B / (ms / 1e3), "path_tflops": flops / (ms / 1e3) / 1e12}
            del reg, plan, d
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            out[key] = {"error": repr(e)[:200]}
    return out


def run_cuda(args) -> None:
    import torch.distributed as dist
    from rap_b200 import _lib
    from rap_b200.engine import Registrar
    from rap_b200.sharding import shard_range, slice_scenes
    from rap_b200.synth import DiTConfig, make_batch, make_state_dict
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, B = args.steps, max(args.warmup, 3), args.pairs or default_pairs(world)      # W >= 3 also covers the one-off graph capture (third call of a plan)

    sd = make_state_dict(DiTConfig(num_layers=LAYERS), seed=0)
    reg = Registrar(sd, num_layers=LAYERS, device=dev, num_steps=FLOW_STEPS, rigidity_forcing=False)
    # The batch of world x B scenes; every rank registers its own block (weak scaling).  Layout tensors stay on the host, the data tensors
    # of the device-timed loop are resident in HBM, the end-to-end loop starts from pinned host memory.
    mine = shard_range(world * B, rank, world)
    host = make_batch(B, VIEWS, PTS, first_scene=mine.start)              # this rank's scenes (make_batch is deterministic per scene index)
    host = type(host)(**{k: v.pin_memory() for k, v in host.__dict__.items()})
    d = host.to(dev)
    plan = reg.plan_for(host.cu_seqlens_batch, host.points_per_part)
    if world > 1:
        # The sharded entry point takes the WHOLE batch (identical on every rank) and registers this rank's block of it.
        g_host = make_batch(world * B, VIEWS, PTS)
        g_host = type(g_host)(**{k: v.pin_memory() for k, v in g_host.__dict__.items()})
        g_dev = type(g_host)(**{k: (v if k in ("cu_seqlens_batch", "cu_seqlens_part", "points_per_part") else v.to(dev)) for k, v in g_host.__dict__.items()})
        assert torch.equal(slice_scenes(g_host, mine.start, mine.stop).cond, host.cond)

    def step_device():
        if world == 1:
            return reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        return reg.register_sharded(g_dev, wait=False)       # compute + ONE NCCL all-gather of the poses on the communication stream

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
        r = step_device()
    if world > 1:
        r.wait()
    barrier()
    clocks = ClockSampler(local)
    launches0 = _lib.lib.rap_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        r = step_device()
    if world > 1:
        r.wait()                                             # all-gathers run in order on one stream: the last one done = all done
    e1.record()
    barrier()
    launches = _lib.lib.rap_launch_count() - launches0
    t_dev = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    clk = clocks.stop()
    assert _lib.device_error() == 0
    if world > 1:
        assert r.poses.shape[0] == world * B and bool(torch.isfinite(r.poses).all())

    # end to end through the public API: pinned host inputs -> H2D -> sample + fit (-> all-gather of the poses) -> D2H of xyz and poses
    def step_e2e():
        return reg.register_host(host) if world == 1 else reg.register_sharded(g_host, wait=True)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = step_e2e()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    h2d = sum(int(getattr(host, k).numel() * getattr(host, k).element_size()) for k in ("cond", "features", "scales", "anchor_indices", "x_1"))
    d2h = int(sum(x.numel() * x.element_size() for x in ((out.points, out.rotations, out.translations) if world == 1 else (out.points, out.poses))))

    pk = peaks()
    f_gemm, f_attn = flops_per_pair()
    line = {"metric": METRIC, "value": world * B * K / t_dev, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_string(B), "parallelism": f"scene-sharded dp{world}", "pairs_per_gpu": B, "rigidity_forcing": False,
                       "api": "Registrar.register_device (value) / register_host (e2e)" if world == 1 else
                              "Registrar.register_sharded: shard by scene, one NCCL all-gather of the poses per step on a side stream (value: device-resident inputs, e2e: pinned host inputs)",
                       "l2": "no explicit flush: per-step working set ~290 MB (171 MB fp16 weights streamed every flow step + activations) exceeds the 126 MB L2"},
            "e2e": {"value": world * B * K / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes_pose_all_gather": world > 1},
            "gpu_launches": int(launches), "clocks": clk,
            "path_frac_per_gpu": (f_gemm + f_attn) * B * K / t_dev / 1e12 / pk["tflops"]}

    if rank == 0:
        with _lib.profile() as prof:                         # separate pass of the same per-GPU work with CUDA events around every launch
            reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        pr = prof.result
        fm = pr.get("fmha_varlen", {"ms": float("nan"), "launches": 1})
        ach = f_attn * B / (fm["ms"] * 1e-3) / 1e12
        total_ms = sum(v["ms"] for v in pr.values())
        line["roofline"] = {"bound": "tensor", "kernel": "fmha_varlen (tcgen05 QK^T/PV, hd=64)", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                            "frac": ach / pk["tflops"], "traffic": None,
                            "traffic_note": "not measurable in-run; dram__bytes per launch of this kernel from ncu --set full: profiles/r02_ncu_summary.md",
                            "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                            "secondary_bound": "MUFU: one exp2 per score at 16/clk/SM = 1191 TFLOP/s-equivalent at 1965 MHz (tools/microbench/exp_rate.cu)",
                            "flops_per_launch": f_attn * B / max(fm["launches"], 1), "avg_launch_ms": fm["ms"] / max(fm["launches"], 1),
                            "share_of_step": fm["ms"] / total_ms if total_ms else None,
                            "path_frac": line["path_frac_per_gpu"]}
        # The profiling pass brackets every launch with an event pair and disables launch overlap; its per-class times exceed the timed
        # (unprofiled) step by a fixed cost per launch (tools/batch_breakdown.py).  Reported next to the raw figures, not instead.
        n_launch = sum(v["launches"] for v in pr.values())
        ev_us = max(0.0, (total_ms - t_dev / K * 1e3) / max(n_launch, 1) * 1e3)
        fm_net_ms = fm["ms"] - ev_us * 1e-3 * fm["launches"]
        line["roofline"]["event_overhead_us_per_launch"] = ev_us
        line["roofline"]["achieved_net_of_event_overhead"] = f_attn * B / (fm_net_ms * 1e-3) / 1e12 if fm_net_ms > 0 else None
        gemm_ms = sum(pr[k]["ms"] for k in pr if k.startswith("gemm"))
        line["kernel_breakdown_ms"] = {k: round(v["ms"], 3) for k, v in pr.items()}
        line["roofline"]["gemm_tflops"] = f_gemm * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
    if rank == 0 and world == 1 and not args.quick:
        try:
            line["comparators"] = comparators_block(dev, sd, d, plan, reg)
        except Exception as e:  # pragma: no cover
            line["comparators"] = {"error": repr(e)[:300]}
        extra = {}
        try:   # the reference's shipped default is rigidity_forcing: true (config/RAP_inference.yaml:58): same workload with it on
            reg_r = Registrar(sd, num_layers=LAYERS, device=dev, num_steps=FLOW_STEPS, rigidity_forcing=True)
            plan_r = reg_r.plan_for(host.cu_seqlens_batch, host.points_per_part)
            ms = timed(lambda: reg_r.register_device(plan_r, d.cond, d.features, d.scales, d.anchor_indices, d.x_1), max(K // 2, 3), 3)
            extra["headline_with_rigidity_forcing"] = {"pairs_per_s": B / (ms / 1e3), "ms_per_step": ms}
            del reg_r, plan_r
            # strict weak-scaling baseline for the N>1 runs (which register 8 pairs per GPU and step)
            d8 = make_batch(8, VIEWS, PTS).to(dev)
            plan8 = reg.plan_for(d8.cu_seqlens_batch.cpu(), d8.points_per_part.cpu())
            ms8 = timed(lambda: reg.register_device(plan8, d8.cond, d8.features, d8.scales, d8.anchor_indices, d8.x_1), 3, 3)
            extra["n1_at_8_pairs_per_gpu"] = {"pairs_per_s": 8 / (ms8 / 1e3), "ms_per_step": ms8,
                                              "note": "per-GPU work of the N>1 runs on one GPU: divide their value by N x this for the strict weak-scaling efficiency"}
            del d8, plan8
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            extra["error"] = repr(e)[:200]
        extra["configs"] = extra_configs(dev, sd)
        line["extra"] = extra
        # once-per-scene pre-processing stage of the path: local features of both views, timed separately like the
        # reference does (it runs in demo.py before the sampler) and folded into an "including features" throughput
        try:
            from rap_b200.dataset_process import FeatureExtractor
            from rap_b200.synth import make_raw_cloud, make_spinnet_state_dict
            import numpy as np
            fx = FeatureExtractor(des_r=1.0, state_dict=make_spinnet_state_dict())
            clouds = [make_raw_cloud(10 + v, 60000, PTS) for v in range(VIEWS)]
            clouds = [(p.to(dev), k.to(dev), np.random.RandomState(v).choice(60000, 60000, replace=False)) for v, (p, k) in enumerate(clouds)]

            def fx_all():
                for p_, k_, perm in clouds:
                    fx.extract_features(p_, k_, perm=perm)
            t_fx = timed(fx_all, 3, 1) / 1e3
            line["extractor"] = {"ms_per_pair": t_fx * 1e3, "workload": f"{VIEWS} views x {PTS} keypoints, 60000-point raw clouds, des_r=1.0 m",
                                 "pairs_per_s_including_features": 1.0 / (t_dev / (K * B) + t_fx)}
        except Exception as e:  # pragma: no cover
            line["extractor"] = {"error": str(e)[:200]}
        # keypoint selection in front of the extractor, steady state, device-resident inputs
        try:
            from rap_b200.dataset_process.utils import dataset_utils as du
            from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
            g = torch.Generator(device=dev).manual_seed(3)
            raw = torch.randn(1_000_000, 3, device=dev, generator=g) * torch.tensor([40.0, 40.0, 6.0], device=dev)
            cloud = torch.randn(2, 100_000, 3, device=dev, generator=g) * torch.tensor([30.0, 30.0, 4.0], device=dev)
            lens, ks = torch.tensor([100_000, 100_000]), torch.tensor([5000, 5000])
            line["keypoint_selection"] = {
                "voxel_downsample_ms": timed(lambda: du.voxel_down_sample_torch(raw, 0.5), 5), "voxel_workload": "1e6 points, 0.5 m voxels",
                "fps_ms": timed(lambda: sample_farthest_points(cloud, lengths=lens, K=ks, random_start_point=False), 3),
                "fps_workload": "2 clouds x 100000 points -> 5000 keypoints each (one 8-CTA cluster per cloud)"}
        except Exception as e:  # pragma: no cover
            line["keypoint_selection"] = {"error": str(e)[:200]}
        if not args.no_cpu_baseline:
            cb = cpu_sample(steps=3, warmup=0, budget_s=15.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
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
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: 1 = configs[1] on one GPU, 8 when N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the comparator / extra-config / pre-processing sections (headline numbers only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if env_int("WORLD_SIZE", 1) != args.gpus and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={env_int('WORLD_SIZE', 1)})")
        run_cuda(args)


if __name__ == "__main__":
    main()

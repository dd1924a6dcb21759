"""Runs the per-GPU share of every BASELINE.json config through the public API and prints throughput (sanity + DESIGN.md table)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200 import _lib
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict
sd = make_state_dict(DiTConfig(num_layers=12))
CFG = [("cfg2 pair 2x4096 s20 b1", 1, 2, 4096, 20, False, 3),
       ("cfg3 4 views x4096 s20 rigid b1", 1, 4, 4096, 20, True, 3),
       ("cfg4 share: 32 pairs 2x4096 s20", 32, 2, 4096, 20, False, 1),
       ("cfg5 share: 8 pairs 2x16384 s50", 8, 2, 16384, 50, False, 1)]
only = os.environ.get("ONLY")
for name, B, V, N, S, rigid, reps in CFG:
    if only and only not in name: continue
    reg = Registrar(sd, num_layers=12, device="cuda", num_steps=S, rigidity_forcing=rigid)
    d = make_batch(B, V, N).to("cuda")
    plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
    r = reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)     # warm-up
    torch.cuda.synchronize(); assert _lib.device_error() == 0 and torch.isfinite(r.points).all() and torch.isfinite(r.rotations).all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / reps
    tokens = B * V * N
    flops = tokens * S * (126.8e6 + 12 * 4 * 512 * (N + V * N))
    det = torch.det(r.rotations[:, :V]).abs()
    print(f"{name}: {t*1e3:.1f} ms/batch, {B/t:.2f} scenes/s per GPU, {flops/t/1e12:.0f} TFLOP/s, |det R| in [{det.min():.4f},{det.max():.4f}], mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB torch", flush=True)
    del reg, plan, d, r
    torch.cuda.empty_cache()

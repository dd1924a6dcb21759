"""Per-kernel-class time per pair at 1 / 8 / 32 pairs per batch (2x4096 pts, 20 steps): where the batch-size dependence comes from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200 import _lib
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict
sd = make_state_dict(DiTConfig(num_layers=12))
for B in (1, 8, 32):
    reg = Registrar(sd, num_layers=12, device="cuda", num_steps=20)
    d = make_batch(B, 2, 4096).to("cuda")
    plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
    reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1); e1.record(); torch.cuda.synchronize()
    with _lib.profile() as pr:
        reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        torch.cuda.synchronize()
    per = {k: round(v["ms"] / B, 2) for k, v in pr.result.items()}
    print(f"B={B}: {e0.elapsed_time(e1)/B:.1f} ms/pair unprofiled; per-pair kernel ms: {per}", flush=True)
    del reg, plan, d
    torch.cuda.empty_cache()

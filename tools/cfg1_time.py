"""configs[0] (1 pair, 2 x 512 points, 5 flow steps: the launch-latency-bound size) through Registrar, ms per registration; environment knobs apply
(e.g. RAP_GEMM_TMA16=0 | 1).  Usage: python tools/cfg1_time.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda")
sd = make_state_dict(DiTConfig(num_layers=bench.LAYERS), seed=0)
reg = Registrar(sd, num_layers=bench.LAYERS, device=dev, num_steps=5, rigidity_forcing=False)
d = make_batch(1, 2, 512).to(dev)
plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
fn = lambda: reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
out = []
for _ in range(3):
    out.append(bench.timed(fn, reps, 3))
print("cfg1 (2 x 512 pts, 5 steps) %s: %s ms per registration" % (" ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("RAP_")), " / ".join(f"{m:.3f}" for m in out)), flush=True)

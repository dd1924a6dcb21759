"""One registered pair (BASELINE configs[1]) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict
steps = int(os.environ.get("FLOW_STEPS", 20))
reg = Registrar(make_state_dict(DiTConfig(num_layers=12)), num_layers=12, device="cuda", num_steps=steps)
d = make_batch(1, 2, 4096).to("cuda")
plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
for _ in range(2):
    reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""Stage-by-stage run of one small forward with a sync after every C call (debug aid for gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200 import _lib
from rap_b200.engine import DiTHandle, Plan
from rap_b200.synth import DiTConfig, make_batch, make_state_dict

def sync(tag):
    torch.cuda.synchronize()
    print("ok:", tag, "deverr", _lib.device_error(), flush=True)

L = int(os.environ.get("L", 2))
sd = {k: v.cuda() for k, v in make_state_dict(DiTConfig(num_layers=L)).items()}
h = DiTHandle(512, L, 8, 32); h.load_state_dict(sd); sync("weights")
b = make_batch(2, 3, [200, 136, 72], max_parts=5)
plan = Plan(h, b.cu_seqlens_batch.numpy(), b.points_per_part.numpy(), 64); sync("plan")
d = b.to("cuda")
plan.set_condition(d.cond, d.features, d.scales, d.anchor_indices); sync("set_condition")
sot = torch.empty(plan.T, dtype=torch.int32, device="cuda")
_lib.check(_lib.lib.rap_debug_read(plan._h, 0, sot.data_ptr(), plan.T * 4)); print("scene_of_tok", sot[:4].tolist(), sot[-4:].tolist(), int(sot.min()), int(sot.max()))
hh = torch.randn(plan.T, 512, device="cuda"); o16 = torch.empty(plan.T, 512, device="cuda", dtype=torch.float16)
tab = torch.randn(2, 4, 1024, device="cuda")
_lib.check(_lib.lib.rap_test_ln_modulate(hh.data_ptr(), o16.data_ptr(), plan.T, tab[0,1].data_ptr(), tab[0,1,512:].data_ptr(), 4096, sot.data_ptr(), 1.0, _lib.stream_ptr())); sync("ln standalone")
ref = torch.nn.functional.layer_norm(hh, (512,)) * (1 + tab[sot.long(), 1, :512]) + tab[sot.long(), 1, 512:]
print("ln err", (o16.float() - ref).abs().max().item())
try:
    v = plan.forward(d.x_1, torch.tensor([0.9, 0.35], device="cuda")); sync("forward")
finally:
    _lib.check(_lib.lib.rap_debug_read(plan._h, 0, sot.data_ptr(), plan.T * 4)); print("scene_of_tok after", sot[40:60].tolist(), int(sot.min()), int(sot.max()))
    print("bad idx", (sot.long() > 1).nonzero().flatten()[:20].tolist(), [hex(x) for x in sot[(sot.long() > 1) | (sot.long() < 0)][:8].tolist()])
print(v[:3])
out = plan.sample(d.x_1, 3, True); sync("sample")
R, t = plan.fit(plan.cond, out["end_point_trajectory"][-1]); sync("fit")
print(R[0, 0], t[0, 0])

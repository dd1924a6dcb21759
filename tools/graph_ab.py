"""Whole-sampler CUDA graph vs stream launches (RAP_GRAPH=1 / 0) at several batch sizes: milliseconds per registered batch, host-timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict
sd = make_state_dict(DiTConfig(num_layers=12))
for (B, V, N, S) in ((1, 2, 512, 5), (1, 2, 1024, 20), (4, 2, 512, 20), (1, 2, 2048, 20), (1, 2, 4096, 20)):
    res = {}
    for mode in ("0", "1"):
        os.environ["RAP_GRAPH"] = mode
        reg = Registrar(sd, num_layers=12, device="cuda", num_steps=S)
        d = make_batch(B, V, N).to("cuda")
        plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
        f = lambda: reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
        for _ in range(3):
            f()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 20 if N < 4096 else 5
        for _ in range(n):
            f()
        torch.cuda.synchronize(); res[mode] = (time.perf_counter() - t0) / n * 1e3
    print("B=%d V=%d N=%d S=%d (T=%d): stream launches %.2f ms, graph %.2f ms per batch" % (B, V, N, S, B * V * N, res["0"], res["1"]), flush=True)

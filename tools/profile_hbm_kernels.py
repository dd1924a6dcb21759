"""The HBM-/latency-bound kernels of the path between cudaProfilerStart/Stop for `ncu --profile-from-start off --set full`:
one flow step of a registered pair with rigidity forcing (ln_modulate, head_final, posenc, procrustes), one extractor view
(patch_kernel = ball query, voxel_feature_kernel, pool_kernel) and one batched FPS call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rap_b200.dataset_process import FeatureExtractor
from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_raw_cloud, make_spinnet_state_dict, make_state_dict
reg = Registrar(make_state_dict(DiTConfig(num_layers=1)), num_layers=1, device="cuda", num_steps=1, rigidity_forcing=True)   # 1 layer: 3 LayerNorm launches
d = make_batch(1, 2, 4096).to("cuda")
plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
fx = FeatureExtractor(des_r=1.0, state_dict=make_spinnet_state_dict())
p_, k_ = make_raw_cloud(10, 60000, 4096)
p_, k_, perm = p_.cuda(), k_.cuda(), np.random.RandomState(0).choice(60000, 60000, replace=False)
g = torch.Generator(device="cuda").manual_seed(3)
cloud = torch.randn(2, 100_000, 3, device="cuda", generator=g) * torch.tensor([30.0, 30.0, 4.0], device="cuda")
lens, ks = torch.tensor([100_000, 100_000]), torch.tensor([5000, 5000])


def once():
    reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
    fx.extract_features(p_, k_, perm=perm)
    sample_farthest_points(cloud, lengths=lens, K=ks, random_start_point=False)


once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

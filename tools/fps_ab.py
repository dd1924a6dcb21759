import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
g = torch.Generator(device="cuda").manual_seed(3)
cloud = torch.randn(2, 100_000, 3, device="cuda", generator=g) * torch.tensor([30.0, 30.0, 4.0], device="cuda")
lens, ks = torch.tensor([100_000, 100_000]), torch.tensor([5000, 5000])
def run():
    return sample_farthest_points(cloud, lengths=lens, K=ks, random_start_point=False)[1]
ref = run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): idx = run()
b.record(); torch.cuda.synchronize()
print(os.environ.get("RAP_FPS_CLUSTER16", "1"), "fps ms", a.elapsed_time(b) / 3, "checksum", int(idx.sum()))

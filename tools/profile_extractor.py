"""One extractor call (2 views x 4096 keypoints, 60000-point clouds) between cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rap_b200.dataset_process import FeatureExtractor
from rap_b200.synth import make_raw_cloud, make_spinnet_state_dict
fx = FeatureExtractor(des_r=1.0, state_dict=make_spinnet_state_dict())
clouds = [make_raw_cloud(10 + v, 60000, 4096) for v in range(2)]
clouds = [(p.cuda(), k.cuda(), np.random.RandomState(v).choice(60000, 60000, replace=False)) for v, (p, k) in enumerate(clouds)]
for p_, k_, perm in clouds:
    fx.extract_features(p_, k_, perm=perm)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for p_, k_, perm in clouds:
    fx.extract_features(p_, k_, perm=perm)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

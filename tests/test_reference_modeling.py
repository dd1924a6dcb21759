"""The reference's own `modeling.py` (`RectifiedPointFlow`) driven through `rap_b200.install()` -- the "drops into sample.py unchanged"
claim of INTEGRATION.md, checked against the reference's code instead of a re-typed stand-in.

`/root/reference` exists only in the build container (no GPU) and never on the GPU box, so:
  * the CPU test runs HERE: the verbatim `sample_rectified_flow` builds its closure, hands it to OUR `get_sampler`, and the fused route must
    recognise it (free-variable names, `self.flow_model`), pass the right tensors and reproduce the closure's side effects -- the CUDA plan is
    replaced by a recording fake, everything in front of it is the real code path;
  * the GPU test (full numerical comparison with the cfg1 golden through `sample_rectified_flow` + the `test_step` pose fit) runs wherever
    a GPU and the reference are both present, and is skipped -- not deleted -- elsewhere.
"""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden import CASES
from oracle.ref_loader import load_reference_modeling, reference_available, unload_reference
from rap_b200.synth import DiTConfig, make_batch, make_state_dict

needs_reference = pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the build container")


def _data_dict(b):
    return {"pointclouds_gt": b.gt, "pointclouds": b.cond, "features": b.features, "scales": b.scales, "anchor_indices": b.anchor_indices,
            "points_per_part": b.points_per_part, "cu_seqlens": b.cu_seqlens_batch}


def _build(ns, L, steps, rigid, device):
    from rap_b200.rectified_point_flow.flow_model import PointCloudDiT
    assert ns.flow_model.PointCloudDiT is PointCloudDiT
    fm = PointCloudDiT(in_dim=0, out_dim=3, embed_dim=512, num_layers=L, num_heads=8, local_feat_dim=32)
    fm.load_state_dict(make_state_dict(DiTConfig(num_layers=L), seed=0), strict=True)
    m = ns.modeling.RectifiedPointFlow(feature_extractor=None, flow_model=fm.to(device), optimizer=object(), encoder_on=False,
                                       inference_sampling_steps=steps, inference_sampler="euler", rigidity_forcing=rigid,
                                       timestep_sampling="uniform")
    return m.eval()


@needs_reference
def test_reference_modeling_binds_to_the_mirrors_and_takes_the_fused_route(monkeypatch):
    from rap_b200.rectified_point_flow import procrustes, sampler
    from rap_b200.rectified_point_flow.flow_model import PointCloudDiT
    try:
        ns = load_reference_modeling()
        # modeling.py:14-15 bound OUR functions
        assert ns.modeling.get_sampler is sampler.get_sampler and ns.modeling.fit_transformations is procrustes.fit_transformations
        S, rigid = 5, True
        b = make_batch(2, 3, [40, 24, 16], max_parts=6)
        model = _build(ns, 2, S, rigid, "cpu")
        seen = {}

        class FakePlan:
            max_timesteps = 64

            def set_condition(self, cond, feats, scales, anchor):
                seen["cond"] = (cond, feats, scales, anchor)

            def sample(self, x_1, num_steps, rigidity_forcing, return_trajectory=True, return_features=False):
                seen["sample"] = (x_1, num_steps, rigidity_forcing, return_trajectory, return_features)
                T = x_1.shape[0]
                return {"end_point_trajectory": torch.arange(num_steps * T * 3, dtype=torch.float32).view(num_steps, T, 3),
                        "trajectory": torch.ones(num_steps, T, 3), "x_final": torch.zeros(T, 3), "transformer_features": torch.full((T, 512), 7.0)}

        def fake_plan(self, cu_b, cu_p, points_per_part=None, max_timesteps=64):
            seen["plan"] = (cu_b, cu_p, points_per_part, max_timesteps)
            return FakePlan()
        monkeypatch.setattr(PointCloudDiT, "plan", fake_plan)
        out = model.sample_rectified_flow(_data_dict(b), None, x_1=b.x_1, return_tarjectory=True, return_transformer_features=True)
        # the fused route ran exactly once with the tensors `_prepare_data` (modeling.py:203-231) produced
        assert seen["sample"][1:] == (S, rigid, True, True) and seen["sample"][0] is b.x_1
        cu_b, cu_p, ppp, _ = seen["plan"]
        assert cu_b.dtype == torch.int32 and cu_b.tolist() == b.cu_seqlens_batch.tolist() and cu_p.tolist() == b.cu_seqlens_part.tolist()
        assert ppp is b.points_per_part and seen["cond"][0] is b.cond and seen["cond"][1] is b.features
        # return structure of modeling.py:722-729 and the closure's side effect (`nonlocal transformer_features`)
        assert set(out) == {"trajectory", "transformer_features"} and set(out["trajectory"]) == {"end_point_trajectory", "trajectory"}
        assert out["trajectory"]["end_point_trajectory"].shape == (S, b.cond.shape[0], 3)
        assert float(out["transformer_features"].mean()) == 7.0
    finally:
        unload_reference()


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1_2x512_s5", "cfg1_2x512_s5_rigid"])
def test_reference_modeling_sample_and_fit_match_the_golden(golden_dir, name):
    """`RectifiedPointFlow.sample_rectified_flow(..., x_1=...)` (modeling.py:633-741) + the `test_step` pose fit (modeling.py:389) of the
    UNMODIFIED reference class over a rap_b200 `PointCloudDiT`, compared with the golden the all-reference run produced."""
    assert torch.cuda.is_available()
    L, B, V, N, S, rigid = CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    try:
        ns = load_reference_modeling()
        b = make_batch(B, V, N, max_parts=V + 2).to("cuda")
        model = _build(ns, L, S, rigid, "cuda")
        out = model.sample_rectified_flow(_data_dict(b), None, x_1=b.x_1, return_tarjectory=True, return_transformer_features=True)
        ep = out["trajectory"]["end_point_trajectory"]
        assert out["transformer_features"].shape == (b.cond.shape[0], 512)
        d = (ep.cpu() - torch.from_numpy(g["end_point_trajectory"])).abs()
        assert d.mean().item() < 2e-2 and d.max().item() < 0.25         # random-weight network: see the k=2 rule in test_gpu_parity.py
        R, t = ns.modeling.fit_transformations(b.cond, ep[-1], b.points_per_part, b.cu_seqlens_batch)
        assert R.shape == (B, V + 2, 3, 3) and torch.isfinite(R).all() and torch.isfinite(t).all()
    finally:
        unload_reference()

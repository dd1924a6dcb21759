"""Pins oracle/rap_oracle.py (the CPU restatement) against the golden vectors produced by the
reference's own sources executed verbatim (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import rap_oracle as orc
from oracle.make_golden import BIG_CASES, CASES, case_params, input_checksum
from rap_b200.synth import DiTConfig, make_batch, make_state_dict


def _load(golden_dir, name):
    L, B, V, N, S, rigid, P, _keep = case_params(name)
    cfg = DiTConfig(num_layers=L)
    sd = make_state_dict(cfg, seed=0)
    batch = make_batch(B, V, N, max_parts=P)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    assert abs(float(g["input_checksum"]) - input_checksum(batch, sd)) < 1e-6, "synthetic inputs drifted; regenerate goldens"
    return cfg, sd, batch, g, (S, rigid)


@pytest.mark.parametrize("name", list(CASES))
def test_sampler_matches_reference(golden_dir, name):
    cfg, sd, b, g, (S, rigid) = _load(golden_dir, name)
    out = orc.register(sd, b.cond, b.features, b.scales, b.anchor_indices, b.points_per_part, b.cu_seqlens_batch,
                       b.cu_seqlens_part, b.x_1, S, rigid, heads=cfg.num_heads)
    # fp32 vs fp32 with different summation orders: tight absolute tolerance in normalised scene units
    np.testing.assert_allclose(out["end_point_trajectory"].numpy(), g["end_point_trajectory"], atol=1e-3, rtol=0)
    np.testing.assert_allclose(out["trajectory"].numpy(), g["trajectory"], atol=1e-3, rtol=0)
    # poses compared by their action on cond (entry-wise R is ill-conditioned with random weights)
    Rg, tg = torch.from_numpy(g["rotations"]), torch.from_numpy(g["translations"])
    for bb, p, s, e in orc.part_offsets(b.points_per_part):
        a = b.cond[s:e] @ out["rotations"][bb, p].t() + out["translations"][bb, p]
        r = b.cond[s:e] @ Rg[bb, p].t() + tg[bb, p]
        assert (a - r).abs().max() < 5e-4
    # empty part slots stay zero (procrustes.py:71-76)
    empty = b.points_per_part == 0
    assert out["rotations"][empty].abs().max() == 0 and out["translations"][empty].abs().max() == 0


@pytest.mark.parametrize("name", ["ragged_b2_v3_L2_s3_rigid", "cfg1_2x512_s5"])
def test_forward_matches_reference(golden_dir, name):
    cfg, sd, b, g, _ = _load(golden_dir, name)
    ts = torch.from_numpy(g["fwd_timesteps"])
    out = orc.dit_forward(sd, b.x_1, ts, b.cond, b.features, b.scales, b.anchor_indices, b.cu_seqlens_batch,
                          b.cu_seqlens_part, heads=cfg.num_heads, return_transformer_features=True)
    np.testing.assert_allclose(out["velocity"].numpy(), g["fwd_velocity"], atol=5e-5, rtol=1e-4)
    np.testing.assert_allclose(out["transformer_features"][:8].numpy(), g["fwd_features_head"], atol=2e-4, rtol=1e-4)
    assert abs(float(out["transformer_features"].double().abs().sum()) / float(g["fwd_features_checksum"]) - 1) < 1e-5


@pytest.mark.parametrize("name", ["small_L2_s70", "cfg1_2x512_s5_p512"])
def test_sampler_matches_reference_compact_cases(golden_dir, name):
    """More than 64 flow steps, and the reference's `max_parts: 512` padding (config/RAP_inference.yaml:33): the compact goldens
    keep a few trajectory steps (`kept_steps`) plus the poses."""
    cfg, sd, b, g, (S, rigid) = _load(golden_dir, name)
    out = orc.register(sd, b.cond, b.features, b.scales, b.anchor_indices, b.points_per_part, b.cu_seqlens_batch,
                       b.cu_seqlens_part, b.x_1, S, rigid, heads=cfg.num_heads)
    keep = g["kept_steps"].tolist()
    tol = 1e-3 if S <= 20 else 4e-3          # fp32 summation-order noise grows along 70 steps of a random-weight network
    np.testing.assert_allclose(out["end_point_trajectory"][keep].numpy(), g["end_point_trajectory"], atol=tol, rtol=0)
    np.testing.assert_allclose(out["trajectory"][keep].numpy(), g["trajectory"], atol=tol, rtol=0)
    np.testing.assert_allclose(np.array([1 - i / S for i in range(S)]), g["call_times"], atol=1e-12)
    Rg, tg = torch.from_numpy(g["rotations"]), torch.from_numpy(g["translations"])
    assert out["rotations"].shape == Rg.shape
    for bb, p, s, e in orc.part_offsets(b.points_per_part):
        a = b.cond[s:e] @ out["rotations"][bb, p].t() + out["translations"][bb, p]
        r = b.cond[s:e] @ Rg[bb, p].t() + tg[bb, p]
        assert (a - r).abs().max() < 5 * tol
    empty = b.points_per_part == 0
    assert out["rotations"][empty].abs().max() == 0 and float(np.abs(g["rotations"][empty.numpy()]).max()) == 0


@pytest.mark.parametrize("name", ["cfg2_2x4096_s20", "cfg3_4x4096_s3_rigid"])
def test_forward_matches_reference_at_baseline_sizes(golden_dir, name):
    """One oracle forward at the BASELINE.json sizes (configs[1]: 2 x 4096, configs[2]: 4 x 4096) against the reference run verbatim;
    the 20-step cfg2 trajectory itself is checked on the GPU (tests/test_gpu_parity.py), where it takes milliseconds."""
    cfg, sd, b, g, _ = _load(golden_dir, name)
    ts = torch.from_numpy(g["fwd_timesteps"])
    with torch.inference_mode():
        out = orc.dit_forward(sd, b.x_1, ts, b.cond, b.features, b.scales, b.anchor_indices, b.cu_seqlens_batch,
                              b.cu_seqlens_part, heads=cfg.num_heads, return_transformer_features=True)
    np.testing.assert_allclose(out["velocity"].numpy(), g["fwd_velocity"], atol=1e-4, rtol=1e-4)
    assert abs(float(out["transformer_features"].double().abs().sum()) / float(g["fwd_features_checksum"]) - 1) < 1e-5


def test_rigidify_matches_reference(golden_dir):
    cfg, sd, b, g, _ = _load(golden_dir, "ragged_b2_v3_L2_s3_rigid")
    pred = torch.from_numpy(g["end_point_trajectory"][-1])
    out = orc.rigidify(pred, b.cond, b.points_per_part)
    np.testing.assert_allclose(out.numpy(), g["rigidified"], atol=2e-5, rtol=0)


def test_procrustes_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, "procrustes_kat.npz"))
    for i in range(5):
        src, tgt = torch.from_numpy(g[f"src{i}"]), torch.from_numpy(g[f"tgt{i}"])
        R, t = orc.solve_procrustes(src, tgt)
        np.testing.assert_allclose(R.numpy(), g[f"R{i}"], atol=2e-5)
        np.testing.assert_allclose(t.numpy(), g[f"t{i}"], atol=2e-5)
        assert abs(float(torch.det(R)) - 1) < 1e-4


def test_sampler_analytic_one_step():
    """v = (x - target)/t makes x0_hat exact after one call; checks the (1-t+dt)/(t-dt) re-interpolation
    weights and that the end-point trajectory is the un-rigidified x0_hat (sampler.py:55-65)."""
    torch.manual_seed(0)
    cond = torch.randn(40, 3)
    ppp = torch.tensor([[25, 15, 0]])
    q, _ = torch.linalg.qr(torch.randn(3, 3))
    q = q * torch.sign(torch.det(q))
    target = torch.cat([cond[:25] @ q.t() + 0.3, cond[25:] + 1.0])
    x1 = torch.randn(40, 3)
    ep, tr = orc.flow_sampler(lambda x, t: (x - target) / t, x1, cond, ppp, 4, True)
    assert (ep[0] - target).abs().max() < 1e-5
    assert (tr[0] - (target * 0.25 + x1 * 0.75)).abs().max() < 1e-4   # t=1, dt=.25 -> weights (.25,.75)
    assert (ep[-1] - target).abs().max() < 1e-4

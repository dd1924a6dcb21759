"""Pins oracle/spinnet_oracle.py against golden vectors from the reference MiniSpinNet executed verbatim."""
import os

import numpy as np
import pytest
import torch

from oracle import spinnet_oracle as so
from oracle.make_golden import SPINNET_CASES
from rap_b200.synth import make_raw_cloud, make_spinnet_state_dict


@pytest.mark.parametrize("name", list(SPINNET_CASES))
def test_descriptor_matches_reference(golden_dir, name):
    idx, n, k, des_r, seed = SPINNET_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    pts, kp = make_raw_cloud(idx, n, k)
    perm = np.random.RandomState(seed).choice(n, n, replace=False)       # what np.random.seed(seed); np.random.choice(N, N, False) yields
    out = so.extract_features(make_spinnet_state_dict(), pts, kp, des_r, perm)
    np.testing.assert_array_equal(out["patches"][:4].numpy(), g["patches_head"])          # index work: bit exact
    assert abs(float(out["patches"].double().abs().sum()) - float(g["patches_checksum"])) < 1e-6
    np.testing.assert_allclose(out["desc"].numpy(), g["desc"], atol=2e-6)


def test_voxel_grid_and_quirks():
    c = so.voxel_centres()
    assert c.shape == (420, 3)
    r = c.norm(dim=1).view(3, 140)
    assert torch.allclose(r, torch.tensor([1 / 6, 3 / 6, 5 / 6])[:, None].expand(3, 140), atol=1e-6)
    # a patch whose every point sits on one voxel centre: slot 0 is zeroed by the mask1 quirk (first neighbour is patch slot 0)
    patch = c[37].repeat(512, 1)[None]
    g = so.sphere_query(patch)
    assert g[0, 37, 0].abs().sum() == 0 and torch.allclose(g[0, 37, 1], c[37])

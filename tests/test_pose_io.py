"""CPU: the `*_transform.txt` wire format against files written by the reference's own Evaluator._save_transformation_files
(tests/golden/eval.npz, oracle/make_golden.py eval)."""
import os
import re

import numpy as np
import pytest

from oracle.make_golden import eval_case

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval.npz"))
LINE = re.compile(r"^( *-?\d+\.\d{8})( +-?\d+\.\d{8}){3}$")


@pytest.mark.parametrize("tag,gen,with_global", [("plain", 0, False), ("global", "selected", True)])
def test_transform_files_match_reference(tmp_path, tag, gen, with_global):
    from rap_b200.rectified_point_flow.eval import load_transform, save_transformation_files
    d = eval_case()
    written = []
    for b in (0, 2):
        written += save_transformation_files(d, b, tmp_path, "synth", 7 + b, generation_idx=gen, rotations_pred=d["rotations_pred"],
                                             translations_pred=d["translations_pred"],
                                             global_rotation=d["global_rotation"] if with_global else None,
                                             global_translation=d["global_translation"] if with_global else None)
    assert sorted(p.name for p in written) == list(G[f"files_{tag}"])
    for name, text in zip(G[f"files_{tag}"], G[f"text_{tag}"]):
        mine = (tmp_path / str(name)).read_text()
        lines = mine.splitlines()
        assert len(lines) == 4 and all(LINE.match(l) and len(l) == 4 * 12 + 3 for l in lines)
        if mine != str(text):                                   # same numpy/LAPACK build -> byte identical; otherwise last-digit only
            np.testing.assert_allclose(np.loadtxt(str(text).splitlines()), load_transform(tmp_path / str(name)), atol=2e-6, rtol=0)


def test_relative_to_reference_and_suffixes(tmp_path):
    from rap_b200.rectified_point_flow.eval import relative_to_reference, save_transformation_files
    from rap_b200.rectified_point_flow.eval.transforms_io import _suffix
    assert _suffix(3) == "generation03" and _suffix("selected") == "generation_selected" and _suffix("generation07") == "generation07"
    rng = np.random.default_rng(0)
    Ts = []
    for _ in range(3):
        q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        T = np.eye(4); T[:3, :3] = q * np.sign(np.linalg.det(q)); T[:3, 3] = rng.standard_normal(3)
        Ts.append(T)
    rel = relative_to_reference(Ts)
    np.testing.assert_allclose(rel[0], np.eye(4), atol=1e-12)
    np.testing.assert_allclose(Ts[0] @ rel[2], Ts[2], atol=1e-12)
    assert relative_to_reference([]) == [] and save_transformation_files({}, 0, tmp_path, "x", 0) == []


def test_pose_errors_match_reference():
    """compute_transform_errors / _direct (reference eval/metrics.py:165-384, run verbatim for the golden) -- batched torch, runs anywhere."""
    from oracle.make_golden import eval_gt_case
    from rap_b200.rectified_point_flow.eval import compute_transform_errors, compute_transform_errors_direct
    d = eval_case()
    g = eval_gt_case(d)
    re_, te_ = compute_transform_errors(d["cond"], g["gt"], g["R_gt2"], g["t_gt2"], g["R_pr2"], g["t_pr2"], g["ppp2"], g["anchor2"],
                                        scale=g["scales2"], cu_seqlens_batch=g["cu2"])
    np.testing.assert_allclose(re_.numpy(), G["rot_err"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(te_.numpy(), G["trans_err"], rtol=1e-5, atol=1e-5)
    re_d, te_d = compute_transform_errors_direct(g["R_gt2"], g["t_gt2"], g["R_pr2"], g["t_pr2"], g["ppp2"], scale=g["scales2"])
    np.testing.assert_allclose(re_d.numpy(), G["rot_err_direct"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(te_d.numpy(), G["trans_err_direct"], rtol=1e-5, atol=1e-5)
    with pytest.raises(NotImplementedError):
        compute_transform_errors(d["cond"], g["gt"], g["R_gt2"], g["t_gt2"], g["R_pr2"], g["t_pr2"], g["ppp2"], g["anchor2"], use_icp=True)

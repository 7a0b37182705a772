"""CPU: the oracle restatement (oracle/edgecape_oracle.py) against the golden vectors frozen
from the unmodified reference (oracle/gen_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import edgecape_oracle
from oracle.gen_golden import CASES, build_case
from edgecape_b200.synthetic import make_state_dict
from edgecape_b200.config import state_dict_shapes

FAST = ["c1_tiny", "tiny_k100_2shot_masked", "tiny_allmasked", "c2_vitb_256_k100"]
SLOW = ["c4_vits_224_5shot", "c5_vitl_384_k200_full", "c2_vitb_256_k100_b16", "c4_vitb_256_5shot"]


def _check(name, golden_dir, dtype, tol):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg, data, wseed = build_case(name)
    sd = make_state_dict(state_dict_shapes(cfg), wseed)
    with torch.no_grad():
        o = edgecape_oracle.detector_forward_test(sd, cfg, data, dtype)
    for k, w in g.items():
        a = o[k].detach().numpy()
        if k in ("feature_q", "encoder_image"):
            a = a[:1]
        if k == "argmax":
            assert np.array_equal(a, w), f"{name}:{k}"      # bit-exact index parity
            continue
        assert a.shape == w.shape, (k, a.shape, w.shape)
        err = np.abs(a.astype(np.float64) - w).max() / (np.abs(w).max() + 1e-12)
        assert err < tol, f"{name}:{k} rel err {err:.3e}"


@pytest.mark.parametrize("name", FAST)
def test_oracle_fp32_matches_reference_golden(name, golden_dir):
    _check(name, golden_dir, torch.float32, 1e-4)


@pytest.mark.slow
@pytest.mark.parametrize("name", SLOW)
def test_oracle_fp32_matches_reference_golden_large(name, golden_dir):
    _check(name, golden_dir, torch.float32, 1e-4)


def test_oracle_fp64_matches_reference_golden(golden_dir):
    _check("tiny_k100_2shot_masked", golden_dir, torch.float64, 1e-4)


def test_all_cases_have_goldens(golden_dir):
    for name in CASES:
        assert os.path.exists(os.path.join(golden_dir, name + ".npz")), name

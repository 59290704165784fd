"""Design check kept green for the next round: the Lanczos formulation of the eigenvalue-clamped Newton solve
(scripts/lanczos_clamped_solve.py) agrees with the oracle's eigh-based safe_invert (cmf_solvers.py:346-356)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from oracle import cmf_oracle as O  # noqa: E402
from lanczos_clamped_solve import cases, lanczos_clamped_solve  # noqa: E402


@pytest.mark.parametrize("k", [7, 32, 128])
def test_lanczos_clamped_solve_matches_safe_invert(k):
    rng = np.random.RandomState(k)
    for name, H in cases(k, rng):
        H = (H + H.T) / 2
        for _ in range(3):
            g = rng.randn(k)
            ref = O.safe_invert(H[None], 0.2)[0] @ g
            got = lanczos_clamped_solve(H, g, 0.2)
            assert np.linalg.norm(got - ref) <= 1e-11 * max(np.linalg.norm(ref), 1e-30), name


@pytest.mark.parametrize("KR", [64, 128])
@pytest.mark.parametrize("kind", ["gram", "indef", "blockdiag"])
def test_register_resident_tridiagonalisation_mapping(KR, kind):
    """Thread-level emulation of tri::tridiag_reg (the clamped solve's Householder steps with the matrix in registers,
    pycmf_b200/csrc/tridiag_solve.cuh): the kernel's index mapping reproduces the spectrum of H, P T P^T = H with the
    reflectors it stores in W, and y = P^T g -- including steps with nothing to annihilate (block-diagonal input)."""
    from tridiag_reg_emulation import check
    ev_err, rec_err, y_err, _ = check(KR, 3, kind)
    assert ev_err < 1e-13 and rec_err < 1e-13 and y_err < 1e-13

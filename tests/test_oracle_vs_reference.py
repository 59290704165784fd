"""Live check of the CPU oracle against the UNMODIFIED reference (smn-ailab/PyCMF imported from /root/reference).

Runs only where the reference is mounted (the build container); on the GPU box it skips -- the committed fixtures under
tests/golden/ (produced by the same driver code, oracle/make_golden.py) carry the pin there.  Beyond the golden cases this
also runs the reference's full `fit_iterative_update` loop with its early-stop test (cmf_solvers.py:165-195) on a shape the
fixtures do not contain (initialisation vs the reference: tests/test_host_logic.py)."""
import warnings

import numpy as np
import pytest

from oracle import cmf_oracle as O
from oracle.cases import CASES, EDGE, make_case, make_edge_case
from oracle.make_golden import run_reference
from oracle.ref_loader import load_reference

from helpers import rel_fro, run_oracle

pytestmark = pytest.mark.skipif(load_reference() is None, reason="/root/reference is not mounted here")


@pytest.mark.parametrize("name", ["mu_dense_reg", "mu_csr", "nt_lin_logit", "nt_logit_logit", "nt_csr_logit_lin",
                                  "nt_noreg_clamp", "nt_sg_csr_lin_logit"])
def test_oracle_tracks_the_live_reference_step_by_step(name):
    case = make_case(name)
    hist_ref, U_ref, V_ref, Z_ref = run_reference(case)
    hist, U, V, Z = run_oracle(case)
    assert np.allclose(hist, hist_ref, rtol=1e-9, atol=1e-11), np.abs(hist - hist_ref).max()
    for got, ref in ((U, U_ref), (V, V_ref), (Z, Z_ref)):
        assert rel_fro(got, ref) < 1e-9


@pytest.mark.parametrize("solver", ["mu", "newton"])
def test_fit_loop_with_early_stop_matches_reference(solver):
    """The whole `fit_iterative_update` loop (error every 10 iterations, stop when the relative decrease drops below tol)
    returns the same n_iter and factors as the reference's."""
    load_reference()
    from pycmf.cmf_solvers import MUSolver, NewtonSolver
    rng = np.random.RandomState(17)
    n, d, l, k = 40, 30, 5, 4
    X = np.abs(rng.randn(n, k) @ rng.randn(k, d)) + 0.05 * np.abs(rng.randn(n, d))
    Y = np.abs(rng.randn(d, l))
    U0, V0, Z0 = 0.5 * np.abs(rng.randn(n, k)), 0.5 * np.abs(rng.randn(d, k)), 0.5 * np.abs(rng.randn(l, k))
    U, V, Z = U0.copy(), V0.copy(), Z0.copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if solver == "mu":
            s = MUSolver(max_iter=200, tol=1e-3, random_state=0)
            kw = dict(solver="mu", max_iter=200, tol=1e-3)
        else:
            p = dict(alpha=0.5, l1_reg=0.0, l2_reg=0.1, hessian_pertubation=0.2, U_non_negative=True, V_non_negative=True,
                     Z_non_negative=True, x_link="linear", y_link="linear", sg_sample_ratio=1.0)
            s = NewtonSolver(max_iter=60, tol=1e-3, random_state=0, **p)
            kw = dict(solver="newton", max_iter=60, tol=1e-3, **p)
        _, _, _, n_iter_ref = s.fit_iterative_update(X, Y, U, V, Z)
    Uo, Vo, Zo = U0.copy(), V0.copy(), Z0.copy()
    n_iter = O.fit_iterative_update(X, Y, Uo, Vo, Zo, **kw)
    n_iter = n_iter[-1] if isinstance(n_iter, tuple) else n_iter
    assert n_iter == n_iter_ref and n_iter_ref < kw["max_iter"]
    for got, ref in ((Uo, U), (Vo, V), (Zo, Z)):
        assert rel_fro(got, ref) < 1e-9


@pytest.mark.parametrize("name", sorted(EDGE))
def test_oracle_matches_the_live_reference_on_degenerate_shapes(name):
    """Ragged / minimal inputs (one row, one label column, rank one, an all-zero CSR row, two columns, sample sets of one
    and of zero indices, single-factor updates): the oracle must follow the reference here too, NumPy's global RNG in
    lock-step for the sampled cases."""
    case = make_edge_case(name)
    hist_ref, U_ref, V_ref, Z_ref = run_reference(case)
    hist, U, V, Z = run_oracle(case)
    assert np.isfinite(hist_ref).all()
    assert np.allclose(hist, hist_ref, rtol=1e-9, atol=1e-11), np.abs(hist - hist_ref).max()
    for got, ref in ((U, U_ref), (V, V_ref), (Z, Z_ref)):
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)

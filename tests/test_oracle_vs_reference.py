"""Live check of the CPU oracle against the UNMODIFIED reference (smn-ailab/PyCMF imported from /root/reference).

Runs only where the reference is mounted (the build container); on the GPU box it skips -- the committed fixtures under
tests/golden/ (produced by the same driver code, oracle/make_golden.py) carry the pin there.  Beyond the golden cases this
also runs the reference's full `fit_iterative_update` loop with its early-stop test (cmf_solvers.py:165-195) on a shape the
fixtures do not contain (initialisation vs the reference: tests/test_host_logic.py)."""
import warnings

import numpy as np
import pytest

from oracle import cmf_oracle as O
from oracle.cases import CASES, make_case
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


def _edge_case(seed, n, d, l, k, sparse, **params):
    """A tiny problem with degenerate dimensions (the generator of oracle/cases.py needs n, d, l > 1)."""
    import scipy.sparse as sp
    rng = np.random.RandomState(seed)
    logit_x, logit_y = params.get("x_link") == "logit", params.get("y_link") == "logit"
    X = rng.rand(n, d) if logit_x else np.abs(rng.randn(n, d))
    if sparse:
        X = X * (rng.rand(n, d) < 0.5)
        X[n // 2, :] = 0.0                                   # an empty row (and, when n == 1, an all-zero matrix)
        X = sp.csr_matrix(X)
    Y = rng.rand(d, l) if logit_y else np.abs(rng.randn(d, l))
    signed = params.get("U_non_negative", True) is False
    f = (lambda a: a) if signed else np.abs
    U0, V0, Z0 = f(0.3 * rng.randn(n, k)), f(0.3 * rng.randn(d, k)), f(0.3 * rng.randn(l, k))
    return dict(name="edge", X=X, Y=Y, U0=U0, V0=V0, Z0=Z0, params=dict(params), iters=4, rng_seed=100 + seed)


_NT = dict(solver="newton", alpha=0.4, l1_reg=0.01, l2_reg=0.1, hessian_pertubation=0.2, U_non_negative=True,
           V_non_negative=True, Z_non_negative=True, sg_sample_ratio=1.0)
EDGE = {
    "mu_one_row": (1, 9, 3, 2, False, dict(solver="mu")),
    "mu_one_label_column": (7, 5, 1, 3, False, dict(solver="mu", l1_reg=0.05)),
    "mu_rank_one": (6, 5, 4, 1, False, dict(solver="mu", l2_reg=0.1)),
    "mu_csr_empty_matrix_row": (1, 6, 2, 2, True, dict(solver="mu")),
    "mu_only_U": (5, 4, 3, 2, False, dict(solver="mu", update_V=False, update_Z=False)),
    "nt_one_row": (1, 7, 3, 2, False, dict(_NT, y_link="logit")),
    "nt_one_label_column": (8, 6, 1, 3, False, dict(_NT, x_link="logit", y_link="logit")),
    "nt_rank_one_signed": (9, 5, 3, 1, False, dict(_NT, U_non_negative=False, V_non_negative=False,
                                                   Z_non_negative=False)),
    "nt_csr_two_columns": (6, 2, 3, 2, True, dict(_NT, x_link="logit")),
    "nt_only_Z": (6, 5, 3, 2, False, dict(_NT, y_link="logit", update_U=False, update_V=False)),
    "nt_sg_samples_of_one": (6, 5, 3, 2, False, dict(_NT, sg_sample_ratio=0.3)),       # int(5*.3) = 1, int(3*.3) = 0
    "nt_sg_csr_logit": (5, 7, 4, 2, True, dict(_NT, x_link="logit", y_link="logit", sg_sample_ratio=0.5)),
}


@pytest.mark.parametrize("name", sorted(EDGE))
def test_oracle_matches_the_live_reference_on_degenerate_shapes(name):
    """Ragged / minimal inputs (one row, one label column, rank one, an all-zero CSR row, two columns, sample sets of one
    and of zero indices, single-factor updates): the oracle must follow the reference here too, NumPy's global RNG in
    lock-step for the sampled cases."""
    n, d, l, k, sparse, params = EDGE[name]
    case = _edge_case(sorted(EDGE).index(name), n, d, l, k, sparse, **params)
    hist_ref, U_ref, V_ref, Z_ref = run_reference(case)
    hist, U, V, Z = run_oracle(case)
    assert np.isfinite(hist_ref).all()
    assert np.allclose(hist, hist_ref, rtol=1e-9, atol=1e-11), np.abs(hist - hist_ref).max()
    for got, ref in ((U, U_ref), (V, V_ref), (Z, Z_ref)):
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)

"""Cross-check against the reference's compiled Cython twin (`pycmf/cmf_newton_solver.pyx`).

The twin is the native module libpycmf_b200.so replaces (`_newton_update_left` .pyx:241-292, `_newton_update_V`
.pyx:296-362).  It is dead code at the reference's HEAD (`USE_CYTHON = False`, cmf_solvers.py:11) and differs from the
live Python solver in ONE deterministic way: Z is updated through `_newton_update_left` (Y passed transposed,
weight 1 - alpha), whose logit Hessian has no `l2 I` term (.pyx:289) -- the C ABI exposes that as
`l2_in_logit_hessian`.  Needs `baseline/_ref` (built by `__graft_entry__.build()` where /root/reference is mounted;
the directory travels to the GPU box).
"""
import numpy as np
import pytest

from oracle import cmf_oracle as O
from oracle.cases import make_case
from oracle.ref_loader import load_reference_cython

pyx = load_reference_cython()
pytestmark = pytest.mark.skipif(pyx is None, reason="compiled cmf_newton_solver not available (baseline/_ref)")

CASES = ["nt_lin_lin", "nt_lin_logit", "nt_logit_lin", "nt_logit_logit", "nt_signed", "nt_noreg_clamp"]


def _params(case):
    p = case["params"]
    return (p["alpha"], p["l1_reg"], p["l2_reg"], p.get("x_link", "linear"), p.get("y_link", "linear"),
            p["hessian_pertubation"])


def _pyx_step(case, U, V, Z):
    """One U, Z, V iteration with the compiled twin, called as cmf_solvers.py:293-311 would (USE_CYTHON branch)."""
    alpha, l1, l2, xl, yl, pert = _params(case)
    p = case["params"]
    X, YT = np.ascontiguousarray(case["X"]), np.ascontiguousarray(case["Y"].T)
    pyx._newton_update_left(U, V, X, alpha, l1, l2, xl, p["U_non_negative"], 1.0, pert)
    pyx._newton_update_left(Z, V, YT, 1 - alpha, l1, l2, yl, p["Z_non_negative"], 1.0, pert)
    pyx._newton_update_V(V, U, Z, X, YT, alpha, l1, l2, xl, yl, p["V_non_negative"], 1.0, pert)


def _oracle_step(case, U, V, Z, z_l2_in_logit):
    alpha, l1, l2, xl, yl, pert = _params(case)
    p = case["params"]
    O.newton_update_U(U, V, case["X"], alpha, l1, l2, xl, p["U_non_negative"], pert)
    O._rows_newton(Z, V, np.asarray(case["Y"]).T, 1 - alpha, l1, l2, yl, p["Z_non_negative"], pert,
                   l2_in_logit_hessian=z_l2_in_logit)
    O.newton_update_V(V, U, Z, case["X"], case["Y"], alpha, l1, l2, xl, yl, p["V_non_negative"], pert)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_compiled_twin(name):
    """Three iterations: the oracle with the twin's Z quirk switched on reproduces the compiled module to 1e-10."""
    case = make_case(name)
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    Ur, Vr, Zr = U.copy(), V.copy(), Z.copy()
    for _ in range(3):
        _pyx_step(case, Ur, Vr, Zr)
        _oracle_step(case, U, V, Z, z_l2_in_logit=False)
    for got, ref in ((U, Ur), (V, Vr), (Z, Zr)):
        assert np.linalg.norm(got - ref) <= 1e-10 * max(1.0, np.linalg.norm(ref))


def test_twin_differs_from_live_solver_only_in_the_z_logit_hessian():
    """With a logit y link and l2 > 0 the twin's Z differs from the live solver's; with a linear y link
    it does not."""
    for name, differs in (("nt_lin_logit", True), ("nt_lin_lin", False)):
        case = make_case(name)
        U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
        Ur, Vr, Zr = U.copy(), V.copy(), Z.copy()
        _pyx_step(case, Ur, Vr, Zr)
        _oracle_step(case, U, V, Z, z_l2_in_logit=True)
        gap = np.linalg.norm(Z - Zr) / np.linalg.norm(Zr)
        assert (gap > 1e-3) if differs else (gap < 1e-10), (name, gap)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cabi_matches_compiled_twin(name):
    """pycmf_newton_left / pycmf_newton_v_* (float64) against the compiled module they replace, one iteration, called
    the way the twin is called: Z through the 'left' entry point with `l2_in_logit_hessian = 0`."""
    from pycmf_b200.device import CudaBackend
    be = CudaBackend(dtype="float64")
    case = make_case(name)
    alpha, l1, l2, xl, yl, pert = _params(case)
    p = case["params"]
    Ur, Vr, Zr = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    _pyx_step(case, Ur, Vr, Zr)
    X, Y = be.ingest(case["X"]), be.ingest(case["Y"])
    U, V, Z = be.to_device(case["U0"]), be.to_device(case["V0"]), be.to_device(case["Z0"])
    be.newton_left(U, V, X, alpha, l1, l2, xl, p["U_non_negative"], pert, l2_in_logit_hessian=False)
    be.newton_left(Z, V, Y, 1 - alpha, l1, l2, yl, p["Z_non_negative"], pert, l2_in_logit_hessian=False, trans=True)
    d = V.shape[0]
    gx, Hx, per_row = be.newton_v_xpart(V, U, X, 0, d, xl, alpha)
    be.newton_v_finish(V, Z, Y, 0, d, yl, alpha, l1, l2, gx, Hx, per_row, p["V_non_negative"], pert)
    for got, ref in ((U, Ur), (V, Vr), (Z, Zr)):
        got = be.to_host(got)
        assert np.linalg.norm(got - ref) <= 1e-9 * max(1.0, np.linalg.norm(ref))

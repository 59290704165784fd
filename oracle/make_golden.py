"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
For every case in oracle/cases.py it builds the reference's own MUSolver / NewtonSolver
(cmf_solvers.py:198, :319), seeds the global RNG the way the solver ctor does (:121-122), then
drives ``update_step`` (:248, :510) + ``compute_error`` (:128) once per iteration and stores the
per-iteration objective, the final factors and the inputs.
"""
import os
import warnings

import numpy as np
import scipy.sparse as sp

from .cases import CASES, make_case
from .ref_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_reference(case):
    ref = load_reference()
    if ref is None:
        raise RuntimeError("reference not available")
    from pycmf.cmf_solvers import MUSolver, NewtonSolver
    p = dict(case["params"])
    solver = p.pop("solver")
    X, Y = case["X"], case["Y"]
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    if solver == "mu":
        s = MUSolver(max_iter=case["iters"], tol=0, l1_reg=p.get("l1_reg", 0.), l2_reg=p.get("l2_reg", 0.),
                     update_U=p.get("update_U", True), update_V=p.get("update_V", True),
                     update_Z=p.get("update_Z", True), random_state=case["rng_seed"])
    else:
        s = NewtonSolver(max_iter=case["iters"], tol=0, random_state=case["rng_seed"], **p)
    hist = [s.compute_error(X, Y, U, V, Z)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(case["iters"]):
            s.update_step(X, Y, U, V, Z, s.l1_reg, s.l2_reg, s.alpha)
            hist.append(s.compute_error(X, Y, U, V, Z))
    return np.asarray(hist, dtype=np.float64), U, V, Z


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        case = make_case(name)
        hist, U, V, Z = run_reference(case)
        X = case["X"]
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            X=X.toarray() if sp.issparse(X) else X, X_is_sparse=sp.issparse(X), Y=case["Y"],
            U0=case["U0"], V0=case["V0"], Z0=case["Z0"], objective=hist, U=U, V=V, Z=Z)
        print("%-22s obj %.6f -> %.6f  finite=%s" % (name, hist[0], hist[-1], np.isfinite(hist).all()))


if __name__ == "__main__":
    main()

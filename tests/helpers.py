"""Shared test helpers: golden loading and the oracle runner."""
import os

import numpy as np
import scipy.sparse as sp

from oracle import cmf_oracle as O
from oracle.cases import CASES, make_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    case = make_case(name)
    X = sp.csr_matrix(g["X"]) if bool(g["X_is_sparse"]) else g["X"]
    case.update(X=X, Y=g["Y"], U0=g["U0"], V0=g["V0"], Z0=g["Z0"])
    return case, g


def solver_kwargs(case):
    p = dict(case["params"])
    return p


def run_oracle(case, masks_per_iter=None):
    """Per-iteration objective + final factors from the CPU oracle (float64)."""
    p = dict(case["params"])
    U, V, Z = case["U0"].copy(), case["V0"].copy(), case["Z0"].copy()
    solver = p["solver"]
    if solver == "mu":
        e = (0.5, "linear", "linear")
    else:
        e = (p["alpha"], p.get("x_link", "linear"), p.get("y_link", "linear"))
    hist = [O.compute_error(case["X"], case["Y"], U, V, Z, *e)]
    h = []
    O.fit_iterative_update(case["X"], case["Y"], U, V, Z, max_iter=case["iters"], tol=0,
                           random_state=case["rng_seed"], history=h,
                           masks_per_iter=masks_per_iter, **p)
    return np.asarray(hist + h), U, V, Z


def draw_masks_for_case(case):
    """Replay the reference's RNG stream (seeded like the solver ctor) to get the per-iteration masks."""
    p = case["params"]
    ratio = p.get("sg_sample_ratio", 1.0)
    if p["solver"] != "newton" or ratio >= 1.0:
        return None
    n, d = case["X"].shape
    l = case["Y"].shape[1]
    np.random.seed(case["rng_seed"])
    return [O.draw_newton_masks(n, d, l, ratio, p.get("update_U", True), p.get("update_Z", True),
                                p.get("update_V", True)) for _ in range(case["iters"])]


def rel_fro(a, b):
    return np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300)

"""Seeded small parity cases shared by oracle/make_golden.py and tests/ -- TEST INFRASTRUCTURE.

Every case is a dict: X (ndarray or csr), Y, U0, V0, Z0 (float64) and ``params`` (solver kwargs in
the reference's vocabulary) plus ``iters``.  Inputs are regenerated from the seed; the golden
files additionally store them so a NumPy RNG change cannot silently move the fixtures.
"""
import numpy as np
import scipy.sparse as sp
from scipy.special import expit


def _init(rng, X, Y, k, non_negative=True):
    """Same scaling as the reference's 'random' init (cmf.py:110-117) and the (V+V_)/2 merge (:425-426)."""
    n, d = X.shape
    l = Y.shape[1]
    ax = np.sqrt(np.abs(X.mean()) / k)
    ay = np.sqrt(np.abs(Y.mean()) / k)
    U = ax * rng.randn(n, k)
    V = (ax * rng.randn(d, k) + ay * rng.randn(d, k)) / 2
    Z = ay * rng.randn(l, k)
    if non_negative:
        U, V, Z = np.abs(U), np.abs(V), np.abs(Z)
    return U, V, Z


def _planted(rng, n, d, l, k, x_kind, y_kind, density=None):
    Ut, Vt, Zt = 0.7 * np.abs(rng.randn(n, k)), 0.7 * np.abs(rng.randn(d, k)), 0.7 * rng.randn(l, k)
    if x_kind == "nonneg":
        X = Ut @ Vt.T + 0.05 * np.abs(rng.randn(n, d))
    elif x_kind == "signed":
        X = rng.randn(n, d)
    elif x_kind == "unit":          # targets in (0, 1) for the logit link
        X = expit(Ut @ Vt.T - 1.0 + 0.1 * rng.randn(n, d))
    if density is not None:
        X = X * (rng.rand(n, d) < density)
        X[0, :] = 0.0              # an empty row
        X[:, 1] = 0.0              # an empty column
        X = sp.csr_matrix(X)
    if y_kind == "nonneg":
        Y = np.abs(Vt @ Zt.T) + 0.05 * np.abs(rng.randn(d, l))
    elif y_kind == "unit":
        Y = expit(Vt @ Zt.T)
    elif y_kind == "binary":
        Y = (rng.rand(d, l) < expit(Vt @ Zt.T)).astype(float)
    return X, Y


_NEWTON = dict(solver="newton", alpha=0.3, l1_reg=0.02, l2_reg=0.1, hessian_pertubation=0.2,
               U_non_negative=True, V_non_negative=True, Z_non_negative=False, sg_sample_ratio=1.0)

# name -> (seed, n, d, l, k, x_kind, y_kind, density, iters, params)
CASES = {
    "mu_dense":        (1, 30, 20, 6, 5, "nonneg", "nonneg", None, 25, dict(solver="mu")),
    "mu_dense_reg":    (2, 30, 20, 6, 5, "nonneg", "nonneg", None, 25, dict(solver="mu", l1_reg=0.1, l2_reg=0.05)),
    "mu_csr":          (3, 40, 25, 4, 6, "nonneg", "nonneg", 0.3, 25, dict(solver="mu")),
    "mu_csr_reg":      (4, 40, 25, 4, 6, "nonneg", "binary", 0.2, 25, dict(solver="mu", l1_reg=0.05, l2_reg=0.02)),
    "mu_k_gt_d":       (5, 12, 7, 3, 9, "nonneg", "nonneg", None, 10, dict(solver="mu")),
    "mu_no_V":         (6, 20, 15, 4, 5, "nonneg", "nonneg", None, 10, dict(solver="mu", update_V=False)),
    "nt_lin_lin":      (11, 28, 18, 5, 4, "nonneg", "nonneg", None, 12, dict(_NEWTON)),
    "nt_lin_logit":    (12, 28, 18, 5, 4, "nonneg", "unit", None, 12, dict(_NEWTON, y_link="logit")),
    "nt_logit_lin":    (13, 28, 18, 5, 4, "unit", "nonneg", None, 12, dict(_NEWTON, x_link="logit")),
    "nt_logit_logit":  (14, 28, 18, 5, 4, "unit", "binary", None, 12, dict(_NEWTON, x_link="logit", y_link="logit")),
    "nt_signed":       (15, 28, 18, 5, 4, "signed", "nonneg", None, 8,
                        dict(_NEWTON, U_non_negative=False, V_non_negative=False, Z_non_negative=False, l1_reg=0.0)),
    "nt_noreg_clamp":  (16, 28, 18, 5, 6, "unit", "unit", None, 6,
                        dict(_NEWTON, x_link="logit", y_link="logit", l1_reg=0.0, l2_reg=0.0, alpha=0.5)),
    "nt_csr_lin_logit": (17, 32, 20, 4, 4, "nonneg", "binary", 0.3, 10, dict(_NEWTON, y_link="logit")),
    "nt_csr_logit_lin": (18, 32, 20, 4, 4, "unit", "nonneg", 0.3, 10, dict(_NEWTON, x_link="logit")),
    "nt_no_V":         (19, 28, 18, 5, 4, "nonneg", "unit", None, 8, dict(_NEWTON, y_link="logit", update_V=False)),
    "nt_sg_lin_lin":   (21, 24, 16, 6, 4, "nonneg", "nonneg", None, 6, dict(_NEWTON, sg_sample_ratio=0.5)),
    "nt_sg_logit_logit": (22, 24, 16, 6, 4, "unit", "binary", None, 6,
                          dict(_NEWTON, x_link="logit", y_link="logit", sg_sample_ratio=0.5)),
    "nt_sg_csr_lin_logit": (23, 24, 16, 6, 4, "nonneg", "binary", 0.35, 6,
                            dict(_NEWTON, y_link="logit", sg_sample_ratio=0.5)),
    "nt_sg_zero_ysample": (24, 48, 40, 6, 3, "unit", "unit", None, 4,
                           dict(_NEWTON, x_link="logit", y_link="logit", sg_sample_ratio=0.15)),  # int(6*0.15) == 0
}


def make_case(name):
    seed, n, d, l, k, x_kind, y_kind, density, iters, params = CASES[name]
    rng = np.random.RandomState(seed)
    X, Y = _planted(rng, n, d, l, k, x_kind, y_kind, density)
    nonneg = params.get("U_non_negative", True)
    U0, V0, Z0 = _init(rng, X, Y, k, non_negative=nonneg)
    if params.get("Z_non_negative", True) is False and nonneg:
        Z0 = Z0 * np.sign(rng.randn(*Z0.shape))
    return dict(name=name, X=X, Y=Y, U0=U0, V0=V0, Z0=Z0, params=dict(params), iters=iters,
                rng_seed=1000 + seed)


# ---- degenerate shapes: live oracle-vs-reference test, host orchestration test, (isolated) GPU test
def _edge_case(seed, n, d, l, k, sparse, **params):
    """A tiny problem with degenerate dimensions (`_planted` above needs n, d, l > 1)."""
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


_NT_EDGE = dict(solver="newton", alpha=0.4, l1_reg=0.01, l2_reg=0.1, hessian_pertubation=0.2, U_non_negative=True,
                V_non_negative=True, Z_non_negative=True, sg_sample_ratio=1.0)
EDGE = {
    "mu_one_row": (1, 9, 3, 2, False, dict(solver="mu")),
    "mu_one_label_column": (7, 5, 1, 3, False, dict(solver="mu", l1_reg=0.05)),
    "mu_rank_one": (6, 5, 4, 1, False, dict(solver="mu", l2_reg=0.1)),
    "mu_csr_empty_matrix_row": (1, 6, 2, 2, True, dict(solver="mu")),
    "mu_only_U": (5, 4, 3, 2, False, dict(solver="mu", update_V=False, update_Z=False)),
    "nt_one_row": (1, 7, 3, 2, False, dict(_NT_EDGE, y_link="logit")),
    "nt_one_label_column": (8, 6, 1, 3, False, dict(_NT_EDGE, x_link="logit", y_link="logit")),
    "nt_rank_one_signed": (9, 5, 3, 1, False, dict(_NT_EDGE, U_non_negative=False, V_non_negative=False,
                                                   Z_non_negative=False)),
    "nt_csr_two_columns": (6, 2, 3, 2, True, dict(_NT_EDGE, x_link="logit")),
    "nt_only_Z": (6, 5, 3, 2, False, dict(_NT_EDGE, y_link="logit", update_U=False, update_V=False)),
    "nt_sg_samples_of_one": (6, 5, 3, 2, False, dict(_NT_EDGE, sg_sample_ratio=0.3)),       # int(5*.3) = 1, int(3*.3) = 0
    "nt_sg_csr_logit": (5, 7, 4, 2, True, dict(_NT_EDGE, x_link="logit", y_link="logit", sg_sample_ratio=0.5)),
}


def make_edge_case(name):
    n, d, l, k, sparse, params = EDGE[name]
    return _edge_case(sorted(EDGE).index(name), n, d, l, k, sparse, **params)

"""CPU oracle for PyCMF's fit loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 NumPy/SciPy restatement of the reference's two solvers and objective.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; ``pycmf_b200`` never does.

Parity status: PINNED.  The reference (smn-ailab/PyCMF) ships no golden vectors
(its tests are property tests only), so this oracle is pinned against outputs of
the *unmodified reference itself*, generated in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (see
``tests/test_oracle_golden.py``), and -- when ``/root/reference`` is present --
live against the reference in ``tests/test_oracle_vs_reference.py``.

Reference lines followed (all into /root/reference/pycmf/cmf_solvers.py):
  * objective ............. :18-42, :128-130  (+ sklearn ``_beta_divergence`` beta=2)
  * fit loop .............. :132-195
  * MU step ............... :212-263   (order V, U, Z)
  * Newton step ........... :321-522   (order U, Z, V; live pure-Python class)
  * sampling RNG order .... :328-344 called from :414, :455-456, :494

Permitted, parity-neutral restatements (each verified against the
live reference in tests/test_oracle_vs_reference.py):
  * ``(U V^T) V -> U (V^T V)``            (MU denominators, :233, :239)
  * ``A^T diag(w) A -> (A * w)^T A``       (no s x s ``np.diag`` temporaries)
  * per-row sampling by index instead of ``X[:, mask]`` copies; masks can be
    injected (``masks=``) or drawn from NumPy's legacy global RNG in the
    reference's call order (``draw_newton_masks``)
  * rows of one factor are updated as a batch (rows are independent, :412, :452, :491)
"""
import time

import numpy as np
import scipy.sparse as sp
from scipy.special import expit

EPSILON = np.finfo(np.float32).eps  # cmf_solvers.py:12


# --------------------------------------------------------------------------- links
def inverse(x, link):
    """cmf_solvers.py:27-33."""
    if link == "linear":
        return x
    elif link == "logit":
        return expit(x)
    raise ValueError("Invalid link function {}".format(link))


def d_sigmoid(x):
    """cmf_solvers.py:22-24."""
    s = expit(x)
    return s * (1 - s)


def _dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M)


# ----------------------------------------------------------------------- objective
def compute_factorization_error(target, left, right_T, link):
    """cmf_solvers.py:36-42 with beta_loss == 2 ('frobenius').

    ``right_T`` is the k x cols right factor (the reference passes ``V.T``).
    linear + dense : ||T - L R||_F            (sklearn _beta_divergence, square_root=True)
    linear + sparse: sqrt(||T||^2 + tr((L^T L)(R R^T)) - 2 tr((T R^T)^T L))
    logit          : ||T - sigmoid(L R)||_F   (dense, also for sparse T)
    """
    if target is None:
        return 0
    if link == "linear":
        if sp.issparse(target):
            norm_t = np.dot(target.data, target.data)
            norm_lr = np.sum((left.T @ left) * (right_T @ right_T.T))
            cross = np.sum(np.asarray(target @ right_T.T) * left)
            return np.sqrt(max(norm_t + norm_lr - 2.0 * cross, 0.0))
        return np.linalg.norm(np.asarray(target) - left @ right_T)
    elif link == "logit":
        return np.linalg.norm(_dense(target) - expit(left @ right_T))
    raise ValueError("Invalid link function {}".format(link))


def compute_error(X, Y, U, V, Z, alpha=0.5, x_link="linear", y_link="linear"):
    """cmf_solvers.py:128-130."""
    return alpha * compute_factorization_error(X, U, V.T, x_link) + \
        (1 - alpha) * compute_factorization_error(Y, V, Z.T, y_link)


# ------------------------------------------------------------------------------ MU
def _regularized_delta(num, den, l1_reg, l2_reg, H):
    """cmf_solvers.py:212-228 with gamma == 1."""
    if l1_reg > 0:
        den = den + l1_reg
    if l2_reg > 0:
        den = den + l2_reg * H
    den = np.where(den == 0, EPSILON, den)
    return num / den


def mu_step(X, Y, U, V, Z, l1_reg=0., l2_reg=0.,
            update_U=True, update_V=True, update_Z=True):
    """One MU iteration, in place, order V -> U -> Z (cmf_solvers.py:248-263)."""
    if update_V:
        num = np.asarray(X.T @ U) + np.asarray(Y @ Z)            # :244
        den = V @ (U.T @ U + Z.T @ Z)                           # :245
        V *= _regularized_delta(num, den, l1_reg, l2_reg, V)
    if update_U:
        num = np.asarray(X @ V)                                 # :232
        den = U @ (V.T @ V)                                     # :233 re-associated
        U *= _regularized_delta(num, den, l1_reg, l2_reg, U)
    if update_Z:
        num = np.asarray(Y.T @ V)                               # :238
        den = Z @ (V.T @ V)                                     # :239 re-associated
        Z *= _regularized_delta(num, den, l1_reg, l2_reg, Z)


# -------------------------------------------------------------------------- Newton
def safe_invert(M, pert):
    """cmf_solvers.py:346-356; accepts a batch (..., k, k)."""
    w, Q = np.linalg.eigh(M)
    w = np.abs(w)
    w = np.where(w < pert, pert, w)
    return (Q / w[..., None, :]) @ np.swapaxes(Q, -1, -2)


def draw_newton_masks(n, d, l, ratio, update_U=True, update_Z=True, update_V=True):
    """Draw one update_step's sample index sets from NumPy's *global legacy RNG*
    in the reference's call order: n x perm(d) [U, :414] ->
    l x perm(d) [Z, :494] -> d x (perm(n), perm(l)) [V, :455-456].
    Returns dict of int32 arrays: U (n, s_d), Z (l, s_d), Vx (d, s_n), Vy (d, s_l).
    """
    if ratio >= 1.:
        return None
    s_d, s_n, s_l = int(d * ratio), int(n * ratio), int(l * ratio)
    out = {}
    if update_U:
        out["U"] = np.stack([np.random.permutation(np.arange(d))[:s_d] for _ in range(n)]
                            ).astype(np.int32).reshape(n, s_d)
    if update_Z:
        out["Z"] = np.stack([np.random.permutation(np.arange(d))[:s_d] for _ in range(l)]
                            ).astype(np.int32).reshape(l, s_d)
    if update_V:
        vx, vy = [], []
        for _ in range(d):
            vx.append(np.random.permutation(np.arange(n))[:s_n])
            vy.append(np.random.permutation(np.arange(l))[:s_l])
        out["Vx"] = np.stack(vx).astype(np.int32).reshape(d, s_n)
        out["Vy"] = np.stack(vy).astype(np.int32).reshape(d, s_l)
    return out


def _rows_newton(F, B, T, weight, l1_reg, l2_reg, link, non_negative, pert,
                 l2_in_logit_hessian, idx=None, chunk=256):
    """Batched restatement of the per-row Newton update of a 'left' factor.

    F (rows x k) is updated in place against B (m x k) and target T (rows x m):
        g_i = weight * (f(f_i B_s^T) - T[i, s]) B_s + l1 sign(f_i) + l2 f_i
        H_i = weight * B_s^T [diag(f'(f_i B_s^T))] B_s (+ l2 I)
        f_i <- f_i - g_i S(H_i);  clamp at 0 if non_negative
    ``idx`` (rows x s) selects B rows / T columns per row (None = all).
    U: cmf_solvers.py:394-430 (l2 in Hessian only for linear);
    Z: :488-508 with T = Y^T (l2 in Hessian for both links).
    """
    rows, k = F.shape
    eye = np.eye(k)
    T_sparse = sp.issparse(T)
    if T_sparse:
        T = T.tocsr()
    F0 = F.copy()
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        f = F0[r0:r1]
        t = T[r0:r1].toarray() if T_sparse else np.asarray(T[r0:r1])
        if idx is None:
            est = f @ B.T                                          # (c, m)
            res = inverse(est, link) - t
            g = weight * (res @ B)
            if link == "linear":
                H = weight * (B.T @ B) + l2_reg * eye
                Hinv = safe_invert(H, pert)
                step = g + l1_reg * np.sign(f) + l2_reg * f
                new = f - step @ Hinv
            else:
                w = d_sigmoid(est)                                 # (c, m)
                H = weight * np.einsum('cj,ja,jb->cab', w, B, B)
                if l2_in_logit_hessian:
                    H = H + l2_reg * eye
                Hinv = safe_invert(H, pert)
                step = g + l1_reg * np.sign(f) + l2_reg * f
                new = f - np.einsum('ca,cab->cb', step, Hinv)
        else:
            ix = idx[r0:r1]                                        # (c, s)
            Bs = B[ix]                                             # (c, s, k)
            est = np.einsum('ck,csk->cs', f, Bs)
            ts = np.take_along_axis(t, ix, axis=1)
            res = inverse(est, link) - ts
            g = weight * np.einsum('cs,csk->ck', res, Bs)
            if link == "linear":
                H = weight * np.einsum('csa,csb->cab', Bs, Bs) + l2_reg * eye
            else:
                w = d_sigmoid(est)
                H = weight * np.einsum('cs,csa,csb->cab', w, Bs, Bs)
                if l2_in_logit_hessian:
                    H = H + l2_reg * eye
            Hinv = safe_invert(H, pert)
            step = g + l1_reg * np.sign(f) + l2_reg * f
            new = f - np.einsum('ca,cab->cb', step, Hinv)
        if non_negative:
            new = np.where(new < 0, 0., new)
        F[r0:r1] = new


def newton_update_U(U, V, X, alpha, l1_reg, l2_reg, link, non_negative, pert, idx=None):
    """cmf_solvers.py:394-430. Logit Hessian has NO l2 term (:428)."""
    _rows_newton(U, V, X, alpha, l1_reg, l2_reg, link, non_negative, pert,
                 l2_in_logit_hessian=False, idx=idx)


def newton_update_Z(Z, V, Y, alpha, l1_reg, l2_reg, link, non_negative, pert, idx=None):
    """cmf_solvers.py:488-508. Weight is (1 - alpha); l2 in Hessian for both links (:501-506)."""
    YT = Y.T.tocsr() if sp.issparse(Y) else np.asarray(Y).T
    _rows_newton(Z, V, YT, 1 - alpha, l1_reg, l2_reg, link, non_negative, pert,
                 l2_in_logit_hessian=True, idx=idx)


def newton_update_V(V, U, Z, X, Y, alpha, l1_reg, l2_reg, x_link, y_link,
                    non_negative, pert, idx_x=None, idx_y=None, chunk=128):
    """cmf_solvers.py:432-486. Uses the already-updated U and Z.
        g_j = alpha (f1(U_s v_j) - X[s, j])^T U_s + (1-alpha) (f2(v_j Z_t^T) - Y[j, t]) Z_t
              + l1 sign(v_j) + l2 v_j
        H_j = alpha U_s^T [D_u] U_s + (1-alpha) Z_t^T [D_z] Z_t + l2 I
    idx_x (d x s_n) indexes rows of U / X; idx_y (d x s_l) indexes rows of Z / columns of Y.
    """
    d, k = V.shape
    eye = np.eye(k)
    X_sparse, Y_sparse = sp.issparse(X), sp.issparse(Y)
    XT = X.T.tocsr() if X_sparse else None
    Yr = Y.tocsr() if Y_sparse else np.asarray(Y)
    V0 = V.copy()
    sampled = idx_x is not None
    for r0 in range(0, d, chunk):
        r1 = min(d, r0 + chunk)
        v = V0[r0:r1]                                              # (c, k)
        xt = XT[r0:r1].toarray() if X_sparse else np.asarray(X)[:, r0:r1].T   # (c, n)
        y = Yr[r0:r1].toarray() if Y_sparse else Yr[r0:r1]                    # (c, l)
        if not sampled:
            est_x = v @ U.T                                        # (c, n)
            est_y = v @ Z.T                                        # (c, l)
            g = alpha * ((inverse(est_x, x_link) - xt) @ U) + \
                (1 - alpha) * ((inverse(est_y, y_link) - y) @ Z)
            if x_link == "logit":
                Hx = np.einsum('ci,ia,ib->cab', d_sigmoid(est_x), U, U)
            else:
                Hx = (U.T @ U)[None]
            if y_link == "logit":
                Hy = np.einsum('ci,ia,ib->cab', d_sigmoid(est_y), Z, Z)
            else:
                Hy = (Z.T @ Z)[None]
        else:
            ix, iy = idx_x[r0:r1], idx_y[r0:r1]
            Us, Zs = U[ix], Z[iy]                                  # (c, s_n, k), (c, s_l, k)
            est_x = np.einsum('ck,csk->cs', v, Us)
            est_y = np.einsum('ck,csk->cs', v, Zs)
            rx = inverse(est_x, x_link) - np.take_along_axis(xt, ix, axis=1)
            ry = inverse(est_y, y_link) - np.take_along_axis(y, iy, axis=1)
            g = alpha * np.einsum('cs,csk->ck', rx, Us) + (1 - alpha) * np.einsum('cs,csk->ck', ry, Zs)
            wx = d_sigmoid(est_x) if x_link == "logit" else np.ones_like(est_x)
            wy = d_sigmoid(est_y) if y_link == "logit" else np.ones_like(est_y)
            Hx = np.einsum('cs,csa,csb->cab', wx, Us, Us)
            Hy = np.einsum('cs,csa,csb->cab', wy, Zs, Zs)
        H = alpha * Hx + (1 - alpha) * Hy + l2_reg * eye
        Hinv = safe_invert(H, pert)
        step = g + l1_reg * np.sign(v) + l2_reg * v
        if Hinv.shape[0] == 1:
            new = v - step @ Hinv[0]
        else:
            new = v - np.einsum('ca,cab->cb', step, Hinv)
        if non_negative:
            new = np.where(new < 0, 0., new)
        V[r0:r1] = new


def newton_step(X, Y, U, V, Z, alpha=0.5, l1_reg=0., l2_reg=0.,
                x_link="linear", y_link="linear",
                U_non_negative=True, V_non_negative=True, Z_non_negative=True,
                hessian_pertubation=0.2, sg_sample_ratio=1.,
                update_U=True, update_V=True, update_Z=True, masks=None):
    """One Newton iteration in place, order U -> Z -> V (cmf_solvers.py:510-522).

    With ``sg_sample_ratio < 1`` the per-row sample index sets are taken from
    ``masks`` (dict as returned by ``draw_newton_masks``) or, if None, drawn from
    NumPy's global legacy RNG in the reference's order.
    """
    n, d = X.shape
    l = Y.shape[1]
    if sg_sample_ratio < 1. and masks is None:
        masks = draw_newton_masks(n, d, l, sg_sample_ratio, update_U, update_Z, update_V)
    m = masks or {}
    if update_U:
        newton_update_U(U, V, X, alpha, l1_reg, l2_reg, x_link, U_non_negative,
                        hessian_pertubation, idx=m.get("U"))
    if update_Z:
        newton_update_Z(Z, V, Y, alpha, l1_reg, l2_reg, y_link, Z_non_negative,
                        hessian_pertubation, idx=m.get("Z"))
    if update_V:
        newton_update_V(V, U, Z, X, Y, alpha, l1_reg, l2_reg, x_link, y_link,
                        V_non_negative, hessian_pertubation,
                        idx_x=m.get("Vx"), idx_y=m.get("Vy"))


# ------------------------------------------------------------------------ fit loop
def fit_iterative_update(X, Y, U, V, Z, solver="mu", max_iter=200, tol=1e-4,
                         alpha=0.5, l1_reg=0., l2_reg=0., x_link="linear", y_link="linear",
                         U_non_negative=True, V_non_negative=True, Z_non_negative=True,
                         hessian_pertubation=0.2, sg_sample_ratio=1., random_state=None,
                         update_U=True, update_V=True, update_Z=True,
                         verbose=0, history=None, masks_per_iter=None):
    """cmf_solvers.py:132-195. Factors are float64 arrays updated in place.

    MU ignores alpha/links for the update and reports the error with alpha=0.5
    and linear links (cmf.py:434-439, cmf_solvers.py:99).
    ``history`` (list) receives the objective after EVERY iteration (parity harness;
    the reference itself evaluates it only every 10th iteration).
    """
    if random_state is not None:
        np.random.seed(random_state)                               # :121-122
    if solver == "mu":
        e_alpha, e_xl, e_yl = 0.5, "linear", "linear"
    else:
        e_alpha, e_xl, e_yl = alpha, x_link, y_link
    start = time.time()
    previous_error = error_at_init = compute_error(X, Y, U, V, Z, e_alpha, e_xl, e_yl)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        if solver == "mu":
            mu_step(X, Y, U, V, Z, l1_reg, l2_reg, update_U, update_V, update_Z)
        elif solver == "newton":
            newton_step(X, Y, U, V, Z, alpha, l1_reg, l2_reg, x_link, y_link,
                        U_non_negative, V_non_negative, Z_non_negative,
                        hessian_pertubation, sg_sample_ratio,
                        update_U, update_V, update_Z,
                        masks=None if masks_per_iter is None else masks_per_iter[n_iter - 1])
        else:
            raise ValueError("No such solver: %s" % solver)
        if history is not None:
            history.append(compute_error(X, Y, U, V, Z, e_alpha, e_xl, e_yl))
        if tol > 0 and n_iter % 10 == 0:
            error = compute_error(X, Y, U, V, Z, e_alpha, e_xl, e_yl)
            if verbose:
                print("Epoch %02d reached after %.3f seconds, error: %f" %
                      (n_iter, time.time() - start, error))
            if (previous_error - error) / error_at_init < tol:
                break
            previous_error = error
    if verbose and (tol == 0 or n_iter % 10 != 0):
        print("Epoch %02d reached after %.3f seconds." % (n_iter, time.time() - start))
    return U, V, Z, n_iter

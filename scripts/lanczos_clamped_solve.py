"""Design check for the next round (CPU, NumPy): the eigenvalue-clamped solve x = S(H) g of the Newton step
(reference _safe_invert, cmf_solvers.py:346-356: S(H) = Q diag(1 / max(|lambda|, p)) Q^T) WITHOUT a full eigendecomposition.

x = f(H) g with f(t) = 1 / max(|t|, p).  k steps of Lanczos on (H, g) with full reorthogonalisation give H Q = Q T exactly
(T tridiagonal, Q^T g = |g| e_1; the recurrence stops early when the Krylov space of g is exhausted, e.g. rank-deficient H),
so  x = |g| Q f(T) e_1 : one k x k tridiagonal eigenproblem (O(k^2) for the values; eigenvectors of T by QL or inverse
iteration) and k matrix-vector products -- ~2 k^3 flop in k sequential, CTA-parallel steps against ~10 Jacobi sweeps of
~3 k^3 flop each with k/2 synchronisations per sweep (the C4 slice spends 1.6 s per iteration in those sweeps at k = 128).

This script measures the accuracy of that formulation against the dense eigh formula on the kinds of Hessians the fit
produces (well conditioned, rank deficient, saturated-logit = nearly zero with a few large eigenvalues, indefinite)."""
import numpy as np
from scipy.linalg import eigh_tridiagonal


def safe_invert_apply(H, g, p):
    w, Q = np.linalg.eigh(H)
    return Q @ ((Q.T @ g) / np.maximum(np.abs(w), p))


def lanczos_clamped_solve(H, g, p, reorth_passes=2):
    k = H.shape[0]
    nrm = np.linalg.norm(g)
    if nrm == 0.0:
        return np.zeros_like(g)
    Q = np.zeros((k, k))
    alpha, beta = np.zeros(k), np.zeros(k)
    q = g / nrm
    scale = max(np.abs(H).sum(1).max(), p)        # ||H||_inf bound: breakdown threshold
    m = 0
    for j in range(k):
        Q[:, j] = q
        m = j + 1
        w = H @ q
        alpha[j] = q @ w
        w -= alpha[j] * q
        if j > 0:
            w -= beta[j - 1] * Q[:, j - 1]
        for _ in range(reorth_passes):               # full reorthogonalisation (classical Gram-Schmidt, twice)
            w -= Q[:, :m] @ (Q[:, :m].T @ w)
        b = np.linalg.norm(w)
        if b <= 1e-14 * scale or j == k - 1:         # Krylov space exhausted: T_m is exact on span(Q)
            break
        beta[j] = b
        q = w / b
    theta, S = eigh_tridiagonal(alpha[:m], beta[:m - 1])
    y = S @ (S[0] / np.maximum(np.abs(theta), p))    # f(T) e_1
    return nrm * (Q[:, :m] @ y)


def cases(k, rng):
    A = rng.randn(k, 3 * k)
    yield "well conditioned Gram", A @ A.T / (3 * k) + 0.5 * np.eye(k)
    A = rng.randn(k, k // 3)
    yield "rank k/3 Gram (clamp on 2k/3 directions)", A @ A.T
    V = 0.3 * np.abs(rng.randn(400, k))
    w = 0.25 * np.exp(-np.abs(3.0 * rng.randn(400)) * 3)          # saturated sigmoids: sigma' mostly tiny
    yield "saturated logit Hessian 0.5 V^T diag(s') V", 0.5 * (V * w[:, None]).T @ V
    yield "same + 0.1 I", 0.5 * (V * w[:, None]).T @ V + 0.1 * np.eye(k)
    B = rng.randn(k, k)
    yield "indefinite symmetric (abs of negative eigenvalues)", (B + B.T) / 4
    yield "zero matrix", np.zeros((k, k))
    d = np.concatenate([np.full(k // 2, 0.2), np.full(k - k // 2, 0.2000001)])
    Qr = np.linalg.qr(rng.randn(k, k))[0]
    yield "eigenvalues clustered at the clamp", (Qr * d) @ Qr.T


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    p = 0.2
    for k in (32, 64, 128):
        print("k = %d" % k)
        for name, H in cases(k, rng):
            H = (H + H.T) / 2
            errs = []
            for _ in range(5):
                g = rng.randn(k)
                ref = safe_invert_apply(H, g, p)
                got = lanczos_clamped_solve(H, g, p)
                errs.append(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
            print("   %-52s max rel err %.1e" % (name, max(errs)))

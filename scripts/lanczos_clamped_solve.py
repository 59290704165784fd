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


def lanczos_clamped_solve(H, g, p, reorth_passes=2, tridiag="ql_first_row"):
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
    if tridiag == "scipy":
        theta, S = eigh_tridiagonal(alpha[:m], beta[:m - 1])
        y = S @ (S[0] / np.maximum(np.abs(theta), p))    # f(T) e_1
    else:
        y = f_of_tridiagonal_e1(alpha[:m], beta[:m - 1], lambda t: 1.0 / np.maximum(np.abs(t), p))
    return nrm * (Q[:, :m] @ y)


def f_of_tridiagonal_e1(diag, off, f):
    """f(T) e_1 for the symmetric tridiagonal T = tridiag(off, diag, off) WITHOUT forming its eigenvectors: implicit QL
    (EISPACK tql2 / Numerical Recipes tqli) on (diag, off) that only carries the FIRST ROW s1 of the eigenvector matrix
    S = G_1 G_2 ... G_N through the Givens rotations and records them; then f(T) e_1 = S (f(theta) * s1) is obtained by
    replaying the rotations in reverse on the vector z = f(theta) * s1.  O(1) state per rotation: the form a single GPU
    thread can run per matrix."""
    n = len(diag)
    d = np.array(diag, dtype=np.float64)
    e = np.zeros(n)
    e[:n - 1] = off
    row0 = np.zeros(n)
    row0[0] = 1.0
    rots = []                                        # (i, c, s): columns i, i + 1 of S
    eps = np.finfo(np.float64).eps
    for l in range(n):
        for sweep in range(60):
            m = l
            while m < n - 1:
                if abs(e[m]) <= eps * (abs(d[m]) + abs(d[m + 1])):
                    break
                m += 1
            if m == l:
                break
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[m] - d[l] + e[l] / (g + (r if g >= 0 else -r))
            sn = cs = 1.0
            pp = 0.0
            broke = False
            for i in range(m - 1, l - 1, -1):
                ff, b = sn * e[i], cs * e[i]
                r = np.hypot(ff, g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= pp
                    e[m] = 0.0
                    broke = True
                    break
                sn, cs = ff / r, g / r
                g = d[i + 1] - pp
                r = (d[i] - g) * sn + 2.0 * cs * b
                pp = sn * r
                d[i + 1] = g + pp
                g = cs * r - b
                a0, a1 = row0[i], row0[i + 1]
                row0[i + 1] = sn * a0 + cs * a1
                row0[i] = cs * a0 - sn * a1
                rots.append((i, cs, sn))
            if not broke:
                d[l] -= pp
                e[l] = g
                e[m] = 0.0
    z = f(d) * row0
    for i, cs, sn in reversed(rots):                 # y = G_1 (G_2 (... (G_N z)))
        a0, a1 = z[i], z[i + 1]
        z[i] = cs * a0 + sn * a1
        z[i + 1] = -sn * a0 + cs * a1
    return z


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
                got2 = lanczos_clamped_solve(H, g, p, tridiag="scipy")
                errs.append((np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300),
                             np.linalg.norm(got2 - ref) / max(np.linalg.norm(ref), 1e-300)))
            print("   %-52s max rel err: QL with first-row tracking %.1e, scipy eigh_tridiagonal %.1e" % (
                name, max(e[0] for e in errs), max(e[1] for e in errs)))

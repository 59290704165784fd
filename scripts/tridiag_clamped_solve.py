"""Prototype: clamped solve x = S(H) g via Householder tridiagonalisation + Sturm bisection + inverse iteration."""
import numpy as np
EPS = np.finfo(float).eps

def tridiagonalise(A):
    """Householder reduction of symmetric A (lower triangle used) -> d, e (e[i] couples i-1,i), reflectors (list of (v, beta))."""
    A = A.copy(); k = A.shape[0]
    refl = []
    for j in range(k - 2):
        x = A[j+1:, j].copy()
        alpha = -np.sign(x[0] if x[0] != 0 else 1.0) * np.linalg.norm(x)
        if alpha == 0.0:
            refl.append(None); continue
        v = x.copy(); v[0] -= alpha
        vn2 = v @ v
        if vn2 == 0.0:
            refl.append(None); continue
        beta = 2.0 / vn2
        S = A[j+1:, j+1:]
        p = beta * (S @ v)
        K = 0.5 * beta * (v @ p)
        q = p - K * v
        S -= np.outer(v, q) + np.outer(q, v)
        A[j+1, j] = alpha; A[j+2:, j] = 0; A[j, j+1] = alpha; A[j, j+2:] = 0
        refl.append((v, beta))
    d = np.diag(A).copy(); e = np.zeros(k); e[1:] = np.diag(A, -1)
    return d, e, refl

def apply_Qt(refl, g):   # y = Q^T g ; Q = H_0 H_1 ... (each acts on trailing part)
    y = g.copy()
    for j, r in enumerate(refl):
        if r is None: continue
        v, beta = r
        y[j+1:] -= beta * (v @ y[j+1:]) * v
    return y

def apply_Q(refl, y):
    x = y.copy()
    for j in reversed(range(len(refl))):
        r = refl[j]
        if r is None: continue
        v, beta = r
        x[j+1:] -= beta * (v @ x[j+1:]) * v
    return x

def sturm_count(d, e2, x, pivmin):
    """number of eigenvalues < x"""
    cnt = 0; q = d[0] - x
    if abs(q) < pivmin: q = -pivmin
    if q < 0: cnt += 1
    for i in range(1, len(d)):
        q = d[i] - x - e2[i] / q
        if abs(q) < pivmin: q = -pivmin
        if q < 0: cnt += 1
    return cnt

def bisect_eig(d, e2, idx, lo, hi, pivmin, tnorm):
    """idx-th smallest eigenvalue (0-based) inside [lo, hi]"""
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if hi - lo <= 2 * EPS * max(abs(lo), abs(hi)) + 2 * pivmin or mid == lo or mid == hi: break
        if sturm_count(d, e2, mid, pivmin) > idx: hi = mid
        else: lo = mid
    return 0.5 * (lo + hi)

def inverse_iteration(d, e, lam, start, tnorm, prev=None):
    """eigenvector of T for eigenvalue lam: LU with partial pivoting of T - lam I (tridiagonal), a few solves."""
    k = len(d)
    # factor: rows a (diag), b (super), c (second super after pivoting), l multipliers, piv flags
    a = d - lam; b = e[1:].copy() if k > 1 else np.zeros(0)   # b[i] = T[i,i+1]
    sub = e[1:].copy()
    u0 = np.zeros(k); u1 = np.zeros(k); u2 = np.zeros(k); l = np.zeros(k); piv = np.zeros(k, bool)
    # Gaussian elimination with partial pivoting
    r0 = a[0]; r1 = b[0] if k > 1 else 0.0; r2 = 0.0
    tiny = EPS * tnorm
    for i in range(k - 1):
        s0 = sub[i]; s1 = a[i+1]; s2 = b[i+1] if i + 2 < k else 0.0   # next row: [s0, s1, s2] at cols i, i+1, i+2
        if abs(s0) > abs(r0):
            piv[i] = True
            u0[i], u1[i], u2[i] = s0, s1, s2
            m = r0 / s0
            l[i] = m
            r0, r1, r2 = r1 - m * s1, r2 - m * s2, 0.0
        else:
            if r0 == 0.0: r0 = tiny
            u0[i], u1[i], u2[i] = r0, r1, r2
            m = s0 / r0
            l[i] = m
            r0, r1, r2 = s1 - m * r1, s2 - m * r2, 0.0
    if r0 == 0.0: r0 = tiny
    u0[k-1] = r0
    x = start.copy()
    for it in range(5):
        # forward (apply L^-1 P) only after first iteration (standard: first iterate solves U x = start)
        if it > 0:
            for i in range(k - 1):
                if piv[i]:
                    x[i], x[i+1] = x[i+1], x[i] - l[i] * x[i+1]
                else:
                    x[i+1] -= l[i] * x[i]
        # back substitution U x = rhs
        for i in reversed(range(k)):
            t = x[i]
            if i + 1 < k: t -= u1[i] * x[i+1]
            if i + 2 < k: t -= u2[i] * x[i+2]
            x[i] = t / u0[i]
        if prev is not None and len(prev):
            for z in prev: x -= (z @ x) * z
        nrm = np.linalg.norm(x)
        x /= nrm
        if it >= 1 and nrm > 1e3 / max(tiny, 1e-300) ** 0: pass
        if it >= 2: break
    return x

def clamped_solve(H, g, p):
    k = H.shape[0]
    d, e, refl = tridiagonalise(H)
    y = apply_Qt(refl, g)
    e2 = e * e
    tnorm = max(np.abs(d) + np.abs(e) + np.abs(np.append(e[1:], 0)))
    pivmin = max(np.finfo(float).tiny * max(e2.max(), 1.0), 1e-290)
    n_neg = sturm_count(d, e2, -p, pivmin)        # eigenvalues < -p
    n_le = sturm_count(d, e2, p, pivmin)          # eigenvalues < p
    idxs = list(range(n_neg)) + list(range(n_le, k))
    gl, gu = -tnorm - 1, tnorm + 1
    lams = [bisect_eig(d, e2, i, gl, gu, pivmin, tnorm) for i in idxs]
    rng = np.random.RandomState(1)
    out = y / p
    prev, prev_l = [], None
    Z = []
    for lam in lams:
        if prev_l is None or abs(lam - prev_l) > 1e-3 * tnorm: prev = []
        z = inverse_iteration(d, e, lam, rng.rand(k) + 0.5, tnorm, prev)
        prev.append(z); prev_l = lam
        Z.append(z)
        out += (1.0 / abs(lam) - 1.0 / p) * (z @ y) * z
    return apply_Q(refl, out), len(lams)

def ref(H, g, p):
    w, Q = np.linalg.eigh(H); w = np.abs(w); w = np.where(w < p, p, w)
    return (Q / w) @ (Q.T @ g)

if __name__ == "__main__":
    worst = 0
    for k in (2, 7, 33, 64, 128):
        rng = np.random.RandomState(k)
        for b in range(12):
            r = max(1, (b * k) // 9) if b < 9 else k
            A = rng.randn(k, min(r, 3 * k)); H = A @ A.T * (0.05 if b % 3 == 0 else 1.0)
            if b == 4: H -= 0.7 * np.eye(k)
            if b == 5: H += 3.0 * np.eye(k)
            if b == 6: H *= 0.15 / np.linalg.norm(H)
            if b == 9:   # clustered: many equal eigenvalues above p
                Q, _ = np.linalg.qr(rng.randn(k, k)); lam = np.where(np.arange(k) % 2 == 0, 1.0, 0.05); H = (Q * lam) @ Q.T
            if b == 10:  # tight clusters near the clamp
                Q, _ = np.linalg.qr(rng.randn(k, k)); lam = 0.2 + 1e-9 * rng.randn(k); H = (Q * lam) @ Q.T
            if b == 11:  # weighted Gram like C4 (float32 rounded)
                V = 0.3 * np.abs(rng.randn(400, k)); w = 0.25 * rng.rand(400); H = (0.5 * (V * w[:, None]).T @ V).astype(np.float32).astype(float)
            H = 0.5 * (H + H.T)
            g = rng.randn(k)
            x, nl = clamped_solve(H, g, 0.2); xr = ref(H, g, 0.2)
            err = np.abs(x - xr).max() / np.abs(xr).max()
            worst = max(worst, err)
            print(k, b, "n_above", nl, "err %.2e" % err)
    print("worst", worst)

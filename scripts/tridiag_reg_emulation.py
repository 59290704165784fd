"""Thread-level emulation of tri::tridiag_reg (pycmf_b200/csrc/tridiag_solve.cuh): the Householder tridiagonalisation with the
matrix in registers, column-owned by lane pairs.  Every array below is what the kernel keeps (a[tid][i], colbuf, vd, qd, red,
y, W rows as reflector storage) and every phase ends where the kernel has a barrier, so an indexing mistake in the kernel's
mapping shows up here on the CPU.  Checked against a plain Householder reduction (scripts/tridiag_clamped_solve.py) and
against eigvalsh; run by tests/test_lanczos_prototype.py."""
import numpy as np


def tridiag_reg_emulated(H, g, KR):
    k, HN = KR, KR // 2
    VP, NB, NW = HN + 2, HN // 8, KR // 16
    NT = 2 * KR
    tid = np.arange(NT)
    lane, warp = tid & 31, tid >> 5
    c, h = tid >> 1, tid & 1
    W = H.copy().reshape(-1)
    colbuf = np.full(KR, np.nan)
    vd = np.full(2 * VP, np.nan)
    qd = np.full(2 * VP, np.nan)
    red = np.zeros(64)
    y = g.copy()
    d, e, beta = np.zeros(k), np.zeros(k), np.zeros(k)
    a = np.empty((NT, HN))
    for i in range(HN):
        a[:, i] = W[(2 * i + h) * k + c]
    colbuf[c[h == 0]] = a[h == 0, 0]

    def publish(j, jb, who):
        """st_shared_if_eq over block jb and element 0 of block jb + 1: the lanes `who` with h == (j + 1) % 2 store a[(j + 1) / 2]."""
        selx = np.where(h == ((j + 1) & 1), (j + 1) >> 1, -1)
        cand = [8 * jb + u for u in range(8)] + ([8 * (jb + 1)] if jb + 1 < NB else [])
        for idx in cand:
            m = who & (selx == idx)
            colbuf[c[m]] = a[m, idx]

    for jb in range(NB):
        jend = 14 if jb == NB - 1 else 16
        for jj in range(jend):
            j = 16 * jb + jj
            # (A) -- norm, warp-redundant: all warps see the same colbuf
            r = np.arange(KR)
            s = float(np.sum(np.where(r > j, colbuf, 0.0) ** 2))
            x0, dj = colbuf[j + 1], colbuf[j]
            tail2 = s - x0 * x0
            if not (tail2 > 0.0):
                d[j], e[j + 1], beta[j] = dj, x0, 0.0
                publish(j, jb, np.ones(NT, bool))
                continue
            alpha = -np.sqrt(s) if x0 >= 0 else np.sqrt(s)
            v0 = x0 - alpha
            bq = 2.0 / (tail2 + v0 * v0)
            for t in range(k):
                vr = 0.0 if t <= j else (v0 if t == j + 1 else colbuf[t])
                vd[(t & 1) * VP + (t >> 1)] = vr
                if t > j:
                    W[j * k + t] = vr
            d[j], e[j + 1], beta[j] = dj, alpha, bq
            # (B)
            live = 16 * warp + 15 > j
            pc = np.zeros(NT)
            for b in range(jb, NB):
                for u in range(8):
                    pc += np.where(live, a[:, 8 * b + u] * vd[h * VP + 8 * b + u], 0.0)
            pc = pc + pc[tid ^ 1]
            vc = vd[(c & 1) * VP + (c >> 1)]
            qc = np.where(c > j, bq * pc, 0.0)
            vp = np.where(h == 0, vc * qc, 0.0)
            vy = np.where(h == 0, vc * y[c], 0.0)
            for w in range(NW):
                red[w] = vp[warp == w].sum()
                red[32 + w] = vy[warp == w].sum()
            # (C)
            svp, svy = red[:NW].sum(), red[32:32 + NW].sum()
            Kc = 0.5 * bq * svp
            qc = qc - Kc * vc
            m0 = h == 0
            qd[(c[m0] & 1) * VP + (c[m0] >> 1)] = qc[m0]
            y[c[m0]] -= bq * svy * vc[m0]
            # (D)
            for b in range(jb, NB):
                for u in range(8):
                    i = 8 * b + u
                    upd = a[:, i] - vd[h * VP + i] * qc - qd[h * VP + i] * vc
                    a[:, i] = np.where(live, upd, a[:, i])
            publish(j, jb, live)
    W[(k - 2 + h) * k + c] = a[:, HN - 1]
    # the shared tail of clamped_solve
    e[0] = 0.0
    d[k - 2] = W[(k - 2) * k + k - 2]
    e[k - 1] = W[(k - 2) * k + k - 1]
    beta[k - 2] = 0.0
    d[k - 1] = W[(k - 1) * k + k - 1]
    beta[k - 1] = 0.0
    return d, e, beta, W.reshape(k, k), y


def check(KR, seed=0, kind="gram"):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tridiag_clamped_solve as P
    rng = np.random.RandomState(seed)
    if kind == "gram":
        A = rng.randn(KR, KR // 3)
        H = A @ A.T + 0.01 * np.eye(KR)
    elif kind == "blockdiag":                         # decoupled blocks: steps with nothing to annihilate
        H = np.zeros((KR, KR))
        for s0 in range(0, KR, 8):
            B = rng.randn(8, 8)
            H[s0:s0 + 8, s0:s0 + 8] = B + B.T
        H[3, 4:] = 0.0
        H[4:, 3] = 0.0
    else:
        B = rng.randn(KR, KR)
        H = B + B.T
    g = rng.randn(KR)
    d, e, beta, W, y = tridiag_reg_emulated(H, g, KR)
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    ev_err = np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(H)).max() / np.abs(np.linalg.eigvalsh(H)).max()
    # the reflectors stored in W reproduce y = P^T g and P T P^T = H
    Q = np.eye(KR)
    for j in range(KR - 2):
        if beta[j] == 0.0:
            continue
        v = np.zeros(KR)
        v[j + 1:] = W[j, j + 1:]
        Q = Q @ (np.eye(KR) - beta[j] * np.outer(v, v))
    rec_err = np.abs(Q @ T @ Q.T - H).max() / np.abs(H).max()
    y_err = np.abs(Q.T @ g - y).max() / np.abs(g).max()
    d0, e0, _ = P.tridiagonalise(H)
    ref_err = max(np.abs(np.abs(d0) - np.abs(d)).max(), np.abs(np.abs(e0) - np.abs(e)).max()) / np.abs(H).max()
    return ev_err, rec_err, y_err, ref_err


if __name__ == "__main__":
    for KR in (64, 128):
        for kind in ("gram", "indef", "blockdiag"):
            print(KR, kind, ["%.1e" % v for v in check(KR, 1, kind)])

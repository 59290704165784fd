"""Factor initialisation ON THE GPU with the reference's semantics (pycmf/cmf.py:41-202).

`init.py` is the host restatement (NumPy + scikit-learn's `randomized_svd`); at the toxic-comments scale (2M x 200k)
that host SVD takes far longer than the whole GPU fit.  Here the same algorithms run on the device-resident, already
ingested matrix: every product with M (the only O(nnz k) / O(n d k) work: `M Q`, `M^T Q`, `Q^T M`) goes through the
library's own GEMM / SpMM kernels, the thin (rows x (k + 10)) LU / QR factorisations and the small SVD through
torch.linalg (cuSOLVER: plumbing on factor-sized operands), and the NNDSVD post-processing is elementwise.

Random numbers are drawn from the SAME NumPy streams as the host path (`check_random_state(random_state)`: the Gaussian
test matrix of the range finder, the 'random' factors, the 'nndsvdar' fill), so both paths start from identical draws and
agree to rounding (tests/test_gpu_api.py::test_device_initialisation_matches_the_host_path).

Algorithm of the randomized SVD (Halko et al. 2009, as scikit-learn's `randomized_svd` runs it with its defaults:
10 oversamples, 7 power iterations when k < 0.1 min(shape) else 4, LU-normalised power iterations, QR at the end,
transposed when rows < cols, deterministic sign flip on the left vectors of the problem as solved).
"""
import warnings

import numpy as np
from sklearn.utils import check_random_state

from .init import NNDSVD_KINDS

OVERSAMPLES = 10
MAX_WIDTH = 256          # the SpMM / GEMM kernels take at most 256 right-hand columns per call


class _Op:
    """M or M^T of an ingested matrix as a linear operator on device tensors."""

    def __init__(self, be, M, transposed=False):
        self.be, self.M, self.transposed = be, M, transposed
        r, c = M.shape
        self.shape = (c, r) if transposed else (r, c)

    @property
    def T(self):
        return _Op(self.be, self.M, not self.transposed)

    def matmul(self, Q):
        """op(M) @ Q for Q (cols x p), in column blocks of at most MAX_WIDTH."""
        be = self.be
        torch = be.torch
        outs = []
        for c0 in range(0, Q.shape[1], MAX_WIDTH):
            Qb = Q[:, c0:c0 + MAX_WIDTH].contiguous()
            if self.M.is_sparse:
                outs.append(be.spmm(self.M, Qb, transposed=self.transposed))
            else:
                outs.append(be.gemm(self.M.t, Qb, trans_a=self.transposed))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)

    def mean(self):
        be = self.be
        if self.M.is_sparse:
            return float(self.M.vals.sum(dtype=be.torch.float64)) / (self.M.shape[0] * self.M.shape[1])
        return float(self.M.t.sum(dtype=be.torch.float64)) / (self.M.shape[0] * self.M.shape[1])


def _lu_normalise(torch, A):
    """P L of the pivoted LU of a tall matrix (scipy.linalg.lu(permute_l=True)): the power iterations' normaliser."""
    P, L, _ = torch.linalg.lu(A)
    return P @ L


def randomized_svd_device(be, op, k, random_state):
    """(U (rows x k), s (k), Vt (k x cols)) of op ~ U diag(s) Vt."""
    torch = be.torch
    rng = check_random_state(random_state)
    p = k + OVERSAMPLES
    n_iter = 7 if k < 0.1 * min(op.shape) else 4
    transpose = op.shape[0] < op.shape[1]
    A = op.T if transpose else op
    Q = be.to_device(rng.normal(size=(A.shape[1], p)))
    for _ in range(n_iter):
        Q = _lu_normalise(torch, A.matmul(Q))
        Q = _lu_normalise(torch, A.T.matmul(Q))
    Q, _ = torch.linalg.qr(A.matmul(Q), mode="reduced")
    B = A.T.matmul(Q).T.contiguous()                               # Q^T A, p x cols
    Uhat, s, Vt = torch.linalg.svd(B, full_matrices=False)
    U = Q @ Uhat
    # deterministic signs: the entry of largest magnitude of every LEFT singular vector of M is positive -- these are the
    # columns of U, or, when the transposed problem was solved, the rows of Vt
    ar = torch.arange(U.shape[1], device=U.device)
    if transpose:
        signs = torch.sign(Vt[ar, torch.argmax(Vt.abs(), dim=1)])
    else:
        signs = torch.sign(U[torch.argmax(U.abs(), dim=0), ar])
    signs = torch.where(signs == 0, torch.ones_like(signs), signs)
    U, Vt = U * signs[None, :], Vt * signs[:, None]
    if transpose:
        return Vt[:k].T.contiguous(), s[:k], U[:, :k].T.contiguous()
    return U[:, :k].contiguous(), s[:k], Vt[:k].contiguous()


def _random_init(be, op, k, random_state, non_negative):
    scale = np.sqrt(abs(op.mean()) / k)                            # cmf.py:111
    rng = check_random_state(random_state)
    A = be.to_device(scale * rng.randn(op.shape[0], k))
    Bt = be.to_device(scale * rng.randn(k, op.shape[1]))
    if non_negative:
        A, Bt = A.abs(), Bt.abs()
    return A, Bt


def _svd_init(be, op, k, random_state):
    torch = be.torch
    rows, cols = op.shape
    if min(rows, cols) < k:
        warnings.warn('The number of components is smaller than the rank in svd initialization.' +
                      'The input will be padded with zeros to compensate for the lack of singular values.')
    Us, s, Vt = randomized_svd_device(be, op, k, random_state)
    if k > cols:                                                   # pad to the requested width (cmf.py:129-138)
        Us = torch.cat([Us, Us.new_zeros(Us.shape[0], k - Us.shape[1])], 1)
        Vt = torch.cat([Vt, Vt.new_zeros(k - Vt.shape[0], Vt.shape[1])], 0)
        s = torch.cat([s, s.new_zeros(k - s.shape[0])])
    root = torch.sqrt(s)
    return Us * root[None, :], root[:, None] * Vt


def _nndsvd_init(be, op, k, kind, eps, random_state, non_negative):
    """Boutsidis & Gallopoulos (2008): the positive or the negative parts of every singular pair, whichever carry more
    mass (cmf.py:145-198), all k pairs at once."""
    torch = be.torch
    if not non_negative:
        warnings.warn('%s results in non-negative constrained factors,' % kind +
                      'so SVD initialization should provide better initial estimate')
    Us, s, Vt = randomized_svd_device(be, op, k, random_state)
    V = Vt.T
    xp, xn = Us.clamp(min=0), (-Us).clamp(min=0)
    yp, yn = V.clamp(min=0), (-V).clamp(min=0)
    xp_n, xn_n = torch.linalg.vector_norm(xp, dim=0), torch.linalg.vector_norm(xn, dim=0)
    yp_n, yn_n = torch.linalg.vector_norm(yp, dim=0), torch.linalg.vector_norm(yn, dim=0)
    pos, neg = xp_n * yp_n, xn_n * yn_n
    take_pos = pos > neg
    u = torch.where(take_pos[None, :], xp / xp_n[None, :], xn / xn_n[None, :])
    v = torch.where(take_pos[None, :], yp / yp_n[None, :], yn / yn_n[None, :])
    scale = torch.sqrt(s * torch.where(take_pos, pos, neg))
    A, B = u * scale[None, :], v * scale[None, :]
    # the leading pair is sign-definite: |u_0|, |v_0| scaled by sqrt(s_0)
    A[:, 0] = torch.sqrt(s[0]) * Us[:, 0].abs()
    B[:, 0] = torch.sqrt(s[0]) * V[:, 0].abs()
    A = torch.nan_to_num(A, nan=0.0)
    B = torch.nan_to_num(B, nan=0.0)
    Bt = B.T.contiguous()
    A[A < eps] = 0
    Bt[Bt < eps] = 0
    if kind == "nndsvda":
        avg = op.mean()
        A[A == 0] = avg
        Bt[Bt == 0] = avg
    elif kind == "nndsvdar":
        rng = check_random_state(random_state)
        avg = op.mean()
        za, zb = A == 0, Bt == 0
        na, nb = int(za.sum()), int(zb.sum())
        A[za] = be.to_device(np.abs(avg * rng.randn(na) / 100), dtype=be.np_dtype)
        Bt[zb] = be.to_device(np.abs(avg * rng.randn(nb) / 100), dtype=be.np_dtype)
    return A, Bt


def initialize_mf_device(be, M, n_components, init=None, eps=1e-6, random_state=None, non_negative=False):
    """Initial guess M ~= A B^T on the device: (A rows x k, B cols x k) as device tensors of the compute dtype.
    `M` is an ingested DenseMatrix / SparseMatrix; same defaults and errors as the reference (cmf.py:41-202)."""
    op = _Op(be, M)
    if non_negative:
        vals = M.vals if M.is_sparse else M.t
        if bool((vals < 0).any()):
            raise ValueError("Negative values in data passed to MF initialization")
    cols = op.shape[1]
    if init is None:
        init = ('nndsvdar' if non_negative else 'svd') if n_components < cols else 'random'
    if init == 'random':
        A, Bt = _random_init(be, op, n_components, random_state, non_negative)
    elif init == 'svd':
        if non_negative:
            raise ValueError('SVD initialization incompatible with NMF (use nndsvd instead)')
        A, Bt = _svd_init(be, op, n_components, random_state)
    elif init in NNDSVD_KINDS:
        A, Bt = _nndsvd_init(be, op, n_components, init, eps, random_state, non_negative)
    else:
        raise ValueError("Invalid init argument")
    return A.contiguous(), Bt.T.contiguous()

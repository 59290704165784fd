"""A float64 NumPy stand-in for CudaBackend, for CPU-only tests of the HOST orchestration (phase
decomposition, row sharding, collectives under gloo).  Test infrastructure: it computes every phase with
the oracle's formulas on torch CPU tensors so torch.distributed(gloo) can all-reduce them."""
import numpy as np
import scipy.sparse as sp
import torch

from oracle import cmf_oracle as O


class M:
    def __init__(self, a):
        self.a = a.tocsr() if sp.issparse(a) else np.asarray(a, dtype=np.float64)
        self.is_sparse = sp.issparse(a)
        self.shape = a.shape


class FakeBackend:
    torch = torch
    device = torch.device("cpu")
    tdtype = torch.float64
    np_dtype = np.dtype("float64")

    def to_device(self, a, dtype=None):
        a = np.array(a, dtype=np.float64 if dtype is None else dtype)
        return torch.from_numpy(a)

    def to_host(self, t):
        return t.numpy().copy()

    def zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=dtype or torch.float64)

    def ingest(self, A):
        return A if isinstance(A, M) else M(A)

    def row_slice(self, T, r0, r1):
        return M(T.a[r0:r1])

    def col_slice(self, T, c0, c1):
        return M(T.a[:, c0:c1])

    def column_block(self, T, comm, r0, col_ranges):
        """Test stand-in of CudaBackend.column_block: every rank's row shard is gathered as an object."""
        shards = [None] * comm.world
        comm.dist.all_gather_object(shards, T.a, group=comm.group)
        c0, c1 = col_ranges[comm.rank]
        full = sp.vstack(shards).tocsr() if T.is_sparse else np.vstack(shards)
        return M(full[:, c0:c1])

    def synchronize(self):
        pass

    @staticmethod
    def _dense(T):
        return T.a.toarray() if T.is_sparse else T.a

    def sqerr(self, A, B, T, link, trans=False):
        t = T.a.T if trans else T.a
        e = O.compute_factorization_error(t, A.numpy(), B.numpy().T, link)
        return torch.tensor([e * e], dtype=torch.float64)

    # MU
    def mu_v_partial(self, X, U):
        u = U.numpy()
        return torch.from_numpy(np.vstack([np.asarray(X.a.T @ u), u.T @ u]))

    def mu_v_apply(self, V, buf, Y, Z, l1, l2):
        v, z, b = V.numpy(), Z.numpy(), buf.numpy()
        d = v.shape[0]
        num = b[:d] + Y.a @ z
        den = v @ (b[d:] + z.T @ z)
        v *= O._regularized_delta(num, den, l1, l2, v)

    def mu_left(self, F, B, T, l1, l2, trans=False):
        f, b = F.numpy(), B.numpy()
        t = T.a.T if trans else T.a
        f *= O._regularized_delta(np.asarray(t @ b), f @ (b.T @ b), l1, l2, f)

    # Newton
    def newton_left(self, F, B, T, weight, l1, l2, link, non_negative, pert, l2_in_logit_hessian, idx=None,
                    trans=False):
        t = T.a.T if trans else T.a
        O._rows_newton(F.numpy(), B.numpy(), t, weight, l1, l2, link, non_negative, pert,
                       l2_in_logit_hessian, idx=None if idx is None else idx.numpy())

    def newton_v_needs_per_row(self, x_link, sampled):
        return bool(sampled) or x_link == "logit"

    def v_chunk_rows(self, d, k, per_row, budget_bytes=0):
        return max(1, d // 2) if per_row else d     # force chunking in the tests

    def newton_v_xpart(self, V, U, X, j0, j1, x_link, alpha, idx=None):
        v, u = V.numpy()[j0:j1], U.numpy()
        xt = self._dense(X)[:, j0:j1].T                       # (rows, n_local)
        k = v.shape[1]
        per_row = self.newton_v_needs_per_row(x_link, idx is not None)
        if idx is None:
            est = v @ u.T
            gx = alpha * ((O.inverse(est, x_link) - xt) @ u)
            if x_link == "logit":
                Hx = alpha * np.einsum('ci,ia,ib->cab', O.d_sigmoid(est), u, u)
            else:
                Hx = alpha * (u.T @ u)[None]
        else:
            ix = idx.numpy()
            gx, Hx = np.zeros((v.shape[0], k)), np.zeros((v.shape[0], k, k))
            for r in range(v.shape[0]):
                sel = ix[r][ix[r] >= 0]
                us = u[sel]
                est = us @ v[r]
                w = O.d_sigmoid(est) if x_link == "logit" else np.ones_like(est)
                gx[r] = alpha * (O.inverse(est, x_link) - xt[r, sel]) @ us
                Hx[r] = alpha * (us * w[:, None]).T @ us
        return torch.from_numpy(gx), torch.from_numpy(np.ascontiguousarray(Hx)), per_row

    def newton_v_finish(self, V, Z, Y, j0, j1, y_link, alpha, l1, l2, gx, Hx, per_row, non_negative, pert, idx=None):
        v, z, y = V.numpy()[j0:j1], Z.numpy(), Y.a[j0:j1]
        k = v.shape[1]
        g, H = gx.numpy().copy(), np.broadcast_to(Hx.numpy(), (v.shape[0], k, k)).copy()
        for r in range(v.shape[0]):
            sel = np.arange(z.shape[0]) if idx is None else idx.numpy()[r]
            zs = z[sel]
            est = zs @ v[r]
            w = O.d_sigmoid(est) if y_link == "logit" else np.ones_like(est)
            g[r] += (1 - alpha) * (O.inverse(est, y_link) - y[r, sel]) @ zs
            H[r] += (1 - alpha) * (zs * w[:, None]).T @ zs + l2 * np.eye(k)
        step = g + l1 * np.sign(v) + l2 * v
        new = v - np.einsum('ca,cab->cb', step, O.safe_invert(H, pert))
        if non_negative:
            new = np.where(new < 0, 0., new)
        V.numpy()[j0:j1] = new

"""B200 solvers behind the reference's solver seam.

Same classes, constructor arguments, methods and error behaviour as the reference's
pycmf/cmf_solvers.py (`_IterativeCMFSolver` :45-195, `MUSolver` :198-263, `NewtonSolver` :319-522):
`fit_iterative_update(X, Y, U, V, Z) -> (U, V, Z, n_iter)` updates the caller's factor arrays in
place, `update_step` / `compute_error` are the sub-seams.  The arithmetic runs in libpycmf_b200.so;
this file only orchestrates phases, collectives and the convergence test (the host loop the
reference also keeps in Python, :170-187).

Backend-only knobs (extra keyword arguments, all optional): `dtype` ('float32' | 'float64'),
`device`, `comm` (sharding.Comm), `sampler` ('numpy' | 'device' | 'auto'), `backend_options`.
"""
import numbers
import time
import warnings

import numpy as np
import scipy.sparse as sp

from .sharding import Comm, default_comm, localize_indices, row_range

EPSILON = np.finfo(np.float32).eps
MAX_COMPONENTS = 256
INTEGER_TYPES = (numbers.Integral, np.integer)


def _beta_loss_to_float(beta_loss):
    """sklearn.decomposition._nmf._beta_loss_to_float (used at cmf_solvers.py:106)."""
    table = {"frobenius": 2, "kullback-leibler": 1, "itakura-saito": 0}
    if isinstance(beta_loss, str):
        if beta_loss not in table:
            raise ValueError("Invalid beta_loss parameter: got %r instead of one of %r, or a float."
                             % (beta_loss, list(table.keys())))
        return table[beta_loss]
    if not isinstance(beta_loss, numbers.Number):
        raise ValueError("Invalid beta_loss parameter: got %r instead of one of %r, or a float."
                         % (beta_loss, list(table.keys())))
    return beta_loss


class FitState:
    """Everything one rank keeps in HBM during a fit: its row shard of X and U, replicated Y / V / Z."""

    def __init__(self, backend, comm, X, Y, U, V, Z, n_total, rows, Xcol=None, cols=None):
        self.be, self.comm = backend, comm
        self.X, self.Y = X, Y
        self.U, self.V, self.Z = U, V, Z
        self.n_total = n_total
        self.r0, self.r1 = rows
        # column block [c0, c1) of X over ALL rows (second copy; the column-sharded Newton V phase)
        self.Xcol = Xcol
        self.c0, self.c1 = cols if cols is not None else (0, 0)
        self.row_counts = None      # rows held by every rank when the caller handed in its own blocks (sharded_input)
        self.iteration = 0

    @property
    def shapes(self):
        return self.n_total, self.V.shape[0], self.Z.shape[0], self.V.shape[1]


class _IterativeCMFSolver:
    """Boilerplate for the iterative solvers (reference cmf_solvers.py:45-195)."""

    def __init__(self, max_iter=200, tol=1e-4, beta_loss="frobenius",
                 l1_reg=0, l2_reg=0, alpha=0.5, verbose=0,
                 U_non_negative=True, V_non_negative=True, Z_non_negative=True,
                 update_U=True, update_V=True, update_Z=True,
                 x_link="linear", y_link="linear", hessian_pertubation=0.2,
                 sg_sample_ratio=1., random_state=None,
                 dtype="float32", device=None, comm=None, sampler="auto", backend=None, backend_options=None,
                 sharded_input=False, use_cuda_graph="auto", v_phase="auto"):
        self.max_iter = max_iter
        self.tol = tol
        self.beta_loss = _beta_loss_to_float(beta_loss)
        self.l1_reg = l1_reg
        self.l2_reg = l2_reg
        self.alpha = alpha
        self.verbose = verbose
        self.U_non_negative = U_non_negative
        self.V_non_negative = V_non_negative
        self.Z_non_negative = Z_non_negative
        self.update_U = update_U
        self.update_V = update_V
        self.update_Z = update_Z
        self.x_link = x_link
        self.y_link = y_link
        self.hessian_pertubation = hessian_pertubation
        self.sg_sample_ratio = sg_sample_ratio
        self.random_state = random_state
        if random_state is not None and isinstance(random_state, INTEGER_TYPES + (np.ndarray, list)):
            np.random.seed(random_state)           # cmf_solvers.py:121-122 (global legacy RNG)
        if self.beta_loss != 2:
            raise NotImplementedError("only the Frobenius loss (beta_loss=2) is implemented, as in the reference")
        self.dtype = np.dtype(dtype)
        self.device = device
        self.comm = comm
        self.sampler = sampler
        self.backend_options = backend_options
        self._backend = backend
        self.sharded_input = sharded_input   # X / U handed in are already this rank's row block
        self.use_cuda_graph = use_cuda_graph # replay one captured iteration (single GPU, no per-iteration host work)
        self.v_phase = v_phase               # Newton V update with per-row Hessians on several ranks: 'rows' | 'columns'
        self.masks_per_iter = None     # test hook: list of per-iteration mask dicts (global indices)
        self.history = None            # test hook: list receiving the objective after every iteration

    # ---- backend / ingest ------------------------------------------------------------------------
    def _get_backend(self):
        if self._backend is None:
            from .device import CudaBackend
            self._backend = CudaBackend(device=self.device, dtype=self.dtype, options=self.backend_options)
        return self._backend

    def prepare(self, X, Y, U, V, Z, column_block=True):
        """Host inputs -> FitState in HBM (row shard of X / U for this rank; `column_block`: also the rank's column block
        of X when the solver's V phase re-partitions -- not needed to evaluate the objective)."""
        k = np.shape(V)[1]
        if k > MAX_COMPONENTS:
            raise ValueError("n_components = %d: the B200 backend keeps a factor row in registers / tensor memory and "
                             "supports n_components <= %d (with n_components=None the reference's default is "
                             "max(X.shape[1], Y.shape[1]): pass n_components explicitly)" % (k, MAX_COMPONENTS))
        be = self._get_backend()
        comm = self.comm if self.comm is not None else default_comm()
        n_local = X.shape[0] if X is not None else np.shape(U)[0]
        if self.sharded_input and comm.world > 1:
            counts = be.torch.zeros(comm.world, dtype=be.torch.int64, device=be.device)
            counts[comm.rank] = n_local
            counts = be.to_host(comm.all_reduce_sum(counts))
            n_total = int(counts.sum())
            r0 = int(counts[:comm.rank].sum())
            r1 = r0 + n_local
            take = slice(None)
            row_counts = [int(c) for c in counts]
        else:
            n_total = n_local
            r0, r1 = row_range(n_total, comm.rank, comm.world)
            take = slice(r0, r1)
            row_counts = None
        Xd = None
        if X is not None:
            if getattr(X, "is_sparse", None) is not None:        # already ingested (device initialisation, bench)
                Xd = X if (r0, r1) == (0, X.shape[0]) or take == slice(None) else be.row_slice(X, r0, r1)
            elif sp.issparse(X):
                Xd = be.ingest(X if take == slice(None) or (r0, r1) == (0, X.shape[0]) else sp.csr_matrix(X)[take])
            else:
                Xd = be.ingest(np.asarray(X)[take])
        Yd = None
        if Y is not None:
            Yd = Y if getattr(Y, "is_sparse", None) is not None else be.ingest(Y.toarray() if sp.issparse(Y) else Y)
        Ud = be.to_device(np.asarray(U)[take])
        Vd = be.to_device(np.asarray(V))
        Zd = be.to_device(np.asarray(Z))
        Xcol, cols = self._prepare_column_block(be, comm, X, Xd, np.shape(V)[0], r0) if column_block else (None, None)
        st = FitState(be, comm, Xd, Yd, Ud, Vd, Zd, n_total, (r0, r1), Xcol, cols)
        st.row_counts = row_counts
        return st

    def _prepare_column_block(self, be, comm, X, Xd, d, r0):
        """Only the Newton solver re-partitions (NewtonSolver._prepare_column_block)."""
        return None, None

    # ---- seam ------------------------------------------------------------------------------------
    def update_step(self, X, Y, U, V, Z, l1_reg, l2_reg, alpha):
        """A single update step for all the matrices in the factorization."""
        raise NotImplementedError("Implement in concrete subclass to use")

    def _step(self, st):
        raise NotImplementedError("Implement in concrete subclass to use")

    def _error_links(self):
        return self.x_link, self.y_link

    def device_error(self, st):
        """alpha ||X - f(UV^T)||_F + (1 - alpha) ||Y - f(VZ^T)||_F  (cmf_solvers.py:128-130)."""
        be = st.be
        x_link, y_link = self._error_links()
        parts = be.zeros(2, dtype=be.torch.float64)
        if st.X is not None:
            parts[0:1] = be.sqerr(st.U, st.V, st.X, x_link)
        st.comm.all_reduce_sum(parts[0:1])
        if st.Y is not None:
            parts[1:2] = be.sqerr(st.V, st.Z, st.Y, y_link)
        ex, ey = np.sqrt(np.maximum(be.to_host(parts), 0.0))
        return float(self.alpha * ex + (1 - self.alpha) * ey)

    def device_error_parts(self, st, x_link, y_link):
        """(||X - f1(U V^T)||_F, ||Y - f2(V Z^T)||_F) on the resident state, with the given links (cmf.py:697-698 evaluates
        the final reconstruction error with the estimator's links, whatever the solver)."""
        be = st.be
        parts = be.zeros(2, dtype=be.torch.float64)
        if st.X is not None:
            parts[0:1] = be.sqerr(st.U, st.V, st.X, x_link)
        st.comm.all_reduce_sum(parts[0:1])
        if st.Y is not None:
            parts[1:2] = be.sqerr(st.V, st.Z, st.Y, y_link)
        ex, ey = np.sqrt(np.maximum(be.to_host(parts), 0.0))
        return float(ex), float(ey)

    def compute_error(self, X, Y, U, V, Z):
        if isinstance(X, FitState):
            return self.device_error(X)
        st = self.prepare(X, Y, U, V, Z, column_block=False)
        return self.device_error(st)

    # 'auto' captures the iteration only for fits long enough to pay for it: capture + instantiate + the graph's private
    # allocations cost 13 - 32 ms (C2 / C3 / C5), the launch gaps they remove ~0.1 ms per iteration
    GRAPH_MIN_ITERS = 64

    def _graphable(self, st, force=False):
        """One iteration can be replayed as a CUDA graph when it needs no host work: no sampling, capturable collectives."""
        if self.use_cuda_graph is False or self.sg_sample_ratio < 1.:
            return False
        if self.use_cuda_graph == "auto" and not force and self.max_iter < self.GRAPH_MIN_ITERS:
            return False
        if st.comm.world > 1 and not getattr(st.comm, "graph_capturable", False):
            return False
        return hasattr(st.be, "capture_step")

    def make_stepper(self, st, force_graph=False):
        """Returns step(): one solver iteration. After two eager iterations (scratch arenas reach their final size)
        the iteration is captured into a CUDA graph and replayed, which removes ~25 launch gaps per iteration
        (`force_graph`: also for short fits -- steady-state timing)."""
        state = {"eager": 0, "graph": None}
        graphable = self._graphable(st, force_graph)

        def step():
            st.iteration += 1
            if state["graph"] is not None:
                state["graph"].replay()
            elif graphable and state["eager"] >= 2:
                try:
                    t0 = time.perf_counter()
                    state["graph"] = st.be.capture_step(lambda: self._step(st))
                    self.capture_seconds_ = time.perf_counter() - t0      # host time of capture + instantiate
                except RuntimeError as e:             # capture not possible (e.g. a collective that cannot be captured)
                    warnings.warn("CUDA-graph capture of the iteration failed, running eagerly: %s" % (e,))
                    state["eager"] = -(10 ** 9)
                    self._step(st)
            else:
                self._step(st)
                state["eager"] += 1
        return step

    def fit_device(self, st):
        """The loop of cmf_solvers.py:165-195 on device-resident state. Returns n_iter."""
        start_time = time.time()
        check = self.tol > 0
        previous_error = error_at_init = self.device_error(st) if check else None
        n_iter = 0
        step = self.make_stepper(st) if self.history is None else None
        for n_iter in range(1, self.max_iter + 1):
            if step is not None:
                step()
            else:
                st.iteration = n_iter
                self._step(st)
            if self.history is not None:
                self.history.append(self.device_error(st))
            if check and n_iter % 10 == 0:
                error = self.device_error(st)
                if self.verbose:
                    print("Epoch %02d reached after %.3f seconds, error: %f" %
                          (n_iter, time.time() - start_time, error))
                if (previous_error - error) / error_at_init < self.tol:
                    break
                previous_error = error
        if self.verbose and (self.tol == 0 or n_iter % 10 != 0):
            st.be.synchronize()
            print("Epoch %02d reached after %.3f seconds." % (n_iter, time.time() - start_time))
        return n_iter

    def fit_iterative_update(self, X, Y, U, V, Z):
        """Compute CMF with iterative methods (reference cmf_solvers.py:132-195).

        X (n x d, ndarray or scipy sparse), Y (d x l), U (n x k), V (d x k), Z (l x k) live on the host;
        the factors are updated in place and returned by identity together with n_iter.
        """
        st = self.prepare(X, Y, U, V, Z)
        n_iter = self.fit_device(st)
        be = st.be
        # final reconstruction errors while X / Y are still resident (CMF.fit_transform, cmf.py:697-698): evaluating them
        # afterwards from the host arrays would upload X a second time (40 GB on C5)
        self.final_errors_ = None
        links = getattr(self, "final_error_links", None)
        if links is not None:
            self.final_errors_ = self.device_error_parts(st, links[0], links[1])
        # only the factors that were updated come back: a factor held fixed (update_V=False in transform(), cmf.py:741)
        # stays bit-identical on the host, as in the reference (tests/test_cmf.py:408) -- a float32 round trip would not
        pairs = []
        if self.update_U:
            pairs.append((U, st.U if self.sharded_input else st.comm.all_gather_rows(st.U, st.n_total)))
        if self.update_V:
            pairs.append((V, st.V))
        if self.update_Z:
            pairs.append((Z, st.Z))
        if pairs and hasattr(be, "to_host_many"):
            for (host, _), got in zip(pairs, be.to_host_many([dev for _, dev in pairs])):
                host[...] = got
        else:
            for host, dev in pairs:
                host[...] = be.to_host(dev)
        return U, V, Z, n_iter


class MUSolver(_IterativeCMFSolver):
    """Multiplicative-update solver (reference cmf_solvers.py:198-263): order V, U, Z; links,
    alpha and the non-negativity flags are ignored exactly as in the reference."""

    # V updates smaller than this many elements stay replicated behind one all-reduce: three collectives cost more than the
    # replicated work saves (C3 slice, d = 25000, k = 64 on 4 GPUs: 0.51 ms replicated, 0.59 ms sharded; C5 width, d = 50000,
    # k = 256: 3.59 -> 3.39 ms)
    SHARD_V_MIN = 1 << 22

    def _error_links(self):
        return "linear", "linear"

    # Slab counts tried by the overlapped V update (first one that divides).  OFF by default (PYCMF_B200_V_SLABS unset):
    # measured on 2 B200s (profiles/r02_overlap_n2.txt) the four slab passes cost more than the hidden collectives save --
    # C5: X^T U 12.3 -> 16.4 ms for ~0.3 ms of exchange; C3: SpMM 0.96 -> 1.82 ms -- see DESIGN section 5.
    V_SLABS = ()

    def _v_slabs(self, st):
        """Number of column slabs of X for the overlapped V update (1 = one pass, no overlap)."""
        import os
        d, k = st.V.shape
        world = st.comm.world
        want = getattr(self, "v_slabs", None) or int(os.environ.get("PYCMF_B200_V_SLABS", "0")) or "auto"
        if world == 1 or d * k < self.SHARD_V_MIN or not getattr(st.comm, "overlap_capable", False) or want == 1:
            return 1
        for s in ((want,) if want != "auto" else self.V_SLABS):
            # every rank gets d / (s * world) rows of every slab; dense slabs must start on a 16-byte boundary (TMA)
            if d % (s * world) == 0 and (d // s) % 4 == 0:
                return s
        return 1

    def _step_v_overlapped(self, st, slabs):
        """The V update (:242-246, :252-255) on several ranks with the exchange hidden behind the pass over X.

        X^T U is computed in `slabs` column slabs of X; while slab i + 1 is in the tensor cores, slab i's partial is
        reduce-scattered by rows of V on a communication stream (NCCL over NVLink).  Every rank then applies the update
        to its d / (slabs * G) rows of each slab (Y Z, V (U^T U + Z^T Z) and the elementwise step shrink by G instead of
        being replicated), and slab i's new rows are all-gathered while slab i + 1 is updated.  What stays exposed is the
        last slab's reduce-scatter and all-gather."""
        be, comm = st.be, st.comm
        torch = be.torch
        d, k = st.V.shape
        world, rank = comm.world, comm.rank
        rows_s = d // slabs
        piece = rows_s // world
        ws = getattr(st, "_v_overlap", None)
        if ws is None or ws["slabs"] != slabs:
            ws = st._v_overlap = dict(
                slabs=slabs, stream=torch.cuda.Stream(device=be.device),
                part=[be.empty(rows_s + k, k) for _ in range(slabs)],
                num=[be.empty(piece + k, k) for _ in range(slabs)],      # [reduced numerator rows ; U^T U]
                send=[be.empty(piece, k) for _ in range(slabs)])
        cs = ws["stream"]
        main = be.stream
        cs.wait_stream(main)
        reduced = []
        for s in range(slabs):
            be.mu_v_partial(be.col_slice(st.X, s * rows_s, (s + 1) * rows_s), st.U, out=ws["part"][s])
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(cs):
                cs.wait_event(ev)
                if s == 0:
                    comm.all_reduce_sum(ws["part"][0][rows_s:])                       # U^T U, k x k
                comm.reduce_scatter_rows(ws["part"][s][:rows_s], out=ws["num"][s][:piece])
                done = torch.cuda.Event()
                done.record(cs)
            reduced.append(done)
        for s in range(slabs):
            main.wait_event(reduced[s])
            ws["num"][s][piece:].copy_(ws["part"][0][rows_s:])
            j0 = s * rows_s + rank * piece
            v_loc = st.V[j0:j0 + piece]
            be.mu_v_apply(v_loc, ws["num"][s], be.row_slice(st.Y, j0, j0 + piece), st.Z, self.l1_reg, self.l2_reg)
            ws["send"][s].copy_(v_loc)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(cs):
                cs.wait_event(ev)
                comm.all_gather_into(st.V[s * rows_s:(s + 1) * rows_s], ws["send"][s])
        main.wait_stream(cs)

    def _step(self, st):
        be = st.be
        slabs = self._v_slabs(st) if self.update_V else 1
        if self.update_V and slabs > 1:
            self._step_v_overlapped(st, slabs)
        elif self.update_V:                                      # :252-255
            buf = be.mu_v_partial(st.X, st.U)                    # [X^T U ; U^T U] of this shard
            d, k = st.V.shape
            world = st.comm.world
            if world > 1 and d % world == 0 and d * k >= self.SHARD_V_MIN:
                # Rows of V are independent in the update: the numerator partial is reduce-scattered by rows of V, every
                # rank applies the update to its d / G rows (Y Z, V (U^T U + Z^T Z) and the elementwise step shrink by G
                # instead of being replicated) and the new rows are all-gathered -- the two collectives together move
                # what the all-reduce moved.
                gram = st.comm.all_reduce_sum(buf[d:])
                num = st.comm.reduce_scatter_rows(buf[:d])
                j0 = st.comm.rank * (d // world)
                j1 = j0 + d // world
                v_loc = st.V[j0:j1]
                be.mu_v_apply(v_loc, be.torch.cat([num, gram], 0), be.row_slice(st.Y, j0, j1), st.Z,
                              self.l1_reg, self.l2_reg)
                st.comm.all_gather_into(st.V, v_loc.clone())
            else:
                st.comm.all_reduce_sum(buf)
                be.mu_v_apply(st.V, buf, st.Y, st.Z, self.l1_reg, self.l2_reg)
        # U and Z both read the new V and nothing of each other: the Z update runs on the auxiliary stream
        side = be.fork() if (self.update_U and self.update_Z and hasattr(be, "fork")) else be
        if self.update_Z:                                        # :261-263
            side.mu_left(st.Z, st.V, st.Y, self.l1_reg, self.l2_reg, trans=True)
        if self.update_U:                                        # :257-259
            be.mu_left(st.U, st.V, st.X, self.l1_reg, self.l2_reg)
        if side is not be:
            be.join()

    def update_step(self, X, Y, U, V, Z, l1_reg, l2_reg, alpha):
        st = X if isinstance(X, FitState) else self.prepare(X, Y, U, V, Z)
        keep = self.l1_reg, self.l2_reg
        self.l1_reg, self.l2_reg = l1_reg, l2_reg
        try:
            self._step(st)
        finally:
            self.l1_reg, self.l2_reg = keep
        if st is not X:
            self._write_back(st, U, V, Z)

    def _write_back(self, st, U, V, Z):
        be = st.be
        if self.update_U:
            U[...] = be.to_host(st.comm.all_gather_rows(st.U, st.n_total))
        if self.update_V:
            V[...] = be.to_host(st.V)
        if self.update_Z:
            Z[...] = be.to_host(st.Z)


def _draw_masks_numpy(n, d, l, ratio, update_U, update_Z, update_V):
    """Per-row sample sets from NumPy's global legacy RNG in the reference's call order
    (cmf_solvers.py:328-344 called from :414 [U], :494 [Z], :455-456 [V: rows of U, then columns of Y])."""
    s_d, s_n, s_l = int(d * ratio), int(n * ratio), int(l * ratio)
    out = {}
    perm = np.random.permutation
    if update_U:
        out["U"] = np.array([perm(np.arange(d))[:s_d] for _ in range(n)], dtype=np.int32).reshape(n, s_d)
    if update_Z:
        out["Z"] = np.array([perm(np.arange(d))[:s_d] for _ in range(l)], dtype=np.int32).reshape(l, s_d)
    if update_V:
        vx = np.empty((d, s_n), dtype=np.int32)
        vy = np.empty((d, s_l), dtype=np.int32)
        for j in range(d):
            vx[j] = perm(np.arange(n))[:s_n]
            vy[j] = perm(np.arange(l))[:s_l]
        out["Vx"], out["Vy"] = vx, vy
    return out


class NewtonSolver(_IterativeCMFSolver):
    """Row-wise Newton-Raphson solver (reference cmf_solvers.py:319-522): order U, Z, V."""

    NUMPY_SAMPLER_LIMIT = 5e7   # 'auto' uses the reference's NumPy RNG stream up to this many drawn indices

    def _masks(self, st):
        """Sample index sets for this iteration as device tensors (or None when sg_sample_ratio == 1)."""
        ratio = self.sg_sample_ratio
        if ratio >= 1.:
            return None
        n, d, l, _ = st.shapes
        be = st.be
        if self.masks_per_iter is not None:
            host = self.masks_per_iter[st.iteration - 1]
        else:
            mode = self.sampler
            if mode == "auto":
                draws = (n * d if self.update_U else 0) + (l * d if self.update_Z else 0) + \
                        (d * (n + l) if self.update_V else 0)
                mode = "numpy" if draws <= self.NUMPY_SAMPLER_LIMIT else "device"
                if mode == "device":
                    warnings.warn("sg_sample_ratio < 1 on a large problem: sampling on device (statistically "
                                  "equivalent to, but not the same stream as, the reference's np.random draws)")
            if mode == "device":
                return self._masks_device(st)
            self._sync_numpy_rng(st)
            host = _draw_masks_numpy(n, d, l, ratio, self.update_U, self.update_Z, self.update_V)
        dev = {}
        by_columns = st.Xcol is not None
        for key, val in host.items():
            val = np.asarray(val)
            if key == "U":
                val = val[st.r0:st.r1]
            elif key in ("Vx", "Vy") and by_columns:
                val = val[st.c0:st.c1]                   # this rank's rows of V; sampled rows of U stay GLOBAL indices
            elif key == "Vx" and st.comm.world > 1:
                val = localize_indices(val, st.r0, st.r1)
            dev[key] = be.to_device(val, np.int32)
        return dev

    def _sync_numpy_rng(self, st):
        """More than one rank drawing from NumPy's global RNG: every rank must see the same stream, or the replicated
        V / Z silently diverge.  An integer / array random_state already seeded every rank alike (constructor); otherwise
        rank 0 draws one seed from its stream and every rank re-seeds with it (once per solver)."""
        if st.comm.world == 1 or getattr(self, "_rng_synced", False):
            return
        self._rng_synced = True
        if self.random_state is not None and isinstance(self.random_state, INTEGER_TYPES + (np.ndarray, list)):
            return
        be = st.be
        seed = be.zeros(1, dtype=be.torch.int64)
        if st.comm.rank == 0:
            seed[0] = int(np.random.randint(0, 2 ** 31 - 1))
        st.comm.all_reduce_sum(seed)
        np.random.seed(int(be.to_host(seed)[0]))

    def _masks_device(self, st):
        n, d, l, _ = st.shapes
        be, ratio = st.be, self.sg_sample_ratio
        seed = 0 if self.random_state is None or not isinstance(self.random_state, INTEGER_TYPES) \
            else int(self.random_state)
        it = st.iteration
        s_d, s_n, s_l = int(d * ratio), int(n * ratio), int(l * ratio)
        out = {}
        # keyed by (seed, stream, GLOBAL row): a rank draws the sets of the rows of U it owns, and of the V update's
        # samples of U rows it keeps the ones inside its row block -- the sets do not depend on the shard count
        if self.update_U:
            out["U"] = be.sample_indices(st.r1 - st.r0, d, s_d, seed, 4 * it + 0, row0=st.r0)
        if self.update_Z:
            out["Z"] = be.sample_indices(l, d, s_d, seed, 4 * it + 1)
        if self.update_V and st.Xcol is not None:
            # column-sharded V phase: the sets of this rank's rows of V only, sampled rows of U as global indices
            out["Vx"] = be.sample_indices(st.c1 - st.c0, n, s_n, seed, 4 * it + 2, row0=st.c0)
            out["Vy"] = be.sample_indices(st.c1 - st.c0, l, s_l, seed, 4 * it + 3, row0=st.c0)
        elif self.update_V:
            out["Vx"] = be.sample_indices(d, n, s_n, seed, 4 * it + 2,
                                          window=(st.r0, st.r1) if st.comm.world > 1 else None)
            out["Vy"] = be.sample_indices(d, l, s_l, seed, 4 * it + 3)
        return out

    # ---- column-sharded V phase ("Newton, logit x-link / sg<1") ---------------------------------------
    def _v_phase_mode(self):
        import os
        mode = self.v_phase
        if mode == "auto":
            mode = os.environ.get("PYCMF_B200_V_PHASE", "auto")
        if mode not in ("rows", "columns", "auto"):
            raise ValueError("v_phase must be 'rows', 'columns' or 'auto', got %r" % (mode,))
        return mode

    def _wants_columns(self, world):
        """Per-row Hessians of the V update sum over ALL rows of U: row shards have to all-reduce d k^2 numbers per
        iteration (13 GB at C4) and every rank repeats all d clamped solves.  'columns' re-partitions for the V phase
        instead: every rank owns d / G rows of V and the matching column block of X over all rows (a second copy), U is
        all-gathered (n k), the new rows of V are all-gathered (d k), and the solves shrink by G.  'rows' keeps the chunked
        all-reduce of the partial Hessians (no second copy of X).  'auto' = 'columns' (PYCMF_B200_V_PHASE overrides)."""
        per_row = self.x_link == "logit" or self.sg_sample_ratio < 1.
        return self._v_phase_mode() != "rows" and world > 1 and self.update_V and per_row

    def _prepare_column_block(self, be, comm, X, Xd, d, r0):
        """This rank's column block of X over all rows (None when the V phase stays row-sharded).  Three sources:
        the whole host matrix (every rank uploads its block: a second PCIe transfer the size of the row shard), the whole
        matrix already in HBM (device initialisation: a view, nothing moves), or -- `sharded_input=True`, a rank holds its
        rows only -- the row shards resident on the ranks, re-partitioned by one all-to-all over NVLink."""
        if X is None or not self._wants_columns(comm.world):
            return None, None
        ranges = [row_range(d, g, comm.world) for g in range(comm.world)]
        c0, c1 = ranges[comm.rank]
        on_device = getattr(X, "is_sparse", None) is not None
        if self.sharded_input:
            return be.column_block(Xd, comm, r0, ranges), (c0, c1)
        if on_device:
            return be.col_slice(X, c0, c1), (c0, c1)
        if sp.issparse(X):
            block = sp.csc_matrix(X)[:, c0:c1].tocsr()
        else:
            block = np.ascontiguousarray(np.asarray(X)[:, c0:c1])
        return be.ingest(block), (c0, c1)

    def _step_v_columns(self, st, idx_x, idx_y, join_z):
        """_newton_update_V (:432-486) for this rank's rows [c0, c1) of V against all rows of U_new."""
        be, comm = st.be, st.comm
        alpha, l1, l2, pert = self.alpha, self.l1_reg, self.l2_reg, self.hessian_pertubation
        d, k = st.V.shape
        c0, c1 = st.c0, st.c1
        U_all = comm.all_gather_rows(st.U, st.n_total, st.row_counts)
        V_loc = st.V[c0:c1]
        Y_loc = be.row_slice(st.Y, c0, c1)
        step = be.v_chunk_rows(max(1, c1 - c0), k, True)
        for j0 in range(0, c1 - c0, step):
            j1 = min(c1 - c0, j0 + step)
            gx, Hx, pr = be.newton_v_xpart(V_loc, U_all, st.Xcol, j0, j1, self.x_link, alpha,
                                           idx=None if idx_x is None else idx_x[j0:j1])
            join_z()
            be.newton_v_finish(V_loc, st.Z, Y_loc, j0, j1, self.y_link, alpha, l1, l2, gx, Hx, pr,
                               self.V_non_negative, pert, idx=None if idx_y is None else idx_y[j0:j1])
        join_z()
        if d % comm.world == 0:
            comm.all_gather_into(st.V, V_loc.clone())
        else:
            st.V.copy_(comm.all_gather_rows(V_loc, d))

    def _step(self, st):
        be = st.be
        m = self._masks(st) or {}
        alpha, l1, l2, pert = self.alpha, self.l1_reg, self.l2_reg, self.hessian_pertubation
        # The U update reads (U, V, X) and the Z update (Z, V, Y): neither sees the other's result, so the order
        # U, Z of the reference is kept by running Z on the auxiliary stream next to U (joined before V).
        side = be.fork() if (self.update_U and self.update_Z and hasattr(be, "fork")) else be
        if self.update_Z:                                        # :515-517 -> _newton_update_Z :488-508
            side.newton_left(st.Z, st.V, st.Y, 1 - alpha, l1, l2, self.y_link, self.Z_non_negative, pert,
                             l2_in_logit_hessian=True, idx=m.get("Z"), trans=True)
        if self.update_U:                                        # :511-513 -> _newton_update_U :394-430
            be.newton_left(st.U, st.V, st.X, alpha, l1, l2, self.x_link, self.U_non_negative, pert,
                           l2_in_logit_hessian=False, idx=m.get("U"))
        # The X part of the V update (gradient / Hessian partials from U_new, V_old, X) does not read Z either: the join with
        # the Z branch is delayed until just before the first newton_v_finish -- the first kernel that needs Z_new and the
        # first that writes V, which the Z branch is still reading.  On several ranks the replicated Z update (0.10 ms on C2)
        # then hides behind the U update, the pass over X and the collectives instead of heading the critical path.
        pending = [side is not be]

        def join_z():
            if pending[0]:
                be.join()
                pending[0] = False

        if not self.update_V:
            join_z()
        if self.update_V:                                        # :519-522 -> _newton_update_V :432-486
            d, k = st.V.shape
            idx_x, idx_y = m.get("Vx"), m.get("Vy")
            per_row = be.newton_v_needs_per_row(self.x_link, idx_x is not None)
            world = st.comm.world
            if st.Xcol is not None:
                self._step_v_columns(st, idx_x, idx_y, join_z)
                return
            if world > 1 and not per_row and idx_y is None and d % world == 0:
                # Shared X-side Hessian: the gradient partial is reduce-scattered by rows of V, every rank finishes its
                # d / G rows (the per-row solves are the expensive, replicated part otherwise) and the new rows are
                # all-gathered: the two collectives together move what one all-reduce moves.
                gx, Hx, pr = be.newton_v_xpart(st.V, st.U, st.X, 0, d, self.x_link, alpha)
                st.comm.all_reduce_sum(Hx)
                gx_loc = st.comm.reduce_scatter_rows(gx)
                j0 = st.comm.rank * (d // world)
                j1 = j0 + d // world
                join_z()
                be.newton_v_finish(st.V, st.Z, st.Y, j0, j1, self.y_link, alpha, l1, l2, gx_loc, Hx, pr,
                                   self.V_non_negative, pert)
                st.comm.all_gather_into(st.V, st.V[j0:j1].clone())
                return
            step = be.v_chunk_rows(d, k, per_row)
            for j0 in range(0, d, step):
                j1 = min(d, j0 + step)
                gx, Hx, pr = be.newton_v_xpart(st.V, st.U, st.X, j0, j1, self.x_link, alpha,
                                               idx=None if idx_x is None else idx_x[j0:j1])
                st.comm.all_reduce_sum(gx)
                st.comm.all_reduce_sum(Hx)
                join_z()
                be.newton_v_finish(st.V, st.Z, st.Y, j0, j1, self.y_link, alpha, l1, l2, gx, Hx, pr,
                                   self.V_non_negative, pert, idx=None if idx_y is None else idx_y[j0:j1])
            join_z()

    def update_step(self, X, Y, U, V, Z, l1_reg, l2_reg, alpha):
        st = X if isinstance(X, FitState) else self.prepare(X, Y, U, V, Z)
        keep = self.l1_reg, self.l2_reg, self.alpha
        self.l1_reg, self.l2_reg, self.alpha = l1_reg, l2_reg, alpha
        try:
            st.iteration += 1
            self._step(st)
        finally:
            self.l1_reg, self.l2_reg, self.alpha = keep
        if st is not X:
            MUSolver._write_back(self, st, U, V, Z)

"""Device side of the fit loop: ingest of X / Y into HBM and the CUDA backend object.

`CudaBackend` is a thin, typed veneer over the C ABI (include/pycmf_b200.h): every method maps to one
entry point and passes torch-owned device pointers.  torch is used for device memory, streams and
(in sharding.py) torch.distributed only.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib

_NP2CODE = {np.dtype("float32"): _lib.F32, np.dtype("float64"): _lib.F64}
LINKS = {"linear": _lib.LINEAR, "logit": _lib.LOGIT}


def link_code(link):
    """Reference strings -> ABI codes; same error as cmf_solvers.py:33."""
    try:
        return LINKS[link]
    except KeyError:
        raise ValueError("Invalid link function {}".format(link))


class DenseMatrix:
    """Row-major dense matrix in HBM (rows x cols)."""
    is_sparse = False

    def __init__(self, tensor):
        self.t = tensor
        self.shape = tuple(tensor.shape)


class SparseMatrix:
    """CSR (row access) + CSC (column access) copies of one matrix in HBM, int32 indices, sorted.

    The CSC arrays are the CSR arrays of the transpose: the MU V update (X^T U, cmf_solvers.py:244)
    and the Newton V update (column j of X, :459) read them.
    """
    is_sparse = True

    def __init__(self, shape, rowptr, colidx, vals, colptr, rowidx, cvals):
        self.shape = tuple(shape)
        self.nnz = int(vals.shape[0])
        # a matrix (or a rank's shard) without nonzeros still hands non-NULL index / value pointers to the C ABI
        colidx, vals, rowidx, cvals = (_non_null(t) for t in (colidx, vals, rowidx, cvals))
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.colptr, self.rowidx, self.cvals = colptr, rowidx, cvals


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _non_null(t):
    """t, or one zero element of its kind when t is empty (never read: the row / column pointers say so)."""
    return t if t is None or t.numel() > 0 else t.new_zeros(1)


def column_block_from_row_shards(M, comm, r0, col_ranges):
    """Row shards -> column blocks on the device, for the column-sharded Newton V phase: every rank holds the
    rows [r0, r0 + M.shape[0]) of X and ends up with the columns `col_ranges[rank]` of X over ALL rows.  One all-to-all
    (NCCL over NVLink) of what already is in HBM; nothing crosses PCIe.  torch only (plumbing).

    Dense: the column slabs of the row shard are sent to their owners, who stack the received slabs by source rank.
    Sparse: in the CSC arrays of a row shard the nonzeros of the columns a rank owns are one contiguous segment, so the
    shard's (row index, value) arrays are the send buffer as they lie; the receiver merges the segments column by column
    with a stable sort on the column id (source ranks, and the rows inside each, stay ascending: sorted CSC).  Only the
    column access path of the block is built (that is all newton_v_xpart reads)."""
    import torch
    world, rank = comm.world, comm.rank
    c0, c1 = col_ranges[rank]
    widths = [b - a for a, b in col_ranges]
    n_loc = M.shape[0]
    if not M.is_sparse:
        dev = M.t.device
        rows_of = [v[0] for v in comm.all_gather_ints([n_loc], dev)]
        send = torch.cat([M.t[:, a:b].reshape(-1) for a, b in col_ranges])
        got = comm.all_to_all_chunks(send, [n_loc * w for w in widths], [r * (c1 - c0) for r in rows_of])
        del send
        block = got.view(sum(rows_of), c1 - c0)       # source blocks are row blocks in rank order
        return DenseMatrix(block)
    dev = M.colptr.device
    colptr = M.colptr.to(torch.int64)
    bounds = [(int(colptr[a]), int(colptr[b])) for a, b in col_ranges]
    seg = [hi - lo for lo, hi in bounds]
    info = comm.all_gather_ints([n_loc] + seg, dev)                       # per source: its rows, its segment sizes
    rows_of = [v[0] for v in info]
    recv_seg = [v[1 + rank] for v in info]
    counts = (colptr[1:] - colptr[:-1]).to(torch.int32)                     # nonzeros per column of this row shard
    got_counts = comm.all_to_all_chunks(counts, widths, [c1 - c0] * world).view(world, c1 - c0)
    nnz_loc = sum(seg)                                                      # (an empty shard holds one dummy element)
    got_rows = comm.all_to_all_chunks(M.rowidx[:nnz_loc] + int(r0), seg, recv_seg)    # global row numbers
    got_vals = comm.all_to_all_chunks(M.cvals[:nnz_loc], seg, recv_seg)
    col_of = torch.repeat_interleave(torch.arange(c1 - c0, device=dev, dtype=torch.int32).repeat(world),
                                     got_counts.reshape(-1).to(torch.int64))
    order = torch.sort(col_of, stable=True).indices
    out = SparseMatrix.__new__(SparseMatrix)
    out.shape = (sum(rows_of), c1 - c0)
    out.rowptr = out.colidx = out.vals = None                               # row access is not defined on a column block
    out.colptr = torch.zeros(c1 - c0 + 1, dtype=torch.int32, device=dev)
    out.colptr[1:] = torch.cumsum(got_counts.sum(0).to(torch.int64), 0).to(torch.int32)
    out.rowidx = got_rows[order].to(torch.int32).contiguous()
    out.cvals = got_vals[order].contiguous()
    out.nnz = int(out.cvals.shape[0])
    out.rowidx, out.cvals = _non_null(out.rowidx), _non_null(out.cvals)
    return out


class CudaBackend:
    """One context (= one rank, one GPU, one stream)."""

    def __init__(self, device=None, dtype="float32", options=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BackendError("pycmf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in _NP2CODE:
            raise ValueError("dtype must be float32 or float64, got %r" % (dtype,))
        self.code = _NP2CODE[self.np_dtype]
        self.tdtype = torch.float32 if self.code == _lib.F32 else torch.float64
        handle = C.c_void_p()
        _lib.check(self.lib.pycmf_create(self.device.index, C.byref(handle)))
        self.ctx = handle
        self._aux = None
        self._options = {}
        self.use_stream(torch.cuda.current_stream(self.device))
        for key, val in (options or {}).items():
            self.set_option(key, val)

    # ---- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_aux", None) is not None:
            self._aux.close()
            self._aux = None
        if getattr(self, "ctx", None) is not None:
            self.lib.pycmf_destroy(self.ctx)
            self.ctx = None

    # ---- a second context on its own stream: phases that do not depend on each other (the U and Z updates of one
    #      iteration: neither reads what the other writes) run side by side, in eager mode and under graph capture
    def fork(self):
        """The auxiliary backend, its stream waiting for everything enqueued so far on this one's."""
        if self._options.get("side_streams", 1.0) == 0.0:
            return self
        if self._aux is None:
            aux = CudaBackend.__new__(CudaBackend)
            aux.torch, aux.lib, aux.device = self.torch, self.lib, self.device
            aux.np_dtype, aux.code, aux.tdtype = self.np_dtype, self.code, self.tdtype
            handle = C.c_void_p()
            _lib.check(self.lib.pycmf_create(self.device.index, C.byref(handle)))
            aux.ctx, aux._aux, aux._options = handle, None, {}
            aux.use_stream(self.torch.cuda.Stream(device=self.device))
            for key, val in self._options.items():
                aux.set_option(key, val)
            self._aux = aux
        self._aux.stream.wait_stream(self.stream)
        return self._aux

    def join(self):
        """This backend's stream waits for the auxiliary one."""
        if self._aux is not None:
            self.stream.wait_stream(self._aux.stream)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_stream(self, stream):
        self.stream = stream
        _lib.check(self.lib.pycmf_set_stream(self.ctx, C.c_void_p(stream.cuda_stream)))

    HOST_OPTIONS = ("pad_pitch",)        # handled in this class, unknown to the C library

    def set_option(self, key, value):
        if key not in self.HOST_OPTIONS:
            _lib.check(self.lib.pycmf_set_option(self.ctx, key.encode(), float(value)))
        self._options[key] = float(value)
        if self._aux is not None:
            self._aux.set_option(key, value)

    def launch_count(self):
        n = int(self.lib.pycmf_launch_count(self.ctx))
        return n + (self._aux.launch_count() if self._aux is not None else 0)

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def profile(self, on=True):
        self._profiling = bool(on)
        _lib.check(self.lib.pycmf_profile_enable(self.ctx, int(bool(on))))
        if self._aux is not None:
            self._aux.profile(on)

    def profile_reset(self):
        _lib.check(self.lib.pycmf_profile_reset(self.ctx))
        if self._aux is not None:
            self._aux.profile_reset()

    def profile_query(self, family):
        """(total milliseconds, launches) of one kernel family since the last reset (both contexts)."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        _lib.check(self.lib.pycmf_profile_query(self.ctx, family.encode(), C.byref(ms), C.byref(cnt)))
        tot, n = ms.value, cnt.value
        if self._aux is not None:
            t2, n2 = self._aux.profile_query(family)
            tot, n = tot + t2, n + n2
        return tot, n

    def capture_step(self, fn):
        """Runs fn() once under CUDA-graph capture (it executes on replay, not now) and returns the graph.
        The library's launches follow the capture stream; event timers are off while capturing.

        torch.cuda.graph() is not used on purpose: its __enter__ empties the device and pinned-host allocator caches
        (cudaFree of every cached block, then fresh cudaMallocs), measured at 12 ms per capture on C2 -- a third of a
        50-iteration fit.  CUDAGraph.capture_begin / capture_end on a side stream do the same capture without it."""
        torch = self.torch
        prev_stream = self.stream
        g = torch.cuda.CUDAGraph()
        was_profiling = bool(getattr(self, "_profiling", False))
        self.profile(False)
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=self.device)
        cap = self._cap_stream
        cap.wait_stream(torch.cuda.current_stream(self.device))
        try:
            with torch.cuda.stream(cap):
                self.use_stream(cap)
                g.capture_begin()
                try:
                    fn()
                finally:
                    g.capture_end()
        finally:
            self.use_stream(prev_stream)
            if was_profiling:
                self.profile(True)
        torch.cuda.current_stream(self.device).wait_stream(cap)
        g.replay()          # the captured iteration has not run yet: run it now
        return g

    # ---- memory ------------------------------------------------------------------------------
    def empty(self, *shape, dtype=None):
        return self.torch.empty(*shape, dtype=dtype or self.tdtype, device=self.device)

    def zeros(self, *shape, dtype=None):
        return self.torch.zeros(*shape, dtype=dtype or self.tdtype, device=self.device)

    def to_device(self, a, dtype=None, raw=False):
        """Host ndarray -> contiguous device tensor of the compute dtype (or `dtype`).

        The bytes are copied as they are (asynchronously when the array lives in pinned memory) and
        the cast to the compute dtype happens on the GPU, so a float64 host matrix costs one PCIe
        transfer and no host-side conversion pass."""
        torch = self.torch
        want = np.dtype(self.np_dtype if dtype is None else dtype)
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64, np.int32, np.int64):
            a = a.astype(want)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
        d = t.to(self.device, non_blocking=t.is_pinned())
        if raw:
            return d
        tw = {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
              np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64}[want]
        return d if d.dtype == tw else d.to(tw)

    def dense(self, t):
        """DenseMatrix of the compute dtype over the values of the device tensor `t` (rows x cols, any float dtype).

        When a row is not a multiple of 128 bytes the matrix is stored with the row pitch padded up to one: the passes
        that read column blocks of X (X^T U, the Newton V gradient) fetch 512-byte row segments, and at an unaligned
        pitch every segment straddles one more 128-byte line (ncu, C2 with pitch 20000 B: 535 MB from HBM for a
        404 MB X).  The cast of a float64 upload goes straight into the padded buffer: no extra pass."""
        torch = self.torch
        rows, cols = t.shape
        es = torch.empty(0, dtype=self.tdtype).element_size()
        pad = self._options.get("pad_pitch", 1.0) != 0.0 and rows * cols >= (1 << 16) and (cols * es) % 128 != 0
        if not pad:
            return DenseMatrix(t.contiguous() if t.dtype == self.tdtype else t.to(self.tdtype))
        ld = (cols * es + 127) // 128 * 128 // es
        buf = torch.empty(rows, ld, dtype=self.tdtype, device=self.device)
        buf[:, cols:].zero_()
        buf[:, :cols].copy_(t)
        return DenseMatrix(buf[:, :cols])

    def dense_empty(self, rows, cols):
        """Uninitialised DenseMatrix (rows x cols) with the padded row pitch of dense(): filled in place by the caller
        (a 40 GB X is generated / uploaded block by block without a second full-size buffer)."""
        torch = self.torch
        es = torch.empty(0, dtype=self.tdtype).element_size()
        pad = self._options.get("pad_pitch", 1.0) != 0.0 and rows * cols >= (1 << 16) and (cols * es) % 128 != 0
        ld = (cols * es + 127) // 128 * 128 // es if pad else cols
        buf = torch.empty(rows, ld, dtype=self.tdtype, device=self.device)
        if ld != cols:
            buf[:, cols:].zero_()
        return DenseMatrix(buf[:, :cols])

    def to_host(self, t):
        return t.detach().to("cpu").numpy()

    def to_host_many(self, tensors):
        """Several device tensors -> host ndarrays through one pinned staging buffer and one synchronisation
        (three separate pageable copies of the factors cost 2 ms on C2, this 0.3 ms)."""
        torch = self.torch
        tensors = [t.detach().contiguous() for t in tensors]
        nbytes = sum(t.numel() * t.element_size() for t in tensors)
        if getattr(self, "_stage", None) is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        views, off = [], 0
        for t in tensors:
            n = t.numel() * t.element_size()
            v = self._stage[off:off + n].view(t.dtype).view(t.shape)
            v.copy_(t, non_blocking=True)
            views.append(v)
            off += (n + 15) & ~15
        torch.cuda.current_stream(self.device).synchronize()
        return [v.numpy() for v in views]          # views of the staging buffer: copy out before the next call

    def ingest(self, M):
        """Host matrix (ndarray / scipy sparse) -> DenseMatrix or SparseMatrix in HBM.
        Replaces check_array(..., accept_sparse=('csr','csc'), dtype=float) of cmf.py:386-388."""
        if isinstance(M, (DenseMatrix, SparseMatrix)):
            return M
        if sp.issparse(M):
            # no host-side passes over the nonzeros that are not needed: a canonical CSR (sorted, no duplicates -- what
            # TfidfVectorizer / check_array hand over) is uploaded as it is and cast on the GPU; at C3 scale every avoided
            # host copy of the 2.5e7-nonzero shard is ~100 ms of a 0.4 s fit
            csr = M if sp.isspmatrix_csr(M) else sp.csr_matrix(M)
            if not csr.has_canonical_format:
                csr = csr.copy() if csr is M else csr
                csr.sum_duplicates()
                csr.sort_indices()
            if csr.nnz >= 2 ** 31 or max(csr.shape) >= 2 ** 31:
                raise ValueError("sparse matrix too large for int32 indices")
            # only the CSR arrays cross PCIe; the CSC copy (the CSR of X^T: X^T U and the column access of the Newton V
            # update) is built on the device by a stable sort on the column index -- rows stay ascending inside a column,
            # i.e. exactly scipy's sorted csr.T.tocsr(), which on the host takes seconds at C3 scale (2e8 nonzeros)
            rowptr = self.to_device(csr.indptr, np.int32)
            colidx = self.to_device(csr.indices, np.int32)
            vals = self.to_device(csr.data)
            colptr, rowidx, cvals = self._csc_from_csr(rowptr, colidx, vals, csr.shape[0], csr.shape[1])
            return SparseMatrix(csr.shape, rowptr, colidx, vals, colptr, rowidx, cvals)
        M = np.asarray(M)
        if M.ndim != 2:
            raise ValueError("Expected 2D array, got %dD array instead" % M.ndim)
        if M.nbytes > self.INGEST_CHUNK_BYTES and M.dtype in (np.float32, np.float64) and M.flags.c_contiguous:
            return self._ingest_dense_chunked(M)
        return self.dense(self.to_device(M, raw=True))

    INGEST_CHUNK_BYTES = 1 << 28

    def _ingest_dense_chunked(self, M):
        """Large dense host matrix -> padded-pitch DenseMatrix through two staging buffers of INGEST_CHUNK_BYTES: the raw
        bytes of a row block cross PCIe while the previous block is cast / pitched into place, and HBM never holds a
        second full-size copy (C5: 40 GB of X next to 40 GB of staging would not leave room for anything else)."""
        torch = self.torch
        rows, cols = M.shape
        out = self.dense_empty(rows, cols)
        step = max(1, self.INGEST_CHUNK_BYTES // max(1, M.strides[0]))
        host = torch.from_numpy(M if M.flags.writeable else M.copy())
        pinned = host.is_pinned()
        stages = [torch.empty(step, cols, dtype=host.dtype, device=self.device) for _ in range(2)]
        copy_stream = getattr(self, "_ingest_stream", None)
        if copy_stream is None:
            copy_stream = self._ingest_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        copy_stream.wait_stream(main)
        done = [None, None]          # event: the cast out of stage i has finished
        for i, r0 in enumerate(range(0, rows, step)):
            r1 = min(rows, r0 + step)
            stg = stages[i & 1][:r1 - r0]
            with torch.cuda.stream(copy_stream):
                if done[i & 1] is not None:
                    copy_stream.wait_event(done[i & 1])
                stg.copy_(host[r0:r1], non_blocking=pinned)
                arrived = torch.cuda.Event()
                arrived.record(copy_stream)
            main.wait_event(arrived)
            out.t[r0:r1].copy_(stg)
            done[i & 1] = torch.cuda.Event()
            done[i & 1].record(main)
        for s in stages:
            s.record_stream(copy_stream)
        return out

    def row_slice(self, M, r0, r1):
        """Rows [r0, r1) of an ingested matrix (view for dense, re-based copy for sparse)."""
        if not M.is_sparse:
            return DenseMatrix(M.t[r0:r1])
        torch = self.torch
        rp = M.rowptr[r0:r1 + 1].clone()
        lo, hi = int(rp[0]), int(rp[-1])
        rp -= lo
        colidx, vals = M.colidx[lo:hi].clone(), M.vals[lo:hi].clone()
        colptr, rowidx, cvals = self._csc_from_csr(rp, colidx, vals, r1 - r0, M.shape[1])
        return SparseMatrix((r1 - r0, M.shape[1]), rp, colidx, vals, colptr, rowidx, cvals)

    def column_block(self, M, comm, r0, col_ranges):
        """This rank's column block of X over all rows, from the row shards resident on the ranks (all-to-all)."""
        blk = column_block_from_row_shards(M, comm, r0, col_ranges)
        return blk if blk.is_sparse else self.dense(blk.t)        # same (padded-pitch) layout as an ingested matrix

    def col_slice(self, M, c0, c1):
        """Columns [c0, c1) of an ingested matrix as the operand of a V-side pass (X^T U over a column slab): a strided
        view for dense X, the CSC arrays with the column pointer sliced (offsets stay absolute) for sparse X."""
        if not M.is_sparse:
            return DenseMatrix(M.t[:, c0:c1])
        S = SparseMatrix.__new__(SparseMatrix)
        S.shape = (M.shape[0], c1 - c0)
        S.rowptr = S.colidx = S.vals = None                      # row access is not defined on a column slab
        S.colptr, S.rowidx, S.cvals = M.colptr[c0:c1 + 1], M.rowidx, M.cvals
        S.nnz = None
        return S

    def _csc_from_csr(self, rowptr, colidx, vals, n_rows, n_cols):
        """(colptr, rowidx, cvals) of the matrix whose CSR arrays are given: stable sort of the nonzeros on the column."""
        torch = self.torch
        counts_r = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
        rows_of = torch.repeat_interleave(torch.arange(n_rows, device=self.device, dtype=torch.int32), counts_r)
        order = torch.sort(colidx, stable=True).indices
        counts = torch.bincount(colidx, minlength=n_cols)
        colptr = torch.zeros(n_cols + 1, dtype=torch.int32, device=self.device)
        colptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        return colptr, rows_of[order].contiguous(), vals[order].contiguous()

    # ---- target argument packing ---------------------------------------------------------------
    def _target_args(self, T, trans=False, rows=None):
        """(dense ptr, ld, trans, rowptr, colidx, vals) for a left-factor target."""
        if T.is_sparse:
            if trans:
                return 0, 0, 0, _ptr(T.colptr), _ptr(T.rowidx), _ptr(T.cvals)
            return 0, 0, 0, _ptr(T.rowptr), _ptr(T.colidx), _ptr(T.vals)
        t = T.t
        assert t.stride(1) == 1
        return t.data_ptr(), t.stride(0), int(bool(trans)), 0, 0, 0

    # ---- objective -----------------------------------------------------------------------------
    def sqerr(self, A, B, T, link, trans=False):
        """sum (T - f(A B^T))^2 as a 0-d float64 device tensor (cmf_solvers.py:36-42)."""
        out = self.zeros(1, dtype=self.torch.float64)
        rows, k = A.shape
        m = B.shape[0]
        tp, ld, tr, rp, ci, vl = self._target_args(T, trans)
        _lib.check(self.lib.pycmf_sqerr(self.ctx, self.code, rows, m, k, _ptr(A), _ptr(B), tp, ld, tr, rp, ci, vl,
                                        link_code(link), _ptr(out)))
        return out

    def resid_pass(self, A, B, T, link, want_left=True, want_right=True, want_sq=False, trans_t=False):
        """(R B, R^T A, sum R^2) with R = f(A B^T) - T for a dense target T (DenseMatrix; stored transposed,
        rows(B) x rows(A), when `trans_t`)."""
        rows, k = A.shape
        m = B.shape[0]
        outL = self.empty(rows, k) if want_left else None
        outR = self.empty(m, k) if want_right else None
        sq = self.zeros(1, dtype=self.torch.float64) if want_sq else None
        t = T.t
        _lib.check(self.lib.pycmf_resid_pass(self.ctx, self.code, rows, m, k, _ptr(A), _ptr(B), t.data_ptr(), t.stride(0),
                                             int(bool(trans_t)), link_code(link), _ptr(outL), _ptr(outR), _ptr(sq)))
        return outL, outR, sq

    # ---- MU ------------------------------------------------------------------------------------
    def mu_v_partial(self, X, U, out=None):
        n, k = U.shape
        d = X.shape[1]
        if out is None:
            out = self.empty(d + k, k)
        if X.is_sparse:
            args = (0, 0, _ptr(X.colptr), _ptr(X.rowidx), _ptr(X.cvals))
        else:
            args = (X.t.data_ptr(), X.t.stride(0), 0, 0, 0)
        _lib.check(self.lib.pycmf_mu_v_partial(self.ctx, self.code, n, d, k, *args, _ptr(U), _ptr(out)))
        return out

    def mu_v_apply(self, V, buf, Y, Z, l1, l2):
        d, k = V.shape
        l = Z.shape[0]
        _lib.check(self.lib.pycmf_mu_v_apply(self.ctx, self.code, d, l, k, _ptr(V), _ptr(buf), Y.t.data_ptr(),
                                             Y.t.stride(0), _ptr(Z), float(l1), float(l2)))

    def mu_left(self, F, B, T, l1, l2, trans=False):
        rows, k = F.shape
        m = B.shape[0]
        tp, ld, tr, rp, ci, vl = self._target_args(T, trans)
        _lib.check(self.lib.pycmf_mu_left(self.ctx, self.code, rows, m, k, _ptr(F), _ptr(B), tp, ld, tr, rp, ci, vl,
                                          float(l1), float(l2)))

    # ---- Newton --------------------------------------------------------------------------------
    def _idx_args(self, idx):
        if idx is None:
            return 0, 0
        assert idx.dtype == self.torch.int32 and idx.is_contiguous()
        n_sample = int(idx.shape[1])
        # a non-NULL pointer is required even for an empty sample set (n_sample == 0 kills the data term)
        return (idx.data_ptr() if n_sample > 0 else self._dummy().data_ptr()), n_sample

    def _dummy(self):
        if not hasattr(self, "_dummy_t"):
            self._dummy_t = self.zeros(4, dtype=self.torch.int32)
        return self._dummy_t

    def newton_left(self, F, B, T, weight, l1, l2, link, non_negative, pert, l2_in_logit_hessian,
                    idx=None, trans=False):
        rows, k = F.shape
        m = B.shape[0]
        tp, ld, tr, rp, ci, vl = self._target_args(T, trans)
        ip, ns = self._idx_args(idx)
        _lib.check(self.lib.pycmf_newton_left(self.ctx, self.code, rows, m, k, _ptr(F), _ptr(B), tp, ld, tr, rp, ci, vl,
                                              float(weight), float(l1), float(l2), link_code(link),
                                              int(bool(non_negative)), float(pert), int(bool(l2_in_logit_hessian)),
                                              ip, ns))

    def newton_v_needs_per_row(self, x_link, sampled):
        return bool(sampled) or x_link == "logit"

    def newton_v_xpart(self, V, U, X, j0, j1, x_link, alpha, idx=None):
        """X part of the V update for V rows [j0, j1): returns (gx, Hx, per_row)."""
        n, k = U.shape
        rows = j1 - j0
        per_row = self.newton_v_needs_per_row(x_link, idx is not None)
        gx = self.empty(rows, k)
        # per-row Hessians in the compute dtype; the shared one (alpha U^T U) always in float64
        Hx = self.empty(rows, k, k) if per_row else self.empty(1, k, k, dtype=self.torch.float64)
        if X.is_sparse:
            xargs = (0, 0, X.colptr[j0:].data_ptr(), _ptr(X.rowidx), _ptr(X.cvals))
        else:
            xargs = (X.t[:, j0:].data_ptr(), X.t.stride(0), 0, 0, 0)
        ip, ns = self._idx_args(idx)
        flag = C.c_int(0)
        _lib.check(self.lib.pycmf_newton_v_xpart(self.ctx, self.code, rows, n, k, V[j0:j1].data_ptr(), _ptr(U), *xargs,
                                                 link_code(x_link), float(alpha), ip, ns, _ptr(gx), _ptr(Hx),
                                                 C.byref(flag)))
        assert bool(flag.value) == per_row
        return gx, Hx, per_row

    def newton_v_finish(self, V, Z, Y, j0, j1, y_link, alpha, l1, l2, gx, Hx, per_row, non_negative, pert, idx=None):
        k = V.shape[1]
        l = Z.shape[0]
        ip, ns = self._idx_args(idx)
        _lib.check(self.lib.pycmf_newton_v_finish(self.ctx, self.code, j1 - j0, l, k, V[j0:j1].data_ptr(), _ptr(Z),
                                                  Y.t[j0:j1].data_ptr(), Y.t.stride(0), link_code(y_link), float(alpha),
                                                  float(l1), float(l2), ip, ns, _ptr(gx), _ptr(Hx), int(per_row),
                                                  int(bool(non_negative)), float(pert)))

    def v_chunk_rows(self, d, k, per_row, budget_bytes=1 << 30):
        if not per_row:
            return d
        return max(1, min(d, budget_bytes // (k * k * self.np_dtype.itemsize)))

    # ---- utilities -----------------------------------------------------------------------------
    def safe_solve(self, H, g, pert):
        """x = S(H) g for float64 tensors H (batch x k x k or k x k shared) and g (batch x k)."""
        batch, k = g.shape
        x = self.empty(batch, k, dtype=self.torch.float64)
        stride = 0 if H.dim() == 2 else k * k
        _lib.check(self.lib.pycmf_safe_solve(self.ctx, batch, k, _ptr(H), stride, _ptr(g), _ptr(x), float(pert)))
        return x

    def sample_indices(self, rows, N, n_sample, seed, stream_id, row0=0, window=None):
        """idx (rows x n_sample): the sample sets of global rows [row0, row0 + rows); `window` = (lo, hi) keeps only the
        indices a rank holds (re-based to lo, the others -1)."""
        idx = self.empty(rows, n_sample, dtype=self.torch.int32)
        lo, hi = window if window is not None else (0, 0)
        _lib.check(self.lib.pycmf_sample_indices_sharded(self.ctx, rows, int(row0), N, n_sample,
                                                         int(seed) & (2 ** 64 - 1), int(stream_id), int(lo), int(hi),
                                                         _ptr(idx)))
        return idx

    def topk_per_column(self, F, topn):
        """(columns x topn) int32: row indices of the topn largest entries of every column of F (rows x k), ascending
        weight == np.argsort(F[:, c], kind="stable")[-topn:] (reference analysis.py:6)."""
        rows, k = F.shape
        topn = int(min(topn, rows))
        out = self.empty(k, topn, dtype=self.torch.int32)
        assert F.stride(1) == 1
        _lib.check(self.lib.pycmf_topk_columns(self.ctx, self.code, rows, k, _ptr(F), F.stride(0), topn, _ptr(out)))
        return out

    def gemm(self, A, B, trans_a=False, alpha=1.0, beta=0.0, out=None):
        m = A.shape[1] if trans_a else A.shape[0]
        p = A.shape[0] if trans_a else A.shape[1]
        q = B.shape[1]
        if out is None:
            out = self.zeros(m, q)
        _lib.check(self.lib.pycmf_gemm(self.ctx, self.code, int(trans_a), m, q, p, _ptr(A), A.stride(0), _ptr(B),
                                       B.stride(0), _ptr(out), out.stride(0), float(alpha), float(beta)))
        return out

    def spmm(self, S, B, transposed=False, alpha=1.0, beta=0.0, out=None):
        rows = S.shape[1] if transposed else S.shape[0]
        k = B.shape[1]
        if out is None:
            out = self.zeros(rows, k)
        rp, ci, vl = (S.colptr, S.rowidx, S.cvals) if transposed else (S.rowptr, S.colidx, S.vals)
        _lib.check(self.lib.pycmf_spmm(self.ctx, self.code, rows, S.shape[0] if transposed else S.shape[1],
                                       _ptr(rp), _ptr(ci), _ptr(vl), _ptr(B), B.stride(0), k, _ptr(out),
                                       out.stride(0), float(alpha), float(beta)))
        return out

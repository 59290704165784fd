/*
 * pycmf_b200.h -- C ABI of libpycmf_b200.so, the sm_100a backend for PyCMF's fit loop.
 *
 * This library takes the place of the reference's only native module, the Cython extension
 * `pycmf.cmf_newton_solver` (reference setup.py:6-8; exports `_newton_update_left`
 * cmf_newton_solver.pyx:241-247 and `_newton_update_V` .pyx:296-303), and additionally moves the
 * NumPy bodies of `MUSolver.update_step` (cmf_solvers.py:248-263), `NewtonSolver.update_step`
 * (cmf_solvers.py:510-522) and `compute_error` (cmf_solvers.py:128-130) onto the GPU.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors on the Python side);
 *     the library only allocates scratch inside its context.  X / Y are never modified, factors
 *     are updated in place (reference ownership rule, cmf_solvers.py:124-126, :255-263).
 *   - matrices are row-major and dense unless a CSR triple (rowptr, colidx, vals) is given;
 *     CSR indices are int32, sorted within a row.  A NULL dense pointer selects the CSR operand.
 *   - dtype: PYCMF_F32 or PYCMF_F64 for every matrix of the call; scalar arguments are double.
 *   - link: PYCMF_LINEAR | PYCMF_LOGIT  (reference strings "linear"/"logit", cmf.py:395-399).
 *   - all work is enqueued on the context's stream (pycmf_set_stream); calls are asynchronous.
 *   - return value 0 = ok; otherwise pycmf_last_error() describes the failure (the Python host
 *     raises ValueError / RuntimeError, mirroring the reference's exceptions).
 *   - one context per host thread / rank; a context is not thread-safe (the reference is
 *     single-threaded under the GIL).
 */
#ifndef PYCMF_B200_H
#define PYCMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYCMF_ABI_VERSION 1

#define PYCMF_F32 0
#define PYCMF_F64 1

#define PYCMF_LINEAR 0
#define PYCMF_LOGIT 1

typedef struct pycmf_ctx pycmf_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
int pycmf_abi_version(void);
const char* pycmf_last_error(void);
int pycmf_create(int device, pycmf_ctx** out);
int pycmf_destroy(pycmf_ctx* ctx);
/* stream is a cudaStream_t (0 = legacy default stream) */
int pycmf_set_stream(pycmf_ctx* ctx, void* stream);
/* options: "chol_fastpath" (0/1, default 1), "dense_path" (0 = generic FMA kernels, 1 = tcgen05
 * 3xTF32 for n_components 32 / 64 / 128 / 192 / 256, 2 = tcgen05 1xTF32 (k = 32 only); default 1),
 * "spmm_path" (0 = generic CSR kernel, 1 = vector / sub-warp-grouped kernels for k = 32 / 64 / 128 / 256,
 * default 1), "solve_path" (clamped solve with the clamp active, k > 32: 0 = one-sided Jacobi, 1 = tridiagonalisation +
 * multi-section + inverse iteration, 2 = the same with the Householder steps on a register-resident matrix for
 * k = 64 / 128; default 2), "hess_mma" (0 = per-row weighted Grams on the FMA pipes, 1 = mma.sync 3xTF32 for k = 64 / 128;
 * default 1), "mu_fused" (0 = separate denominator GEMM + elementwise MU step, default 1), "max_scratch_mb",
 * "side_streams" (0/1), and for tests and tuning: "tc_max_splits", "tc_ctas", "tc_chain", "tc_x_promotion",
 * "tc_prefetch", "tc_trace", "finish_minblocks", "solve_threads", "spmm_blocks_per_sm", "spmm_unroll", "spmm_lean" */
int pycmf_set_option(pycmf_ctx* ctx, const char* key, double value);
/* number of kernels this context launched since creation (bench.py's gpu_launches) */
int64_t pycmf_launch_count(pycmf_ctx* ctx);

/* per-kernel-family device timers (bench.py's live roofline): enable, run, then query by family name
 * ("resid_left", "resid_right", "gemm", "spmm", "row_grad_hess", "safe_solve", "tc_resid_left",
 * "tc_resid_right", "tc_xv", "tc_xtu", "tc_ytv", "tc_factor", ...).
 * query synchronises the stream; total_ms / count cover everything since the last reset. */
int pycmf_profile_enable(pycmf_ctx* ctx, int on);
int pycmf_profile_query(pycmf_ctx* ctx, const char* family, double* total_ms, int64_t* count);
int pycmf_profile_reset(pycmf_ctx* ctx);

/* diagnostics: with option "tc_trace" = 1 every tcgen05 pass over X records clock64 stamps of its pipeline
 * events for the first 96 tiles of CTA 0 ([event][tile] int64; events of the fused-residual pass: see
 * scripts/tc_trace.py, of the MU pass: scripts/tc_mu_trace.py); this copies the last trace to the host. */
int pycmf_debug_tc_trace(pycmf_ctx* ctx, int64_t* host, int64_t max_words);

/* ---- primitives (used by the phases below; exported for tests and composition) ---------- */
/* C (m x q) = alpha * op(A) * B + beta * C;  op(A) = A (m x p) or A^T with A stored (p x m).
 * Replaces np.dot / safe_sparse_dot on dense operands (cmf_solvers.py:232-245). */
int pycmf_gemm(pycmf_ctx* ctx, int dtype, int trans_a, int64_t m, int64_t q, int64_t p,
               const void* A, int64_t lda, const void* B, int64_t ldb,
               void* C, int64_t ldc, double alpha, double beta);
/* C (rows x k) = alpha * S * B + beta * C with S (rows x cols) in CSR.
 * Replaces safe_sparse_dot(X, V) / safe_sparse_dot(X.T, U) (cmf_solvers.py:232, :244; the
 * transposed product is the same call on the CSC arrays of X, i.e. the CSR of X^T). */
int pycmf_spmm(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t cols,
               const int32_t* rowptr, const int32_t* colidx, const void* vals,
               const void* B, int64_t ldb, int64_t k, void* C, int64_t ldc,
               double alpha, double beta);

/* Fused residual pass: R = f(A B^T) - T (never materialised);  outL = R B (rows x k), outR = R^T A (m x k),
 * *sq += sum R^2 (device double).  Any of outL / outR / sq may be NULL.  This is the reference's
 * `res = inverse(np.dot(U, V.T), link) - X; np.dot(res, V); np.dot(res.T, U)` (cmf_solvers.py:399-400, :436-440). */
int pycmf_resid_pass(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k, const void* A, const void* B,
                     const void* T, int64_t ldt, int trans_t, int link, void* outL, void* outR, double* sq);

/* ---- objective (cmf_solvers.py:36-42; sklearn _beta_divergence beta=2) -------------------- */
/* *out_sq (device double) = sum_ij (T_ij - f(a_i . b_j))^2 over ALL entries, for
 *   T dense (rows x m, ld ldt; if trans_t the array is stored m x rows and read transposed)
 *   or T in CSR (T == NULL).  A is rows x k, B is m x k.
 * Sparse + linear uses ||T||^2 + tr((A^T A)(B^T B)) - 2 sum_nz t_ij a_i.b_j; sparse + logit a
 * dense pass plus a nonzero correction.  The caller takes sqrt (and sums shards first). */
int pycmf_sqerr(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k,
                const void* A, const void* B,
                const void* T, int64_t ldt, int trans_t,
                const int32_t* rowptr, const int32_t* colidx, const void* vals,
                int link, double* out_sq);

/* ---- multiplicative-update solver (cmf_solvers.py:212-263) ------------------------------- */
/* Shard-local part of the V update: out is a ((d + k) x k) buffer, rows [0,d) = X^T U
 * (cmf_solvers.py:244 first term), rows [d, d+k) = U^T U (:245).  X dense (n x d, ldx) or the
 * CSC arrays of X (= CSR of X^T, d rows).  With row-sharded X/U the caller all-reduces `out`. */
int pycmf_mu_v_partial(pycmf_ctx* ctx, int dtype, int64_t n, int64_t d, int64_t k,
                       const void* X, int64_t ldx,
                       const int32_t* csc_colptr, const int32_t* csc_rowidx, const void* csc_vals,
                       const void* U, void* out);
/* V *= (XtU + Y Z) / (V (UtU + Z^T Z) + l1 + l2 V), zero denominators -> float32 eps
 * (cmf_solvers.py:212-228, :242-246, :255).  xtu_utu is the (all-reduced) buffer from above. */
int pycmf_mu_v_apply(pycmf_ctx* ctx, int dtype, int64_t d, int64_t l, int64_t k,
                     void* V, const void* xtu_utu, const void* Y, int64_t ldy, const void* Z,
                     double l1_reg, double l2_reg);
/* F *= (T B) / (F (B^T B) + l1 + l2 F) for a left factor F (rows x k) against B (m x k):
 * U update with T = X (cmf_solvers.py:230-234, :259) and Z update with T = Y^T (trans_t = 1 on
 * the d x l array Y; :236-240, :263).  (F B^T) B is evaluated as F (B^T B). */
int pycmf_mu_left(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k,
                  void* F, const void* B,
                  const void* T, int64_t ldt, int trans_t,
                  const int32_t* rowptr, const int32_t* colidx, const void* vals,
                  double l1_reg, double l2_reg);

/* ---- Newton solver (cmf_solvers.py:321-522; replaces cmf_newton_solver.pyx:241-362) -------- */
/* Row-wise Newton update of a left factor F (rows x k) against B (m x k) and target T (rows x m):
 *   g_i = weight (f(f_i B_s^T) - T[i,s]) B_s + l1 sign(f_i) + l2 f_i
 *   H_i = weight B_s^T [diag f'(f_i B_s^T)] B_s  (+ l2 I unless link == logit && !l2_in_logit_hessian)
 *   f_i <- f_i - g_i S(H_i),  S = eigenvalue-clamped inverse (_safe_invert, cmf_solvers.py:346-356)
 *   negatives -> 0 if non_negative (:321-326).
 * U: T = X, weight = alpha, l2_in_logit_hessian = 0 (:394-430).
 * Z: T = Y^T (trans_t = 1), weight = 1 - alpha, l2_in_logit_hessian = 1 (:488-508).
 * sample_idx (rows x n_sample int32, NULL = use every column) are the per-row sample sets of
 * _stochastic_sample (:328-344); with n_sample == 0 and sample_idx != NULL the data terms vanish. */
int pycmf_newton_left(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k,
                      void* F, const void* B,
                      const void* T, int64_t ldt, int trans_t,
                      const int32_t* rowptr, const int32_t* colidx, const void* vals,
                      double weight, double l1_reg, double l2_reg, int link,
                      int non_negative, double hessian_pertubation, int l2_in_logit_hessian,
                      const int32_t* sample_idx, int64_t n_sample);
/* Shard-local X part of the V update for V rows [0, d_rows) of the given slice (:432-486):
 *   gx_j = alpha sum_{i in s} (f1(u_i . v_j) - X[i, j]) u_i                     -> gx (d_rows x k)
 *   Hx   = alpha U^T U (k x k, shared, ALWAYS FLOAT64 whatever `dtype`: rounded to float32 it costs 6e-8 x cond(H)
 *          on the Newton step) if x_link is linear and no sampling, else
 *   Hx_j = alpha sum_{i in s} f1'(u_i . v_j) u_i u_i^T                          -> (d_rows x k x k, compute dtype)
 * X is dense (n x d_total, ldx; the slice's first column is X + col0) or CSC arrays already
 * offset to the slice.  *hx_per_row tells the caller which Hx layout was written.
 * With row-sharded X/U the caller all-reduces gx and Hx before pycmf_newton_v_finish. */
int pycmf_newton_v_xpart(pycmf_ctx* ctx, int dtype, int64_t d_rows, int64_t n, int64_t k,
                         const void* V, const void* U,
                         const void* Xcols, int64_t ldx,
                         const int32_t* csc_colptr, const int32_t* csc_rowidx, const void* csc_vals,
                         int x_link, double alpha,
                         const int32_t* sample_idx_x, int64_t n_sample_x,
                         void* gx, void* Hx, int* hx_per_row);
/* Adds the Y part, regularisation and solves: V rows updated in place (:432-486).
 *   g_j = gx_j + (1-alpha)(f2(v_j Z_t^T) - Y[j,t]) Z_t + l1 sign(v_j) + l2 v_j
 *   H_j = Hx(_j) + (1-alpha) Z_t^T [D] Z_t + l2 I */
int pycmf_newton_v_finish(pycmf_ctx* ctx, int dtype, int64_t d_rows, int64_t l, int64_t k,
                          void* V, const void* Z, const void* Yrows, int64_t ldy,
                          int y_link, double alpha, double l1_reg, double l2_reg,
                          const int32_t* sample_idx_y, int64_t n_sample_y,
                          const void* gx, const void* Hx, int hx_per_row,
                          int non_negative, double hessian_pertubation);

/* ---- utilities -------------------------------------------------------------------------- */
/* x = S(H) g for a batch of symmetric k x k matrices (float64): the eigenvalue-clamped inverse of
 * _safe_invert (cmf_solvers.py:346-356) applied to a vector.  h_stride = 0 shares one matrix. */
int pycmf_safe_solve(pycmf_ctx* ctx, int64_t batch, int64_t k, const double* H, int64_t h_stride,
                     const double* g, double* x, double hessian_pertubation);
/* On-device sampler (K10): idx (rows x n_sample) <- n_sample distinct indices in [0, N) per row,
 * keyed by (seed, stream_id, row) through a cycle-walking Feistel permutation. */
int pycmf_sample_indices(pycmf_ctx* ctx, int64_t rows, int64_t N, int64_t n_sample,
                         uint64_t seed, uint64_t stream_id, int32_t* idx);
/* Topic-term extraction (reference analysis.py:1-16, `topic.argsort()[-10:]` per column of the term-topic matrix):
 * out (k x topn int32) <- for every column c of F (rows x k, row pitch ld) the row indices of its topn largest entries
 * in ascending weight order, ties by ascending index; -1 where the column has fewer than topn rows. */
int pycmf_topk_columns(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t k, const void* F, int64_t ld,
                       int64_t topn, int32_t* out);
/* The same sampler for a row-sharded fit: this rank draws the sets of global rows [row0, row0 + rows) (the key uses the
 * global row, so every shard count sees the same sets), and, when hi > lo, keeps only the indices inside the window
 * [lo, hi) re-based to lo -- the others become -1 and are skipped by the kernels (the rows of U a rank does not hold,
 * cmf_solvers.py:455). */
int pycmf_sample_indices_sharded(pycmf_ctx* ctx, int64_t rows, int64_t row0, int64_t N, int64_t n_sample,
                                 uint64_t seed, uint64_t stream_id, int64_t lo, int64_t hi, int32_t* idx);

#ifdef __cplusplus
}
#endif
#endif /* PYCMF_B200_H */

// extern "C" surface of libpycmf_b200.so (see include/pycmf_b200.h) and the phase compositions.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace pycmf {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

void* scratch(pycmf_ctx* ctx, int slot, size_t bytes) {
    Scratch& s = ctx->arena[slot];
    if (bytes == 0) bytes = 256;
    if (s.bytes < bytes) {
        // growing an arena frees memory: never under stream capture (it would invalidate the capture, and a graph captured
        // earlier has the old address baked in -- callers run two eager iterations first so that arenas have their size)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        PYCMF_CUDA(cudaStreamIsCapturing(ctx->stream, &cap));
        PYCMF_CHECK(cap == cudaStreamCaptureStatusNone,
                    "scratch arena would have to grow during CUDA-graph capture: run the same step eagerly first");
        PYCMF_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s.ptr) ctx->retired.push_back(s.ptr);      // not freed: a graph captured earlier may replay with this address
        s.ptr = nullptr;
        s.bytes = 0;
        size_t want = std::max(bytes, size_t(1) << 20);
        want = (want + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
        PYCMF_CUDA(cudaMalloc(&s.ptr, want));
        s.bytes = want;
    }
    return s.ptr;
}

pycmf_ctx* fork_side(pycmf_ctx* ctx) {
    if (ctx->root != nullptr || !ctx->side_streams) return ctx;       // no nesting; option off: serial on the main stream
    if (ctx->side == nullptr) {
        pycmf_ctx* s = new pycmf_ctx();
        s->device = ctx->device;
        s->num_sms = ctx->num_sms;
        s->max_smem_optin = ctx->max_smem_optin;
        s->root = ctx;
        PYCMF_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        PYCMF_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        PYCMF_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        ctx->side = s;
    }
    pycmf_ctx* s = ctx->side;
    s->chol_fastpath = ctx->chol_fastpath;
    s->dense_path = ctx->dense_path;
    s->spmm_path = ctx->spmm_path;
    s->spmm_blocks_per_sm = ctx->spmm_blocks_per_sm;
    s->spmm_unroll = ctx->spmm_unroll;
    s->spmm_lean = ctx->spmm_lean;
    s->mu_fused = ctx->mu_fused;
    s->solve_path = ctx->solve_path;
    s->hess_mma = ctx->hess_mma;
    s->solve_threads = ctx->solve_threads;
    s->tc_max_splits = ctx->tc_max_splits;
    s->tc_ctas = ctx->tc_ctas;
    s->tc_chain = ctx->tc_chain;
    s->tc_x_promotion = ctx->tc_x_promotion;
    s->tc_prefetch = ctx->tc_prefetch;
    s->max_scratch = ctx->max_scratch;
    s->finish_minblocks = ctx->finish_minblocks;
    PYCMF_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    PYCMF_CUDA(cudaStreamWaitEvent(s->stream, ctx->ev_fork, 0));
    return s;
}

void join_side(pycmf_ctx* ctx) {
    if (ctx->root != nullptr || !ctx->side_streams || ctx->side == nullptr) return;
    PYCMF_CUDA(cudaEventRecord(ctx->ev_join, ctx->side->stream));
    PYCMF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
}

namespace {

constexpr int SLOT_T0 = 4, SLOT_T1 = 5, SLOT_T2 = 6, SLOT_T3 = 7;

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename F>
int guarded(pycmf_ctx* ctx, F&& f) {
    try {
        PYCMF_CHECK(ctx != nullptr, "null context");
        DeviceGuard g(ctx->device);
        f();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return 1;
    }
}

template <typename T>
void copy_async(pycmf_ctx* ctx, T* dst, const T* src, size_t n) {
    PYCMF_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
}

// ---- objective ---------------------------------------------------------------------------------
template <typename T>
void sqerr_impl(pycmf_ctx* ctx, int64_t rows, int64_t m, int64_t k, const T* A, const T* B,
                const T* Tg, int64_t ldt, bool trans_t, const int32_t* rowptr, const int32_t* colidx,
                const T* vals, int link, double* out) {
    PYCMF_CUDA(cudaMemsetAsync(out, 0, sizeof(double), ctx->stream));
    if (rows <= 0 || m <= 0) return;
    if (Tg != nullptr) {
        resid_pass<T>(ctx, rows, m, k, A, B, Tg, ldt, trans_t, link, nullptr, nullptr, out);
        return;
    }
    PYCMF_CHECK(rowptr != nullptr && colidx != nullptr && vals != nullptr, "sqerr: no target given");
    if (link == PYCMF_LINEAR) {
        // ||T||^2 + tr((A^T A)(B^T B)) - 2 sum_nz t_ij a_i.b_j   (sklearn _beta_divergence sparse branch)
        double* G = static_cast<double*>(scratch(ctx, SLOT_T0, sizeof(double) * 2 * k * k));
        gram_f64<T>(ctx, rows, k, A, G);
        gram_f64<T>(ctx, m, k, B, G + k * k);
        dot_f64<double>(ctx, k * k, G, G + k * k, 1.0, out, true);
        sddmm_reduce<T>(ctx, 2, rows, rowptr, colidx, vals, A, B, k, 1.0, out);
        // cross term sum_nz t_ij a_i . b_j = <A, T B>: one SpMM (the nonzero-balanced kernel) and a dot product in float64;
        // the warp-per-row SDDMM reduction it replaces cost 1.18 ms on the C3 shard against 0.41 ms for the SpMM
        T* N = static_cast<T*>(scratch(ctx, SLOT_T1, sizeof(T) * size_t(rows) * k));
        spmm<T>(ctx, rows, rowptr, colidx, vals, B, k, k, N, k, T(1), T(0), m);
        dot_f64<T>(ctx, rows * k, A, N, -2.0, out, true);
    } else {
        // sum_all sigma^2 + sum_nz [ (t - sigma)^2 - sigma^2 ]
        resid_pass<T>(ctx, rows, m, k, A, B, nullptr, 0, false, link, nullptr, nullptr, out);
        sddmm_reduce<T>(ctx, 1, rows, rowptr, colidx, vals, A, B, k, 1.0, out);
    }
}

// ---- MU ------------------------------------------------------------------------------------------
// out (contiguous, k columns) = X Q (trans == false) or X^T Q on the tensor cores when the shape qualifies (fp32, k = 32 uses
// tc_resid.cu's pass, k = 64 .. 256 tc_mu.cu); false = caller runs the generic kernel.
template <typename T>
bool tc_try(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, int64_t k, const T* X, int64_t ldx, const T* Q, T* out) {
    if constexpr (std::is_same<T, float>::value) {
        if (tc_mu_eligible(ctx, rows, cols, k, X, ldx, false)) {
            tc_mu_xmul(ctx, trans, rows, cols, k, X, ldx, Q, out, "tc_factor");
            return true;
        }
    }
    return false;
}

template <typename T>
void mu_v_partial_impl(pycmf_ctx* ctx, int64_t n, int64_t d, int64_t k, const T* X, int64_t ldx,
                       const int32_t* colptr, const int32_t* rowidx, const T* cvals, const T* U, T* out) {
    // U^T U on the side stream, next to the pass over X
    pycmf_ctx* sc = fork_side(ctx);
    if (!tc_try<T>(sc, true, n, k, k, U, k, U, out + d * k))
        gemm<T>(sc, true, k, k, n, U, k, U, k, out + d * k, k, T(1), T(0));
    if (X != nullptr) {
        bool done = false;
        if constexpr (std::is_same<T, float>::value) {
            if (tc_dense_eligible(ctx, n, d, k, X, ldx, false)) {
                tc_xmul(ctx, true, n, d, X, ldx, U, out);
                done = true;
            } else if (tc_mu_eligible(ctx, n, d, k, X, ldx, false)) {
                tc_mu_xmul(ctx, true, n, d, k, X, ldx, U, out);
                done = true;
            }
        }
        if (!done) gemm<T>(ctx, true, d, k, n, X, ldx, U, k, out, k, T(1), T(0));
    } else {
        PYCMF_CHECK(colptr && rowidx && cvals, "mu_v_partial: neither dense X nor CSC arrays given");
        spmm<T>(ctx, d, colptr, rowidx, cvals, U, k, k, out, k, T(1), T(0), n);
    }
    join_side(ctx);
}

template <typename T>
void mu_v_apply_impl(pycmf_ctx* ctx, int64_t d, int64_t l, int64_t k, T* V, const T* xtu_utu, const T* Y,
                     int64_t ldy, const T* Z, double l1, double l2) {
    T* N = static_cast<T*>(scratch(ctx, SLOT_T0, sizeof(T) * size_t(d) * k));
    T* D = static_cast<T*>(scratch(ctx, SLOT_T1, sizeof(T) * size_t(d) * k));
    T* G = static_cast<T*>(scratch(ctx, SLOT_T2, sizeof(T) * size_t(k) * k));
    copy_async<T>(ctx, N, xtu_utu, size_t(d) * k);
    copy_async<T>(ctx, G, xtu_utu + d * k, size_t(k) * k);
    if (tc_try<T>(ctx, false, d, l, k, Y, ldy, Z, D))                   // + Y Z       (:244)
        axpby<T>(ctx, d * k, T(1), N, T(1), D, N);
    else
        gemm<T>(ctx, false, d, k, l, Y, ldy, Z, k, N, k, T(1), T(1));
    gemm<T>(ctx, true, k, k, l, Z, k, Z, k, G, k, T(1), T(1));           // + Z^T Z     (:245)
    if (mu_fused_apply<T>(ctx, d, k, V, N, G, l1, l2)) return;          // V (UtU+ZtZ) and the ratio in one kernel
    if (!tc_try<T>(ctx, false, d, k, k, V, k, G, D))                    // V (UtU+ZtZ) (:245)
        gemm<T>(ctx, false, d, k, k, V, k, G, k, D, k, T(1), T(0));
    mu_apply<T>(ctx, d, k, V, N, D, l1, l2);
}

template <typename T>
void mu_left_impl(pycmf_ctx* ctx, int64_t rows, int64_t m, int64_t k, T* F, const T* B, const T* Tg, int64_t ldt,
                  bool trans_t, const int32_t* rowptr, const int32_t* colidx, const T* vals, double l1, double l2) {
    T* N = static_cast<T*>(scratch(ctx, SLOT_T0, sizeof(T) * size_t(rows) * k));
    // denominator F (B^T B) on the side stream, next to the pass over the target
    pycmf_ctx* sc = fork_side(ctx);
    T* D = static_cast<T*>(scratch(sc, SLOT_T1, sizeof(T) * size_t(rows) * k));
    T* G = static_cast<T*>(scratch(sc, SLOT_T2, sizeof(T) * size_t(k) * k));
    if (!tc_try<T>(sc, true, m, k, k, B, k, B, G))                      // B^T B
        gemm<T>(sc, true, k, k, m, B, k, B, k, G, k, T(1), T(0));
    const bool fused = k <= 128 && ctx->mu_fused != 0;                  // F (B^T B) inside the ratio kernel below
    if (!fused && !tc_try<T>(sc, false, rows, k, k, F, k, G, D))        // F (B^T B)
        gemm<T>(sc, false, rows, k, k, F, k, G, k, D, k, T(1), T(0));
    if (Tg != nullptr) {
        bool done = false;
        if constexpr (std::is_same<T, float>::value) {
            if (tc_dense_eligible(ctx, rows, m, k, Tg, ldt, trans_t)) {
                tc_xmul(ctx, false, rows, m, Tg, ldt, B, N);
                done = true;
            } else if (!trans_t && tc_mu_eligible(ctx, rows, m, k, Tg, ldt, false)) {
                tc_mu_xmul(ctx, false, rows, m, k, Tg, ldt, B, N);
                done = true;
            } else if (trans_t && tc_mu_eligible(ctx, m, rows, k, Tg, ldt, false)) {
                tc_mu_xmul(ctx, true, m, rows, k, Tg, ldt, B, N, "tc_ytv");      // target stored transposed (m x rows): N = Tg^T B
                done = true;
            }
        }
        if (!done) gemm<T>(ctx, trans_t, rows, k, m, Tg, ldt, B, k, N, k, T(1), T(0));
    } else {
        PYCMF_CHECK(rowptr && colidx && vals, "mu_left: neither dense target nor CSR arrays given");
        spmm<T>(ctx, rows, rowptr, colidx, vals, B, k, k, N, k, T(1), T(0), m);
    }
    join_side(ctx);
    if (fused && mu_fused_apply<T>(ctx, rows, k, F, N, G, l1, l2)) return;
    if (fused) {                                                        // not eligible after all (shared memory): unfused
        if (!tc_try<T>(ctx, false, rows, k, k, F, k, G, D))
            gemm<T>(ctx, false, rows, k, k, F, k, G, k, D, k, T(1), T(0));
    }
    mu_apply<T>(ctx, rows, k, F, N, D, l1, l2);
}

// ---- Newton --------------------------------------------------------------------------------------
int64_t rows_per_chunk(pycmf_ctx* ctx, int64_t rows, int64_t k, size_t elem) {
    size_t per_row = size_t(k) * k * elem;
    int64_t c = int64_t(std::max<size_t>(1, (ctx->max_scratch / 2) / per_row));
    return std::max<int64_t>(1, std::min<int64_t>(rows, c));
}

template <typename T>
void newton_left_impl(pycmf_ctx* ctx, int64_t rows, int64_t m, int64_t k, T* F, const T* B, const T* Tg,
                      int64_t ldt, bool trans_t, const int32_t* rowptr, const int32_t* colidx, const T* vals,
                      double weight, double l1, double l2, int link, bool non_negative, double pert,
                      bool l2_in_logit, const int32_t* idx, int64_t n_sample) {
    if (rows <= 0) return;
    const bool sampled = idx != nullptr;
    const bool sparse = Tg == nullptr;
    if (sparse && !(sampled && n_sample == 0))
        PYCMF_CHECK(rowptr && colidx && vals, "newton_left: neither dense target nor CSR arrays given");
    const double l2_diag = (link == PYCMF_LINEAR || l2_in_logit) ? l2 : 0.0;
    T* g = static_cast<T*>(scratch(ctx, SLOT_T0, sizeof(T) * size_t(rows) * k));

    if (!sampled) {
        // shared Hessian weight * B^T B + l2 I (cmf_solvers.py:407-410): Gram in float64 and its clamped inverse on the
        // side stream, next to the gradient pass
        const double* Hinv = nullptr;
        if (link == PYCMF_LINEAR) {
            pycmf_ctx* sc = fork_side(ctx);
            double* G64 = static_cast<double*>(scratch(sc, SLOT_T1, sizeof(double) * size_t(k) * k));
            gram_f64<T>(sc, m, k, B, G64);
            Hinv = shared_inverse64(sc, k, G64, weight, l2_diag, pert);
        }
        // data gradient for every row from the pre-update factor (cmf_solvers.py:399-400)
        if (!sparse) {
            resid_pass<T>(ctx, rows, m, k, F, B, Tg, ldt, trans_t, link, g, nullptr, nullptr);
            axpby<T>(ctx, rows * k, T(weight), g, T(0), nullptr, g);
        } else {
            if (link == PYCMF_LINEAR) {
                T* G = static_cast<T*>(scratch(ctx, SLOT_T2, sizeof(T) * size_t(k) * k));
                gemm<T>(ctx, true, k, k, m, B, k, B, k, G, k, T(1), T(0));
                gemm<T>(ctx, false, rows, k, k, F, k, G, k, g, k, T(weight), T(0));
            } else {
                resid_pass<T>(ctx, rows, m, k, F, B, nullptr, 0, false, link, g, nullptr, nullptr);
                axpby<T>(ctx, rows * k, T(weight), g, T(0), nullptr, g);
            }
            spmm<T>(ctx, rows, rowptr, colidx, vals, B, k, k, g, k, T(-weight), T(1), m);
        }
        if (link == PYCMF_LINEAR) {
            join_side(ctx);
            apply_shared_inverse<T>(ctx, rows, k, F, g, Hinv, l1, l2, non_negative);
            return;
        }
    }
    // per-row Hessians (logit link and / or sampling), processed in bounded row chunks
    const int64_t chunk = rows_per_chunk(ctx, rows, k, sizeof(T));
    T* H = static_cast<T*>(scratch(ctx, SLOT_T1, sizeof(T) * size_t(chunk) * k * k));
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t rc = std::min(chunk, rows - r0);
        const T* Tc = Tg == nullptr ? nullptr : (trans_t ? Tg + r0 : Tg + r0 * ldt);
        const int32_t* rp = rowptr ? rowptr + r0 : nullptr;
        const int32_t* ix = idx ? idx + r0 * n_sample : nullptr;
        row_grad_hess<T>(ctx, rc, m, k, F + r0 * k, B, Tc, ldt, trans_t, rp, colidx, vals, link, weight, ix,
                         n_sample, sampled ? g + r0 * k : nullptr, H, false);
        // H_i = weight * (PSD weighted Gram) + l2_diag I: lambda_min >= l2_diag, so the clamp test can be skipped
        newton_solve_rows<T>(ctx, rc, k, F + r0 * k, g + r0 * k, H, k * k, l1, l2, l2_diag, pert, non_negative,
                             weight >= 0.0 && l2_diag >= pert);
    }
}

template <typename T>
void newton_v_xpart_impl(pycmf_ctx* ctx, int64_t d_rows, int64_t n, int64_t k, const T* V, const T* U,
                         const T* Xc, int64_t ldx, const int32_t* colptr, const int32_t* rowidx, const T* cvals,
                         int x_link, double alpha, const int32_t* idx, int64_t n_sample, T* gx, void* Hx_any,
                         int* hx_per_row) {
    // Hx: per-row Hessians (d_rows x k x k, compute dtype) or ONE shared k x k matrix in FLOAT64 (linear link, no
    // sampling): alpha U^T U rounded to fp32 costs 6e-8 x cond(H) on the Newton step -- 8e-3 on V in C2's first iteration
    T* Hx = static_cast<T*>(Hx_any);
    const bool sampled = idx != nullptr;
    const bool sparse = Xc == nullptr;
    if (sparse && !(sampled && n_sample == 0))
        PYCMF_CHECK(colptr && rowidx && cvals, "newton_v_xpart: neither dense X nor CSC arrays given");
    const bool per_row = sampled || x_link == PYCMF_LOGIT;
    if (hx_per_row) *hx_per_row = per_row ? 1 : 0;
    if (d_rows <= 0) return;
    if (sampled) {
        // element (j, i) of the target is X[i, j]: dense X read transposed, or row j of the CSC arrays
        row_grad_hess<T>(ctx, d_rows, n, k, V, U, Xc, ldx, true, colptr, rowidx, cvals, x_link, alpha, idx,
                         n_sample, gx, Hx, false);
        return;
    }
    if (x_link == PYCMF_LINEAR) {
        // shared Hessian part alpha U^T U on the side stream, next to the gradient pass
        pycmf_ctx* sc = fork_side(ctx);
        double* Hx64 = static_cast<double*>(Hx_any);
        gram_f64<T>(sc, n, k, U, Hx64);
        axpby<double>(sc, k * k, alpha, Hx64, 0.0, nullptr, Hx64);
    }
    if (!sparse) {
        resid_pass<T>(ctx, n, d_rows, k, U, V, Xc, ldx, false, x_link, nullptr, gx, nullptr);
        axpby<T>(ctx, d_rows * k, T(alpha), gx, T(0), nullptr, gx);
    } else {
        if (x_link == PYCMF_LINEAR) {
            T* G = static_cast<T*>(scratch(ctx, SLOT_T2, sizeof(T) * size_t(k) * k));
            gemm<T>(ctx, true, k, k, n, U, k, U, k, G, k, T(1), T(0));
            gemm<T>(ctx, false, d_rows, k, k, V, k, G, k, gx, k, T(alpha), T(0));
        } else {
            resid_pass<T>(ctx, n, d_rows, k, U, V, nullptr, 0, false, x_link, nullptr, gx, nullptr);
            axpby<T>(ctx, d_rows * k, T(alpha), gx, T(0), nullptr, gx);
        }
        spmm<T>(ctx, d_rows, colptr, rowidx, cvals, U, k, k, gx, k, T(-alpha), T(1), n);
    }
    if (x_link == PYCMF_LINEAR) {
        join_side(ctx);
    } else {
        row_grad_hess<T>(ctx, d_rows, n, k, V, U, nullptr, 0, false, nullptr, nullptr, nullptr, x_link, alpha,
                         nullptr, 0, nullptr, Hx, false);
    }
}

template <typename T>
void newton_v_finish_impl(pycmf_ctx* ctx, int64_t d_rows, int64_t l, int64_t k, T* V, const T* Z, const T* Yr,
                          int64_t ldy, int y_link, double alpha, double l1, double l2, const int32_t* idx,
                          int64_t n_sample, const T* gx, const void* Hx_any, bool hx_per_row, bool non_negative,
                          double pert) {
    if (d_rows <= 0) return;
    const bool sampled = idx != nullptr;
    const double wy = 1.0 - alpha;
    if (!sampled && newton_finish_small<T>(ctx, d_rows, l, k, V, Z, Yr, ldy, y_link, wy, gx, Hx_any, hx_per_row, l1, l2,
                                           l2, pert, non_negative))
        return;
    const bool per_row = hx_per_row || sampled || y_link == PYCMF_LOGIT;
    // a shared X-side Hessian arrives in float64 and stays there: it is added to the per-row label parts inside the
    // solve (Hbase), or, with a shared label part as well, the whole shared Hessian is inverted once in float64
    const T* Hx = hx_per_row ? static_cast<const T*>(Hx_any) : nullptr;
    const double* Hx64 = hx_per_row ? nullptr : static_cast<const double*>(Hx_any);
    T* g = static_cast<T*>(scratch(ctx, SLOT_T0, sizeof(T) * size_t(d_rows) * k));
    if (!per_row) {
        T* gy = static_cast<T*>(scratch(ctx, SLOT_T2, sizeof(T) * size_t(d_rows) * k));
        resid_pass<T>(ctx, d_rows, l, k, V, Z, Yr, ldy, false, y_link, gy, nullptr, nullptr);
        axpby<T>(ctx, d_rows * k, T(wy), gy, T(1), gx, g);
        double* H64 = static_cast<double*>(scratch(ctx, SLOT_T1, sizeof(double) * size_t(k) * k));   // (arena 7 = the inverse)
        gram_f64<T>(ctx, l, k, Z, H64);                                           // Z^T Z
        axpby<double>(ctx, k * k, wy, H64, 1.0, Hx64, H64);                       // (1 - alpha) Z^T Z + alpha U^T U
        const double* Hinv = shared_inverse64(ctx, k, H64, 1.0, l2, pert);
        apply_shared_inverse<T>(ctx, d_rows, k, V, g, Hinv, l1, l2, non_negative);
        return;
    }
    T* H = static_cast<T*>(scratch(ctx, SLOT_T1, sizeof(T) * size_t(d_rows) * k * k));
    if (hx_per_row) copy_async<T>(ctx, H, Hx, size_t(d_rows) * k * k);
    else PYCMF_CUDA(cudaMemsetAsync(H, 0, sizeof(T) * size_t(d_rows) * k * k, ctx->stream));
    if (sampled) {
        copy_async<T>(ctx, g, gx, size_t(d_rows) * k);
        row_grad_hess<T>(ctx, d_rows, l, k, V, Z, Yr, ldy, false, nullptr, nullptr, nullptr, y_link, wy, idx,
                         n_sample, g, H, true);
    } else {
        T* gy = static_cast<T*>(scratch(ctx, SLOT_T2, sizeof(T) * size_t(d_rows) * k));
        resid_pass<T>(ctx, d_rows, l, k, V, Z, Yr, ldy, false, y_link, gy, nullptr, nullptr);
        axpby<T>(ctx, d_rows * k, T(wy), gy, T(1), gx, g);
        if (y_link == PYCMF_LOGIT) {
            row_grad_hess<T>(ctx, d_rows, l, k, V, Z, nullptr, 0, false, nullptr, nullptr, nullptr, y_link, wy,
                             nullptr, 0, nullptr, H, true);
        } else {
            T* Gz = static_cast<T*>(scratch(ctx, SLOT_T3, sizeof(T) * size_t(k) * k));
            gemm<T>(ctx, true, k, k, l, Z, k, Z, k, Gz, k, T(wy), T(0));
            broadcast_add<T>(ctx, d_rows, k * k, H, Gz, T(1), false);
        }
    }
    newton_solve_rows<T>(ctx, d_rows, k, V, g, H, k * k, l1, l2, l2, pert, non_negative, false, Hx64);
}

}  // namespace
}  // namespace pycmf

using namespace pycmf;

#define DISPATCH(dtype, CALL_F32, CALL_F64)                                    \
    do {                                                                       \
        if ((dtype) == PYCMF_F32) { CALL_F32; }                                \
        else if ((dtype) == PYCMF_F64) { CALL_F64; }                           \
        else throw pycmf::Error("unknown dtype code");                         \
    } while (0)

static void check_link(int link) {
    if (link != PYCMF_LINEAR && link != PYCMF_LOGIT) throw pycmf::Error("Invalid link function code");
}

static void clear_timers(pycmf_ctx* ctx) {
    for (auto& kv : ctx->timers)
        for (auto& pr : kv.second) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    ctx->timers.clear();
}

extern "C" {

int pycmf_abi_version(void) { return PYCMF_ABI_VERSION; }

const char* pycmf_last_error(void) { return g_last_error.c_str(); }

int pycmf_create(int device, pycmf_ctx** out) {
    try {
        PYCMF_CHECK(out != nullptr, "null out pointer");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0)
            throw pycmf::Error("pycmf_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
        PYCMF_CHECK(device >= 0 && device < count, "bad device index");
        DeviceGuard g(device);
        cudaDeviceProp prop;
        PYCMF_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw pycmf::Error(std::string("pycmf_b200 is built for sm_100a only; device is sm_") +
                               std::to_string(prop.major) + std::to_string(prop.minor));
        pycmf_ctx* c = new pycmf_ctx();
        c->device = device;
        c->num_sms = prop.multiProcessorCount;
        c->max_smem_optin = int(prop.sharedMemPerBlockOptin);
        *out = c;
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return 1;
    }
}

int pycmf_destroy(pycmf_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    clear_timers(ctx);
    for (auto& a : ctx->arena)
        if (a.ptr) cudaFree(a.ptr);
    for (void* p : ctx->retired) cudaFree(p);
    if (ctx->side != nullptr) {
        for (auto& a : ctx->side->arena)
            if (a.ptr) cudaFree(a.ptr);
        for (void* p : ctx->side->retired) cudaFree(p);
        cudaStreamDestroy(ctx->side->stream);
        delete ctx->side;
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
    return 0;
}

int pycmf_set_stream(pycmf_ctx* ctx, void* stream) {
    return guarded(ctx, [&] { ctx->stream = static_cast<cudaStream_t>(stream); });
}

int pycmf_set_option(pycmf_ctx* ctx, const char* key, double value) {
    return guarded(ctx, [&] {
        std::string k(key ? key : "");
        if (k == "chol_fastpath") ctx->chol_fastpath = value != 0.0;
        else if (k == "dense_path") ctx->dense_path = int(value);
        else if (k == "tc_max_splits") ctx->tc_max_splits = int(value);
        else if (k == "spmm_path") ctx->spmm_path = int(value);
        else if (k == "spmm_blocks_per_sm") ctx->spmm_blocks_per_sm = int(value);
        else if (k == "spmm_unroll") ctx->spmm_unroll = int(value);
        else if (k == "spmm_lean") ctx->spmm_lean = int(value);
        else if (k == "mu_fused") ctx->mu_fused = int(value);
        else if (k == "solve_path") ctx->solve_path = int(value);
        else if (k == "hess_mma") ctx->hess_mma = int(value);
        else if (k == "solve_threads") ctx->solve_threads = int(value);
        else if (k == "tc_trace") ctx->tc_trace = int(value);
        else if (k == "tc_ctas") ctx->tc_ctas = int(value);
        else if (k == "tc_chain") ctx->tc_chain = int(value);
        else if (k == "tc_x_promotion") ctx->tc_x_promotion = int(value);
        else if (k == "tc_prefetch") ctx->tc_prefetch = int(value);
        else if (k == "side_streams") ctx->side_streams = value != 0.0;
        else if (k == "finish_minblocks") ctx->finish_minblocks = int(value);
        else if (k == "max_scratch_mb") ctx->max_scratch = size_t(std::max(16.0, value)) << 20;
        else throw pycmf::Error("unknown option: " + k);
    });
}

int64_t pycmf_launch_count(pycmf_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pycmf_debug_tc_trace(pycmf_ctx* ctx, int64_t* host, int64_t max_words) {
    return guarded(ctx, [&] {
        PYCMF_CUDA(cudaStreamSynchronize(ctx->stream));
        PYCMF_CHECK(ctx->arena[2].ptr != nullptr, "no trace recorded (set option tc_trace = 1 first)");
        int64_t n = std::min<int64_t>(max_words, tc_trace_words());
        PYCMF_CUDA(cudaMemcpy(host, ctx->arena[2].ptr, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
    });
}

int pycmf_profile_enable(pycmf_ctx* ctx, int on) {
    return guarded(ctx, [&] { ctx->profile = on != 0; });
}

int pycmf_profile_reset(pycmf_ctx* ctx) {
    return guarded(ctx, [&] {
        PYCMF_CUDA(cudaStreamSynchronize(ctx->stream));
        clear_timers(ctx);
    });
}

int pycmf_profile_query(pycmf_ctx* ctx, const char* family, double* total_ms, int64_t* count) {
    return guarded(ctx, [&] {
        PYCMF_CUDA(cudaStreamSynchronize(ctx->stream));
        double tot = 0.0;
        int64_t n = 0;
        auto it = ctx->timers.find(family ? family : "");
        if (it != ctx->timers.end()) {
            for (auto& pr : it->second) {
                float ms = 0.f;
                PYCMF_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
                tot += ms;
                n++;
            }
        }
        if (total_ms) *total_ms = tot;
        if (count) *count = n;
    });
}

int pycmf_gemm(pycmf_ctx* ctx, int dtype, int trans_a, int64_t m, int64_t q, int64_t p, const void* A,
               int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, double alpha, double beta) {
    return guarded(ctx, [&] {
        DISPATCH(dtype,
                 gemm<float>(ctx, trans_a != 0, m, q, p, (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc,
                             float(alpha), float(beta)),
                 gemm<double>(ctx, trans_a != 0, m, q, p, (const double*)A, lda, (const double*)B, ldb, (double*)C,
                              ldc, alpha, beta));
    });
}

int pycmf_spmm(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t cols, const int32_t* rowptr,
               const int32_t* colidx, const void* vals, const void* B, int64_t ldb, int64_t k, void* C,
               int64_t ldc, double alpha, double beta) {
    (void)cols;
    return guarded(ctx, [&] {
        DISPATCH(dtype,
                 spmm<float>(ctx, rows, rowptr, colidx, (const float*)vals, (const float*)B, ldb, k, (float*)C, ldc,
                             float(alpha), float(beta), cols),
                 spmm<double>(ctx, rows, rowptr, colidx, (const double*)vals, (const double*)B, ldb, k, (double*)C,
                              ldc, alpha, beta, cols));
    });
}

int pycmf_resid_pass(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k, const void* A, const void* B,
                     const void* T, int64_t ldt, int trans_t, int link, void* outL, void* outR, double* sq) {
    return guarded(ctx, [&] {
        check_link(link);
        DISPATCH(dtype,
                 resid_pass<float>(ctx, rows, m, k, (const float*)A, (const float*)B, (const float*)T, ldt, trans_t != 0,
                                   link, (float*)outL, (float*)outR, sq),
                 resid_pass<double>(ctx, rows, m, k, (const double*)A, (const double*)B, (const double*)T, ldt,
                                    trans_t != 0, link, (double*)outL, (double*)outR, sq));
    });
}

int pycmf_sqerr(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k, const void* A, const void* B,
                const void* T, int64_t ldt, int trans_t, const int32_t* rowptr, const int32_t* colidx,
                const void* vals, int link, double* out_sq) {
    return guarded(ctx, [&] {
        check_link(link);
        DISPATCH(dtype,
                 sqerr_impl<float>(ctx, rows, m, k, (const float*)A, (const float*)B, (const float*)T, ldt,
                                   trans_t != 0, rowptr, colidx, (const float*)vals, link, out_sq),
                 sqerr_impl<double>(ctx, rows, m, k, (const double*)A, (const double*)B, (const double*)T, ldt,
                                    trans_t != 0, rowptr, colidx, (const double*)vals, link, out_sq));
    });
}

int pycmf_mu_v_partial(pycmf_ctx* ctx, int dtype, int64_t n, int64_t d, int64_t k, const void* X, int64_t ldx,
                       const int32_t* csc_colptr, const int32_t* csc_rowidx, const void* csc_vals,
                       const void* U, void* out) {
    return guarded(ctx, [&] {
        DISPATCH(dtype,
                 mu_v_partial_impl<float>(ctx, n, d, k, (const float*)X, ldx, csc_colptr, csc_rowidx,
                                          (const float*)csc_vals, (const float*)U, (float*)out),
                 mu_v_partial_impl<double>(ctx, n, d, k, (const double*)X, ldx, csc_colptr, csc_rowidx,
                                           (const double*)csc_vals, (const double*)U, (double*)out));
    });
}

int pycmf_mu_v_apply(pycmf_ctx* ctx, int dtype, int64_t d, int64_t l, int64_t k, void* V, const void* xtu_utu,
                     const void* Y, int64_t ldy, const void* Z, double l1_reg, double l2_reg) {
    return guarded(ctx, [&] {
        DISPATCH(dtype,
                 mu_v_apply_impl<float>(ctx, d, l, k, (float*)V, (const float*)xtu_utu, (const float*)Y, ldy,
                                        (const float*)Z, l1_reg, l2_reg),
                 mu_v_apply_impl<double>(ctx, d, l, k, (double*)V, (const double*)xtu_utu, (const double*)Y, ldy,
                                         (const double*)Z, l1_reg, l2_reg));
    });
}

int pycmf_mu_left(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k, void* F, const void* B,
                  const void* T, int64_t ldt, int trans_t, const int32_t* rowptr, const int32_t* colidx,
                  const void* vals, double l1_reg, double l2_reg) {
    return guarded(ctx, [&] {
        DISPATCH(dtype,
                 mu_left_impl<float>(ctx, rows, m, k, (float*)F, (const float*)B, (const float*)T, ldt, trans_t != 0,
                                     rowptr, colidx, (const float*)vals, l1_reg, l2_reg),
                 mu_left_impl<double>(ctx, rows, m, k, (double*)F, (const double*)B, (const double*)T, ldt,
                                      trans_t != 0, rowptr, colidx, (const double*)vals, l1_reg, l2_reg));
    });
}

int pycmf_newton_left(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t m, int64_t k, void* F, const void* B,
                      const void* T, int64_t ldt, int trans_t, const int32_t* rowptr, const int32_t* colidx,
                      const void* vals, double weight, double l1_reg, double l2_reg, int link, int non_negative,
                      double hessian_pertubation, int l2_in_logit_hessian, const int32_t* sample_idx,
                      int64_t n_sample) {
    return guarded(ctx, [&] {
        check_link(link);
        DISPATCH(dtype,
                 newton_left_impl<float>(ctx, rows, m, k, (float*)F, (const float*)B, (const float*)T, ldt,
                                         trans_t != 0, rowptr, colidx, (const float*)vals, weight, l1_reg, l2_reg,
                                         link, non_negative != 0, hessian_pertubation, l2_in_logit_hessian != 0,
                                         sample_idx, n_sample),
                 newton_left_impl<double>(ctx, rows, m, k, (double*)F, (const double*)B, (const double*)T, ldt,
                                          trans_t != 0, rowptr, colidx, (const double*)vals, weight, l1_reg, l2_reg,
                                          link, non_negative != 0, hessian_pertubation, l2_in_logit_hessian != 0,
                                          sample_idx, n_sample));
    });
}

int pycmf_newton_v_xpart(pycmf_ctx* ctx, int dtype, int64_t d_rows, int64_t n, int64_t k, const void* V,
                         const void* U, const void* Xcols, int64_t ldx, const int32_t* csc_colptr,
                         const int32_t* csc_rowidx, const void* csc_vals, int x_link, double alpha,
                         const int32_t* sample_idx_x, int64_t n_sample_x, void* gx, void* Hx, int* hx_per_row) {
    return guarded(ctx, [&] {
        check_link(x_link);
        DISPATCH(dtype,
                 newton_v_xpart_impl<float>(ctx, d_rows, n, k, (const float*)V, (const float*)U, (const float*)Xcols,
                                            ldx, csc_colptr, csc_rowidx, (const float*)csc_vals, x_link, alpha,
                                            sample_idx_x, n_sample_x, (float*)gx, Hx, hx_per_row),
                 newton_v_xpart_impl<double>(ctx, d_rows, n, k, (const double*)V, (const double*)U,
                                             (const double*)Xcols, ldx, csc_colptr, csc_rowidx,
                                             (const double*)csc_vals, x_link, alpha, sample_idx_x, n_sample_x,
                                             (double*)gx, Hx, hx_per_row));
    });
}

int pycmf_newton_v_finish(pycmf_ctx* ctx, int dtype, int64_t d_rows, int64_t l, int64_t k, void* V, const void* Z,
                          const void* Yrows, int64_t ldy, int y_link, double alpha, double l1_reg, double l2_reg,
                          const int32_t* sample_idx_y, int64_t n_sample_y, const void* gx, const void* Hx,
                          int hx_per_row, int non_negative, double hessian_pertubation) {
    return guarded(ctx, [&] {
        check_link(y_link);
        DISPATCH(dtype,
                 newton_v_finish_impl<float>(ctx, d_rows, l, k, (float*)V, (const float*)Z, (const float*)Yrows, ldy,
                                             y_link, alpha, l1_reg, l2_reg, sample_idx_y, n_sample_y,
                                             (const float*)gx, Hx, hx_per_row != 0, non_negative != 0,
                                             hessian_pertubation),
                 newton_v_finish_impl<double>(ctx, d_rows, l, k, (double*)V, (const double*)Z, (const double*)Yrows,
                                              ldy, y_link, alpha, l1_reg, l2_reg, sample_idx_y, n_sample_y,
                                              (const double*)gx, Hx, hx_per_row != 0,
                                              non_negative != 0, hessian_pertubation));
    });
}

int pycmf_safe_solve(pycmf_ctx* ctx, int64_t batch, int64_t k, const double* H, int64_t h_stride, const double* g,
                     double* x, double hessian_pertubation) {
    return guarded(ctx, [&] { safe_solve_f64(ctx, batch, k, H, h_stride, g, x, hessian_pertubation); });
}

int pycmf_sample_indices(pycmf_ctx* ctx, int64_t rows, int64_t N, int64_t n_sample, uint64_t seed,
                         uint64_t stream_id, int32_t* idx) {
    return guarded(ctx, [&] { sample_indices(ctx, rows, 0, N, n_sample, seed, stream_id, 0, 0, idx); });
}

int pycmf_topk_columns(pycmf_ctx* ctx, int dtype, int64_t rows, int64_t k, const void* F, int64_t ld, int64_t topn,
                       int32_t* out) {
    return guarded(ctx, [&] {
        DISPATCH(dtype, topk_columns<float>(ctx, rows, k, (const float*)F, ld, int(topn), out),
                 topk_columns<double>(ctx, rows, k, (const double*)F, ld, int(topn), out));
    });
}

int pycmf_sample_indices_sharded(pycmf_ctx* ctx, int64_t rows, int64_t row0, int64_t N, int64_t n_sample, uint64_t seed,
                                 uint64_t stream_id, int64_t lo, int64_t hi, int32_t* idx) {
    return guarded(ctx, [&] { sample_indices(ctx, rows, row0, N, n_sample, seed, stream_id, lo, hi, idx); });
}

}  // extern "C"

// Fused finish of the Newton V update for small n_components (k <= 32): one WARP per row of V does
//   g_j = gx_j + w (f2(v_j Z^T) - Y[j,:]) Z + l1 sign(v_j) + l2 v_j
//   H_j = Hx(_j) + w Z^T diag(f2'(v_j Z^T)) Z + l2 I
//   v_j <- v_j - S(H_j) g_j ; optional clamp                           (reference cmf_solvers.py:432-486)
// entirely on chip: the row of H lives in registers (lane = Hessian row), the factorisation runs in a
// per-warp shared-memory tile with warp-synchronous Cholesky (fast path, lambda_min(H) > pert) or a
// warp-level one-sided Jacobi (eigenvalue clamp active).  Replaces five launches of the generic path
// (small fused-residual pass, axpby, Hessian broadcast, per-row Hessian, batched solve) and the
// d x k x k Hessian round trip through HBM.  All arithmetic in float64.
#include "warp_solve.cuh"

namespace pycmf {
namespace {

using wsolve::KS;
using wsolve::WLD;
using wsolve::shfl_d;
constexpr int WARPS = 4;
constexpr int LMAX = 128;         // max rows of the small factor (labels)

// pd_mode: 0 = test every row (Cholesky of H - pert I), 1 = read the shared verdict from *pd_flag, 2 = known PD
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
newton_finish_small_kernel(int64_t rows, int l, int k, T* __restrict__ F, const T* __restrict__ Z,
                           const T* __restrict__ Y, int64_t ldy, int y_link, T wy,
                           const T* __restrict__ gx, const T* __restrict__ Hx, int64_t hx_stride,
                           double l1, double l2, double l2_diag, double pert, bool non_negative, bool chol_fastpath,
                           int pd_mode, const int* __restrict__ pd_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Wall = reinterpret_cast<double*>(smem_raw);          // WARPS x KS x WLD solve tiles
    T* Zs = reinterpret_cast<T*>(Wall + WARPS * KS * WLD);       // l x (k + 1)
    T* Hs = Zs + size_t(l) * (k + 1);                            // k x k (shared Hessian part, if hx_stride == 0)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kz = k + 1;
    for (int e = threadIdx.x; e < l * k; e += blockDim.x) Zs[(e / k) * kz + (e % k)] = Z[e];
    if (hx_stride == 0)
        for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
            int r = e / k, c = e % k;
            Hs[e] = Hx[(r > c ? r : c) * k + (r > c ? c : r)];   // lower triangle, like eigh
        }
    __syncthreads();
    double* W = Wall + warp * (KS * WLD);
    const bool act = lane < k;
    const bool known_pd = pd_mode == 2 || (pd_mode == 1 && *pd_flag != 0);
    for (int64_t row = int64_t(blockIdx.x) * WARPS + warp; row < rows; row += int64_t(gridDim.x) * WARPS) {
        const T v = act ? F[row * k + lane] : T(0);
        T g = act ? gx[row * k + lane] : T(0);
        // ---- estimates for the labels c2 = lane + 32 t
        T res[LMAX / 32], wgt[LMAX / 32];
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            const int c2 = lane + 32 * t;
            if (32 * t >= l) { res[t] = T(0); wgt[t] = T(0); continue; }
            T d = T(0);
            for (int a = 0; a < k; a++) d = fma(T(__shfl_sync(0xffffffffu, v, a)), (c2 < l) ? Zs[c2 * kz + a] : T(0), d);
            T est = d, fp = T(1);
            if (y_link == PYCMF_LOGIT) { est = sigmoid_<T>(d); fp = est * (T(1) - est); }
            const T y = (c2 < l) ? Y[row * ldy + c2] : T(0);
            res[t] = (c2 < l) ? wy * (est - y) : T(0);
            wgt[t] = (c2 < l) ? wy * fp : T(0);
        }
        // ---- row `lane` of the Hessian in registers (compute dtype)
        T Wr[KS];
#pragma unroll
        for (int c = 0; c < KS; c++) {
            T h = T(0);
            if (act && c < k) {
                if (hx_stride == 0) h = Hs[lane * k + c];
                else {
                    const int hi = lane > c ? lane : c, lo = lane > c ? c : lane;
                    h = Hx[row * hx_stride + hi * k + lo];
                }
            }
            Wr[c] = h;
        }
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            const int lim = min(32, l - 32 * t);
            for (int cc = 0; cc < lim; cc++) {
                const int c2 = 32 * t + cc;
                const T r = __shfl_sync(0xffffffffu, res[t], cc), w = __shfl_sync(0xffffffffu, wgt[t], cc);
                const T za = act ? Zs[c2 * kz + lane] : T(0);
                g = fma(r, za, g);
                const T wza = w * za;
#pragma unroll
                for (int c = 0; c < KS; c++)
                    if (c < k) Wr[c] = fma(wza, Zs[c2 * kz + c], Wr[c]);
            }
        }
        const double vd = double(v);
        const double sgn = vd > 0.0 ? 1.0 : (vd < 0.0 ? -1.0 : 0.0);
        const double gfull = act ? double(g) + l1 * sgn + l2 * vd : 0.0;
        // ---- l2 on the diagonal, then the clamped solve in float64
        double Wd[KS];
#pragma unroll
        for (int c = 0; c < KS; c++) Wd[c] = double(Wr[c]) + ((c == lane) ? l2_diag : 0.0);
        const double x = wsolve::safe_solve_warp(Wd, k, lane, gfull, pert, chol_fastpath, known_pd, W);
        if (act) {
            double f = vd - x;
            if (non_negative && f < 0.0) f = 0.0;
            F[row * k + lane] = T(f);
        }
    }
}

// ---- CTA-cooperative variants (32 x 32 threads, thread (r, c) owns one matrix element): used when there are only a
// few matrices, where the one-warp-per-matrix kernels are pure latency (~30 us for a 32 x 32 factorisation).
// LDL^T without square roots: step j updates W[r][c] -= W[r][j] W[c][j] / d_j for j < c <= r, then scales column j.
// Returns false (uniformly) when a pivot is <= floor.  On success: unit-lower L below the diagonal, D on it.
__device__ bool ldl_cta(double* W, int k, int r, int c, double floor) {
    for (int j = 0; j < k; j++) {
        __syncthreads();
        const double piv = W[j * wsolve::WLD + j];
        if (!(piv > floor)) return false;
        if (r < k && c > j && c <= r) W[r * wsolve::WLD + c] -= W[r * wsolve::WLD + j] * W[c * wsolve::WLD + j] / piv;
        __syncthreads();
        if (c == j && r > j && r < k) W[r * wsolve::WLD + j] /= piv;
    }
    __syncthreads();
    return true;
}

template <typename T>
__device__ void load_tile_cta(double* W, const T* __restrict__ H, int k, int r, int c, double scale, double diag) {
    if (r < k && c <= r) W[r * wsolve::WLD + c] = scale * double(H[r * k + c]) + (r == c ? diag : 0.0);
}

// *flag = 1 iff scale * H + (diag - pert) I is positive definite (H is k x k, lower triangle used)
template <typename T>
__global__ void __launch_bounds__(1024)
pd_flag_kernel(int k, const T* __restrict__ H, double scale, double diag, double pert, int* flag) {
    __shared__ double W[KS * WLD];
    __shared__ double red[32];
    const int r = threadIdx.y, c = threadIdx.x;
    load_tile_cta<T>(W, H, k, r, c, scale, diag - pert);
    __syncthreads();
    double tr = (r == 0 && c < k) ? fabs(W[c * WLD + c]) : 0.0;
    tr = warp_sum(tr);
    if (r == 0 && c == 0) red[0] = tr;
    __syncthreads();
    const bool ok = ldl_cta(W, k, r, c, 1e-13 * (red[0] + pert));
    if (r == 0 && c == 0) *flag = ok ? 1 : 0;
}

// One CTA per matrix, nrhs right-hand sides per matrix (warp w takes rhs w, w + 32, ...).
// MODE 0: X[b][q] = S(scale H_b + diag I) G[b][q].   MODE 1 (nrhs == 1): Newton row update of out[b].
template <typename T, int MODE>
__global__ void __launch_bounds__(1024)
safe_solve_cta_kernel(int64_t batch, int k, const T* __restrict__ H, int64_t h_stride, const T* __restrict__ g,
                      T* __restrict__ out, int nrhs, double l1, double l2, double l2_diag, double pert,
                      bool non_negative, bool chol_fastpath, double h_scale, bool known_pd) {
    __shared__ double W[KS * WLD];
    __shared__ double red[32];
    __shared__ int ok_s;
    const int r = threadIdx.y, c = threadIdx.x, lane = c, warp = r;
    const bool act = lane < k;
    for (int64_t b = blockIdx.x; b < batch; b += gridDim.x) {
        const T* Hb = H + b * h_stride;
        bool ok = false;
        __syncthreads();
        if (chol_fastpath) {
            ok = known_pd;
            if (!ok) {
                load_tile_cta<T>(W, Hb, k, r, c, h_scale, l2_diag - pert);
                __syncthreads();
                double tr = (r == 0 && c < k) ? fabs(W[c * WLD + c]) : 0.0;
                tr = warp_sum(tr);
                if (r == 0 && c == 0) red[0] = tr;
                __syncthreads();
                ok = ldl_cta(W, k, r, c, 1e-13 * (red[0] + pert));
                __syncthreads();
            }
            if (ok) {
                load_tile_cta<T>(W, Hb, k, r, c, h_scale, l2_diag);
                ok = ldl_cta(W, k, r, c, 0.0);
            }
        }
        if (!ok) {
            // eigenvalue clamp active: symmetric tile, Jacobi sweeps by warp 0, then every warp applies it to its rhs
            __syncthreads();
            if (r < k && c < k) {
                const int hi = r > c ? r : c, lo = r > c ? c : r;
                W[r * WLD + c] = h_scale * double(Hb[hi * k + lo]) + (r == c ? l2_diag : 0.0);
            }
            __syncthreads();
            if (warp == 0) wsolve::jacobi_sweeps_tile(W, k, lane, pert);
            __syncthreads();
        }
        for (int q = warp; q < nrhs; q += 32) {
            const int64_t gi = (b * nrhs + q) * k + lane;
            double gr = act ? double(g[gi]) : 0.0, f = 0.0;
            if (MODE == 1 && act) {
                f = double(out[gi]);
                gr += l1 * (f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0)) + l2 * f;
            }
            double x;
            if (ok) {
                double bv = gr;
                for (int j = 0; j < k; j++) {                       // L y = b (unit lower)
                    const double y = wsolve::shfl_d(bv, j);
                    if (lane > j && act) bv = fma(-W[lane * WLD + j], y, bv);
                }
                if (act) bv /= W[lane * WLD + lane];                // D z = y
                for (int j = k - 1; j >= 0; j--) {                  // L^T x = z
                    const double xv = wsolve::shfl_d(bv, j);
                    if (lane < j) bv = fma(-W[j * WLD + lane], xv, bv);
                }
                x = bv;
            } else {
                x = wsolve::jacobi_apply_tile(W, k, lane, gr, pert);
            }
            if (act) {
                if (MODE == 0) out[gi] = T(x);
                else {
                    double fn = f - x;
                    if (non_negative && fn < 0.0) fn = 0.0;
                    out[gi] = T(fn);
                }
            }
        }
    }
}

// Warp-per-matrix clamped solve for k <= 32.  MODE 0: x_b = S(scale H_b + diag I) g_b.  MODE 1: Newton row update.
template <typename T, int MODE>
__global__ void __launch_bounds__(WARPS * 32)
safe_solve_small_kernel(int64_t batch, int k, const T* __restrict__ H, int64_t h_stride, const T* __restrict__ g,
                        T* __restrict__ out, double l1, double l2, double l2_diag, double pert, bool non_negative,
                        bool chol_fastpath, double h_scale, bool known_pd) {
    __shared__ double Wall[WARPS * KS * WLD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* W = Wall + warp * (KS * WLD);
    const bool act = lane < k;
    for (int64_t b = int64_t(blockIdx.x) * WARPS + warp; b < batch; b += int64_t(gridDim.x) * WARPS) {
        const T* Hb = H + b * h_stride;
        double Hr[KS];
#pragma unroll
        for (int c = 0; c < KS; c++) {
            double h = 0.0;
            if (act && c < k) {
                const int hi = lane > c ? lane : c, lo = lane > c ? c : lane;
                h = h_scale * double(Hb[hi * k + lo]);
                if (c == lane) h += l2_diag;
            }
            Hr[c] = h;
        }
        double gr = act ? double(g[b * k + lane]) : 0.0;
        double f = 0.0;
        if (MODE == 1 && act) {
            f = double(out[b * k + lane]);
            gr += l1 * (f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0)) + l2 * f;
        }
        const double x = wsolve::safe_solve_warp(Hr, k, lane, gr, pert, chol_fastpath, known_pd, W);
        if (act) {
            if (MODE == 0) {
                out[b * k + lane] = T(x);
            } else {
                double fn = f - x;
                if (non_negative && fn < 0.0) fn = 0.0;
                out[b * k + lane] = T(fn);
            }
        }
    }
}

}  // namespace

template <typename T, int MODE>
bool safe_solve_small(pycmf_ctx* ctx, int64_t batch, int64_t k, const T* H, int64_t h_stride, const T* g, T* out,
                      double l1, double l2, double l2_diag, double pert, bool non_negative, double h_scale, bool known_pd) {
    if (k > KS || batch < 1) return false;
    Timed timer(ctx, "safe_solve");
    if (batch <= 2 * ctx->num_sms || (h_stride == 0 && MODE == 0)) {
        // few matrices (or one matrix with many right-hand sides): CTA-cooperative factorisation
        const bool shared = h_stride == 0 && MODE == 0;
        const int64_t nb = shared ? 1 : batch;
        const int nrhs = shared ? int(batch) : 1;
        safe_solve_cta_kernel<T, MODE><<<(unsigned)nb, dim3(32, 32), 0, ctx->stream>>>(
            nb, int(k), H, h_stride, g, out, nrhs, l1, l2, l2_diag, pert, non_negative, ctx->chol_fastpath != 0, h_scale,
            known_pd);
        PYCMF_LAUNCH_CHECK(ctx);
        return true;
    }
    int64_t grid = std::min<int64_t>(ceil_div(batch, WARPS), int64_t(16) * ctx->num_sms);
    safe_solve_small_kernel<T, MODE><<<(unsigned)grid, WARPS * 32, 0, ctx->stream>>>(
        batch, int(k), H, h_stride, g, out, l1, l2, l2_diag, pert, non_negative, ctx->chol_fastpath != 0, h_scale, known_pd);
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

template <typename T>
bool newton_finish_small(pycmf_ctx* ctx, int64_t rows, int64_t l, int64_t k, T* F, const T* Z, const T* Y, int64_t ldy,
                         int y_link, double wy, const T* gx, const T* Hx, bool hx_per_row, double l1, double l2,
                         double l2_diag, double pert, bool non_negative) {
    if (k > KS || l > LMAX || l < 1 || rows < 1) return false;
    size_t smem = sizeof(double) * size_t(WARPS) * KS * WLD + sizeof(T) * (size_t(l) * (k + 1) + KS * KS);
    if (smem > size_t(ctx->max_smem_optin)) return false;
    // Definiteness shortcut: H_j = Hx(_j) + wy Z^T D Z + l2 I with the label term PSD when wy >= 0.
    //   l2 >= pert                      -> every H_j has lambda_min >= pert (pd_mode 2, if Hx is PSD: weights >= 0)
    //   shared Hx: test Hx + l2 I once  -> verdict read by every row (pd_mode 1)
    int pd_mode = 0;
    int* flag = nullptr;
    if (wy >= 0.0 && ctx->chol_fastpath) {
        if (!hx_per_row) {
            flag = static_cast<int*>(scratch(ctx, 2, 256));
            pd_flag_kernel<T><<<1, dim3(32, 32), 0, ctx->stream>>>(int(k), Hx, 1.0, l2_diag, pert, flag);
            PYCMF_LAUNCH_CHECK(ctx);
            pd_mode = 1;
        } else if (l2_diag >= pert) {
            pd_mode = 2;
        }
    }
    auto kern = newton_finish_small_kernel<T>;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t grid = std::min<int64_t>(ceil_div(rows, WARPS), int64_t(16) * ctx->num_sms);
    Timed timer(ctx, "newton_finish_small");
    kern<<<(unsigned)grid, WARPS * 32, smem, ctx->stream>>>(rows, int(l), int(k), F, Z, Y, ldy, y_link, T(wy), gx, Hx,
                                                           hx_per_row ? k * k : 0, l1, l2, l2_diag, pert, non_negative,
                                                           ctx->chol_fastpath != 0, pd_mode, flag);
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

template bool safe_solve_small<float, 1>(pycmf_ctx*, int64_t, int64_t, const float*, int64_t, const float*, float*, double,
                                         double, double, double, bool, double, bool);
template bool safe_solve_small<double, 0>(pycmf_ctx*, int64_t, int64_t, const double*, int64_t, const double*, double*,
                                          double, double, double, double, bool, double, bool);
template bool safe_solve_small<double, 1>(pycmf_ctx*, int64_t, int64_t, const double*, int64_t, const double*, double*,
                                          double, double, double, double, bool, double, bool);
template bool newton_finish_small<float>(pycmf_ctx*, int64_t, int64_t, int64_t, float*, const float*, const float*,
                                         int64_t, int, double, const float*, const float*, bool, double, double, double,
                                         double, bool);
template bool newton_finish_small<double>(pycmf_ctx*, int64_t, int64_t, int64_t, double*, const double*, const double*,
                                          int64_t, int, double, const double*, const double*, bool, double, double,
                                          double, double, bool);

}  // namespace pycmf

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

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
newton_finish_small_kernel(int64_t rows, int l, int k, T* __restrict__ F, const T* __restrict__ Z,
                           const T* __restrict__ Y, int64_t ldy, int y_link, double wy,
                           const T* __restrict__ gx, const T* __restrict__ Hx, int64_t hx_stride,
                           double l1, double l2, double l2_diag, double pert, bool non_negative, bool chol_fastpath) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Zs = reinterpret_cast<double*>(smem_raw);            // l x (k + 1)
    double* Hs = Zs + size_t(l) * (k + 1);                       // k x k   (shared Hessian part, if hx_stride == 0)
    double* Wall = Hs + KS * KS;                                 // WARPS x KS x WLD (Jacobi fallback tiles)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kz = k + 1;
    for (int e = threadIdx.x; e < l * k; e += blockDim.x) Zs[(e / k) * kz + (e % k)] = double(Z[e]);
    if (hx_stride == 0)
        for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
            int r = e / k, c = e % k;
            Hs[e] = double(Hx[(r > c ? r : c) * k + (r > c ? c : r)]);   // lower triangle, like eigh
        }
    __syncthreads();
    double* W = Wall + warp * (KS * WLD);
    const bool act = lane < k;
    for (int64_t row = int64_t(blockIdx.x) * WARPS + warp; row < rows; row += int64_t(gridDim.x) * WARPS) {
        const double v = act ? double(F[row * k + lane]) : 0.0;
        double g = act ? double(gx[row * k + lane]) : 0.0;
        // ---- estimates for the labels c2 = lane + 32 t
        double res[LMAX / 32], wgt[LMAX / 32];
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            const int c2 = lane + 32 * t;
            if (32 * t >= l) { res[t] = 0.0; wgt[t] = 0.0; continue; }
            double d = 0.0;
            for (int a = 0; a < k; a++) d = fma(shfl_d(v, a), (c2 < l) ? Zs[c2 * kz + a] : 0.0, d);
            double est = d, fp = 1.0;
            if (y_link == PYCMF_LOGIT) { est = 1.0 / (1.0 + exp(-d)); fp = est * (1.0 - est); }
            const double y = (c2 < l) ? double(Y[row * ldy + c2]) : 0.0;
            res[t] = (c2 < l) ? wy * (est - y) : 0.0;
            wgt[t] = (c2 < l) ? wy * fp : 0.0;
        }
        // ---- row `lane` of the Hessian in registers
        double Wr[KS];
#pragma unroll
        for (int c = 0; c < KS; c++) {
            double h = 0.0;
            if (act && c < k) {
                if (hx_stride == 0) h = Hs[lane * k + c];
                else {
                    const int hi = lane > c ? lane : c, lo = lane > c ? c : lane;
                    h = double(Hx[row * hx_stride + hi * k + lo]);
                }
                if (c == lane) h += l2_diag;
            }
            Wr[c] = h;
        }
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            const int lim = min(32, l - 32 * t);
            for (int cc = 0; cc < lim; cc++) {
                const int c2 = 32 * t + cc;
                const double r = shfl_d(res[t], cc), w = shfl_d(wgt[t], cc);
                const double za = act ? Zs[c2 * kz + lane] : 0.0;
                g = fma(r, za, g);
                const double wza = w * za;
#pragma unroll
                for (int c = 0; c < KS; c++)
                    if (c < k) Wr[c] = fma(wza, Zs[c2 * kz + c], Wr[c]);
            }
        }
        const double sgn = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0);
        const double gfull = act ? g + l1 * sgn + l2 * v : 0.0;
        // ---- solve (registers; Jacobi fallback in the per-warp tile)
        const double x = wsolve::safe_solve_warp(Wr, k, lane, gfull, pert, chol_fastpath, W);
        if (act) {
            double f = v - x;
            if (non_negative && f < 0.0) f = 0.0;
            F[row * k + lane] = T(f);
        }
    }
}

// Warp-per-matrix clamped solve for k <= 32.  MODE 0: x_b = S(scale H_b + diag I) g_b.  MODE 1: Newton row update.
template <typename T, int MODE>
__global__ void __launch_bounds__(WARPS * 32)
safe_solve_small_kernel(int64_t batch, int k, const T* __restrict__ H, int64_t h_stride, const T* __restrict__ g,
                        T* __restrict__ out, double l1, double l2, double l2_diag, double pert, bool non_negative,
                        bool chol_fastpath, double h_scale) {
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
        const double x = wsolve::safe_solve_warp(Hr, k, lane, gr, pert, chol_fastpath, W);
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
                      double l1, double l2, double l2_diag, double pert, bool non_negative, double h_scale) {
    if (k > KS || batch < 1) return false;
    int64_t grid = std::min<int64_t>(ceil_div(batch, WARPS), int64_t(16) * ctx->num_sms);
    Timed timer(ctx, "safe_solve");
    safe_solve_small_kernel<T, MODE><<<(unsigned)grid, WARPS * 32, 0, ctx->stream>>>(
        batch, int(k), H, h_stride, g, out, l1, l2, l2_diag, pert, non_negative, ctx->chol_fastpath != 0, h_scale);
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

template <typename T>
bool newton_finish_small(pycmf_ctx* ctx, int64_t rows, int64_t l, int64_t k, T* F, const T* Z, const T* Y, int64_t ldy,
                         int y_link, double wy, const T* gx, const T* Hx, bool hx_per_row, double l1, double l2,
                         double l2_diag, double pert, bool non_negative) {
    if (k > KS || l > LMAX || l < 1 || rows < 1) return false;
    size_t smem = sizeof(double) * (size_t(l) * (k + 1) + KS * KS + size_t(WARPS) * KS * WLD);
    if (smem > size_t(ctx->max_smem_optin)) return false;
    auto kern = newton_finish_small_kernel<T>;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t grid = std::min<int64_t>(ceil_div(rows, WARPS), int64_t(8) * ctx->num_sms);
    Timed timer(ctx, "newton_finish_small");
    kern<<<(unsigned)grid, WARPS * 32, smem, ctx->stream>>>(rows, int(l), int(k), F, Z, Y, ldy, y_link, wy, gx, Hx,
                                                           hx_per_row ? k * k : 0, l1, l2, l2_diag, pert, non_negative,
                                                           ctx->chol_fastpath != 0);
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

template bool safe_solve_small<float, 1>(pycmf_ctx*, int64_t, int64_t, const float*, int64_t, const float*, float*, double,
                                         double, double, double, bool, double);
template bool safe_solve_small<double, 0>(pycmf_ctx*, int64_t, int64_t, const double*, int64_t, const double*, double*,
                                          double, double, double, double, bool, double);
template bool safe_solve_small<double, 1>(pycmf_ctx*, int64_t, int64_t, const double*, int64_t, const double*, double*,
                                          double, double, double, double, bool, double);
template bool newton_finish_small<float>(pycmf_ctx*, int64_t, int64_t, int64_t, float*, const float*, const float*,
                                         int64_t, int, double, const float*, const float*, bool, double, double, double,
                                         double, bool);
template bool newton_finish_small<double>(pycmf_ctx*, int64_t, int64_t, int64_t, double*, const double*, const double*,
                                          int64_t, int, double, const double*, const double*, bool, double, double,
                                          double, double, bool);

}  // namespace pycmf

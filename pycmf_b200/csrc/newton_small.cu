// Fused finish of the Newton V update for small n_components (k <= 32): one WARP per row of V does
//   g_j = gx_j + w (f2(v_j Z^T) - Y[j,:]) Z + l1 sign(v_j) + l2 v_j
//   H_j = Hx(_j) + w Z^T diag(f2'(v_j Z^T)) Z + l2 I
//   v_j <- v_j - S(H_j) g_j ; optional clamp                           (reference cmf_solvers.py:432-486)
// entirely on chip: the row of H lives in registers (lane = Hessian row), the factorisation runs in a
// per-warp shared-memory tile with warp-synchronous Cholesky (fast path, lambda_min(H) > pert) or a
// warp-level one-sided Jacobi (eigenvalue clamp active).  Replaces five launches of the generic path
// (small fused-residual pass, axpby, Hessian broadcast, per-row Hessian, batched solve) and the
// d x k x k Hessian round trip through HBM.  All arithmetic in float64.
#include "common.cuh"

namespace pycmf {
namespace {

constexpr int KS = 32;            // max n_components
constexpr int WLD = KS + 1;       // padded leading dimension of the per-warp tile
constexpr int WARPS = 8;
constexpr int LMAX = 128;         // max rows of the small factor (labels)

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// In-place Cholesky of the lower triangle of W (k x k, ld WLD); lane == row.  Uniform return value.
__device__ bool warp_cholesky(double* W, int k, int lane, double floor) {
    for (int j = 0; j < k; j++) {
        const double piv = W[j * WLD + j];
        if (!(piv > floor)) return false;
        const double inv = 1.0 / sqrt(piv);
        if (lane >= j && lane < k) W[lane * WLD + j] *= inv;
        __syncwarp();
        if (lane > j && lane < k) {
            const double lrj = W[lane * WLD + j];
            for (int c = j + 1; c <= lane; c++) W[lane * WLD + c] -= lrj * W[c * WLD + j];
        }
        __syncwarp();
    }
    return true;
}

// Solve L L^T x = b; b is distributed (lane r holds b_r); returns x_r in lane r.
__device__ double warp_chol_solve(const double* W, int k, int lane, double b) {
    for (int j = 0; j < k; j++) {
        const double y = shfl_d(b, j) / W[j * WLD + j];
        if (lane == j) b = y;
        if (lane > j && lane < k) b -= W[lane * WLD + j] * y;
    }
    for (int j = k - 1; j >= 0; j--) {
        const double x = shfl_d(b, j) / W[j * WLD + j];
        if (lane == j) b = x;
        if (lane < j) b -= W[j * WLD + lane] * x;
    }
    return b;
}

// Eigenvalue-clamped solve by one-sided Jacobi on the columns of the symmetric W (lane == row index).
__device__ double warp_jacobi_solve(double* W, int k, int lane, double g, double pert) {
    const bool act = lane < k;
    const double tol = 1e-15, skip2 = (1e-3 * pert) * (1e-3 * pert);
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < k - 1; p++) {
            for (int q = p + 1; q < k; q++) {
                const double a = act ? W[lane * WLD + p] : 0.0, b = act ? W[lane * WLD + q] : 0.0;
                const double al = warp_sum(a * a), be = warp_sum(b * b), ga = warp_sum(a * b);
                if (ga == 0.0 || fmax(al, be) < skip2) continue;
                if (fabs(ga) <= tol * sqrt(al * be)) continue;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                if (act) {
                    W[lane * WLD + p] = c * a - s * b;
                    W[lane * WLD + q] = s * a + c * b;
                }
                rotated = true;
            }
        }
        __syncwarp();
        if (!rotated) break;
    }
    double x = g / pert;
    for (int i = 0; i < k; i++) {
        const double w = act ? W[lane * WLD + i] : 0.0;
        const double al = warp_sum(w * w), dg = warp_sum(w * (act ? g : 0.0));
        const double sigma = sqrt(al);
        if (sigma >= pert) x += (1.0 / sigma - 1.0 / pert) * dg / al * w;
    }
    return x;
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
newton_finish_small_kernel(int64_t rows, int l, int k, T* __restrict__ F, const T* __restrict__ Z,
                           const T* __restrict__ Y, int64_t ldy, int y_link, double wy,
                           const T* __restrict__ gx, const T* __restrict__ Hx, int64_t hx_stride,
                           double l1, double l2, double l2_diag, double pert, bool non_negative, bool chol_fastpath) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Zs = reinterpret_cast<double*>(smem_raw);            // l x (k + 1)
    double* Hs = Zs + size_t(l) * (k + 1);                       // k x k   (shared Hessian part, if hx_stride == 0)
    double* Wall = Hs + KS * KS;                                 // WARPS x KS x WLD
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
        // ---- solve
        double x;
        bool done = false;
        if (chol_fastpath) {
            double tr = 0.0;
#pragma unroll
            for (int c = 0; c < KS; c++) {
                if (c < k && act) W[lane * WLD + c] = Wr[c] - (c == lane ? pert : 0.0);
                if (c == lane && act) tr = fabs(Wr[c] - pert);
            }
            tr = warp_sum(tr);
            __syncwarp();
            bool ok = warp_cholesky(W, k, lane, 1e-13 * (tr + pert));
            __syncwarp();
            if (ok) {
#pragma unroll
                for (int c = 0; c < KS; c++)
                    if (c < k && act) W[lane * WLD + c] = Wr[c];
                __syncwarp();
                ok = warp_cholesky(W, k, lane, 0.0);
                if (ok) {
                    x = warp_chol_solve(W, k, lane, gfull);
                    done = true;
                }
            }
            __syncwarp();
        }
        if (!done) {
            // symmetric tile from the lower triangle: W[r][c] = H[max][min]
#pragma unroll
            for (int c = 0; c < KS; c++)
                if (c < k && act && c <= lane) W[lane * WLD + c] = Wr[c];
            __syncwarp();
            for (int c = lane + 1; c < k; c++)
                if (act) W[lane * WLD + c] = W[c * WLD + lane];
            __syncwarp();
            x = warp_jacobi_solve(W, k, lane, gfull, pert);
            __syncwarp();
        }
        if (act) {
            double f = v - x;
            if (non_negative && f < 0.0) f = 0.0;
            F[row * k + lane] = T(f);
        }
    }
}

}  // namespace

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

template bool newton_finish_small<float>(pycmf_ctx*, int64_t, int64_t, int64_t, float*, const float*, const float*,
                                         int64_t, int, double, const float*, const float*, bool, double, double, double,
                                         double, bool);
template bool newton_finish_small<double>(pycmf_ctx*, int64_t, int64_t, int64_t, double*, const double*, const double*,
                                          int64_t, int, double, const double*, const double*, bool, double, double,
                                          double, double, bool);

}  // namespace pycmf

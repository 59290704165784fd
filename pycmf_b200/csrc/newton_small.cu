// Small-n_components (k <= 32) Newton kernels: one WARP per row / per matrix, register-resident Cholesky (warp_solve.cuh).
//
// newton_finish_small : fused finish of the V update.  One warp per row of V does
//   g_j = gx_j + w (f2(v_j Z^T) - Y[j,:]) Z + l1 sign(v_j) + l2 v_j
//   H_j = Hx(_j) + w Z^T diag(f2'(v_j Z^T)) Z + l2 I
//   v_j <- v_j - S(H_j) g_j ; optional clamp                           (reference cmf_solvers.py:432-486)
// entirely on chip: the row of H is built in registers (lane = Hessian row) from 128-bit broadcast reads of Z,
// factorised in float64 and applied.  Replaces five launches of the generic path (small fused-residual pass, axpby,
// Hessian broadcast, per-row Hessian, batched solve) and the d x k x k Hessian round trip through HBM.
//
// safe_solve_small    : x_b = S(scale H_b + diag I) g_b (MODE 0) or the Newton row update (MODE 1), one warp per
//                       right-hand side (a shared matrix is factorised redundantly by each warp: ~5 us of latency,
//                       cheaper than any cross-warp hand-off at k <= 32).
// Templates on KT = 8 / 16 / 32 (k padded with identity rows) so that all register arrays have compile-time indices.
#include "warp_solve.cuh"

namespace pycmf {
namespace {

using wsolve::KS;
using wsolve::WLD;
using wsolve::TILE;
using wsolve::shfl_d;
constexpr int WARPS = 4;
constexpr int LMAX = 128;         // max rows of the small factor (labels)

// four consecutive elements of a shared-memory row (16-byte aligned for float, two 16-byte loads for double)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    const double2 t0 = *reinterpret_cast<const double2*>(p), t1 = *reinterpret_cast<const double2*>(p + 2);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
}

// pd_mode: 0 = test every row (Cholesky of H - pert I), 1 = read the shared verdict from *pd_flag, 2 = known PD
template <typename T, int KT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
newton_finish_small_kernel(int64_t rows, int l, int k, T* __restrict__ F, const T* __restrict__ Z,
                           const T* __restrict__ Y, int64_t ldy, int y_link, T wy,
                           const T* __restrict__ gx, const T* __restrict__ Hx, const double* __restrict__ Hx64,
                           int64_t hx_stride,
                           double l1, double l2, double l2_diag, double pert, bool non_negative, bool chol_fastpath,
                           int pd_mode, const int* __restrict__ pd_flag) {
    constexpr int ZLD = KT + 4;       // row stride of Z in shared memory: 16-byte aligned rows, conflict-free columns
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Wall = reinterpret_cast<double*>(smem_raw);          // WARPS solve tiles
    constexpr int HLD = KT + 1;       // odd pitch: lane r reads row r, so the 32 lanes of a warp hit 32 different banks
    double* Hs = Wall + WARPS * TILE;                            // KT x HLD float64 (shared Hessian part, if hx_stride == 0)
    T* Zs = reinterpret_cast<T*>(Hs + KT * HLD);                 // l x ZLD (columns >= k zero); KT (KT + 1) is even: 16-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < l * ZLD + 32; e += blockDim.x) {
        const int r = e / ZLD, c = e % ZLD;
        Zs[e] = (r < l && c < k) ? Z[r * k + c] : T(0);
    }
    if (hx_stride == 0)
        for (int e = threadIdx.x; e < KT * KT; e += blockDim.x) {
            const int r = e / KT, c = e % KT;
            Hs[r * HLD + c] = (r < k && c < k) ? Hx64[(r > c ? r : c) * k + (r > c ? c : r)] : 0.0;   // lower triangle, like eigh
        }
    __syncthreads();
    double* W = Wall + warp * TILE;
    const bool act = lane < k;
    const bool known_pd = pd_mode == 2 || (pd_mode == 1 && *pd_flag != 0);
    for (int64_t row = int64_t(blockIdx.x) * WARPS + warp; row < rows; row += int64_t(gridDim.x) * WARPS) {
        const T v = act ? F[row * k + lane] : T(0);
        T g = act ? gx[row * k + lane] : T(0);
        // ---- estimates for the labels c2 = lane + 32 t : d = v . z_c2
        T res[LMAX / 32], wgt[LMAX / 32];
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            res[t] = T(0);
            wgt[t] = T(0);
            if (32 * t >= l) continue;
            const int c2 = lane + 32 * t;
            const T* zr = Zs + (c2 < l ? c2 : 0) * ZLD;
            T d = T(0);
#pragma unroll
            for (int a0 = 0; a0 < KT; a0 += 4) {
                T z4[4];
                load4(zr + a0, z4);
#pragma unroll
                for (int u = 0; u < 4; u++) d = fma(T(__shfl_sync(0xffffffffu, v, a0 + u)), z4[u], d);
            }
            T est = d, fp = T(1);
            if (y_link == PYCMF_LOGIT) { est = sigmoid_<T>(d); fp = est * (T(1) - est); }
            if (c2 < l) {
                res[t] = wy * (est - Y[row * ldy + c2]);
                wgt[t] = wy * fp;
            }
        }
        // ---- row `lane` of the Hessian in registers (compute dtype)
        T Wr[KT];
        if (hx_stride == 0) {
            // the shared part stays in float64 shared memory and is added inside the factorisation: rounding alpha U^T U
            // to fp32 cost 8e-3 on V in C2's first iteration (6e-8 x cond(H)); only the label part is built in T
#pragma unroll
            for (int c = 0; c < KT; c++) Wr[c] = T(0);
        } else {
#pragma unroll
            for (int c = 0; c < KT; c++) {
                const int hi = lane > c ? lane : c, lo = lane > c ? c : lane;
                Wr[c] = (act && c < k) ? Hx[row * hx_stride + hi * k + lo] : T(0);
            }
        }
#pragma unroll
        for (int t = 0; t < LMAX / 32; t++) {
            const int lim = min(32, l - 32 * t);
            for (int cc = 0; cc < lim; cc++) {
                const T* zr = Zs + (32 * t + cc) * ZLD;
                const T r = __shfl_sync(0xffffffffu, res[t], cc), w = __shfl_sync(0xffffffffu, wgt[t], cc);
                const T za = zr[lane < KT ? lane : 0];
                g = fma(r, za, g);
                const T wza = w * za;
#pragma unroll
                for (int c = 0; c < KT; c += 4) {
                    T z4[4];
                    load4(zr + c, z4);
#pragma unroll
                    for (int u = 0; u < 4; u++) Wr[c + u] = fma(wza, z4[u], Wr[c + u]);
                }
            }
        }
        const double vd = double(v);
        const double sgn = vd > 0.0 ? 1.0 : (vd < 0.0 ? -1.0 : 0.0);
        const double gfull = act ? double(g) + l1 * sgn + l2 * vd : 0.0;
        // ---- the clamped solve in float64 (l2 on the diagonal)
        double a[KT], dinv;
        const bool fac = wsolve::safe_factor_warp<KT, T>(Wr, l2_diag, k, lane, pert, chol_fastpath, known_pd, W, a, &dinv,
                                                         hx_stride == 0 ? Hs + (lane < KT ? lane : 0) * HLD : nullptr);
        const double x = wsolve::safe_apply_warp<KT>(fac, a, dinv, k, lane, gfull, pert, W);
        __syncwarp();
        if (act) {
            double f = vd - x;
            if (non_negative && f < 0.0) f = 0.0;
            F[row * k + lane] = T(f);
        }
    }
}

// *flag = 1 iff scale * H + (diag - pert) I is positive definite (H is k x k, lower triangle used).  One warp.
template <typename T, int KT>
__global__ void __launch_bounds__(32)
pd_flag_kernel(int k, const T* __restrict__ H, double scale, double diag, double pert, int* flag) {
    __shared__ __align__(16) double W[TILE];
    const int lane = threadIdx.x;
    const bool act = lane < k;
    double a[KT], tr = 0.0;
#pragma unroll
    for (int c = 0; c < KT; c++) {
        a[c] = (act && c <= lane) ? scale * double(H[lane * k + c]) + (c == lane ? diag - pert : 0.0)
                                  : (c == lane ? 1.0 : 0.0);
        if (c == lane && act) tr = fabs(a[c]);
    }
    tr = warp_sum(tr);
    double dinv;
    const bool ok = wsolve::chol_reg<KT>(a, lane, 1e-13 * (tr + pert), W, &dinv);
    if (lane == 0) *flag = ok ? 1 : 0;
}

// Warp per right-hand side.  h_stride == 0: one shared matrix, rhs / outputs indexed by b.
// MODE 0: x_b = S(scale H_b + diag I) g_b.  MODE 1: Newton row update of out[b].
template <typename T, int MODE, int KT>
__global__ void __launch_bounds__(WARPS * 32)
safe_solve_small_kernel(int64_t batch, int k, const T* __restrict__ H, int64_t h_stride, const T* __restrict__ g,
                        T* __restrict__ out, double l1, double l2, double l2_diag, double pert, bool non_negative,
                        bool chol_fastpath, double h_scale, bool known_pd) {
    __shared__ __align__(16) double Wall[WARPS * TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // 1 or WARPS warps per CTA
    double* W = Wall + warp * TILE;
    const bool act = lane < k;
    double a[KT], dinv = 1.0;
    bool fac = false, have = false;
    for (int64_t b = int64_t(blockIdx.x) * nwarps + warp; b < batch; b += int64_t(gridDim.x) * nwarps) {
        if (!have || h_stride != 0) {
            const T* Hb = H + b * h_stride;
            double Hr[KT];
#pragma unroll
            for (int c = 0; c < KT; c++)
                Hr[c] = (act && c <= lane) ? h_scale * double(Hb[lane * k + c]) : 0.0;
            fac = wsolve::safe_factor_warp<KT, double>(Hr, l2_diag, k, lane, pert, chol_fastpath, known_pd, W, a, &dinv);
            have = true;
        }
        double gr = act ? double(g[b * k + lane]) : 0.0;
        double f = 0.0;
        if (MODE == 1 && act) {
            f = double(out[b * k + lane]);
            gr += l1 * (f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0)) + l2 * f;
        }
        const double x = wsolve::safe_apply_warp<KT>(fac, a, dinv, k, lane, gr, pert, W);
        __syncwarp();
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
    // few right-hand sides: one warp per CTA so that they spread over the SMs (pure latency otherwise)
    const int warps = batch <= 2 * ctx->num_sms ? 1 : WARPS;
    const int64_t grid = std::min<int64_t>(ceil_div(batch, warps), int64_t(16) * ctx->num_sms);
    const bool fast = ctx->chol_fastpath != 0;
#define LAUNCH(KT)                                                                                                      \
    safe_solve_small_kernel<T, MODE, KT><<<(unsigned)grid, warps * 32, 0, ctx->stream>>>(                               \
        batch, int(k), H, h_stride, g, out, l1, l2, l2_diag, pert, non_negative, fast, h_scale, known_pd)
    const int kt = wsolve::pick_kt(int(k));
    if (kt == 8) LAUNCH(8);
    else if (kt == 16) LAUNCH(16);
    else LAUNCH(32);
#undef LAUNCH
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

template <typename T>
bool newton_finish_small(pycmf_ctx* ctx, int64_t rows, int64_t l, int64_t k, T* F, const T* Z, const T* Y, int64_t ldy,
                         int y_link, double wy, const T* gx, const void* Hx_any, bool hx_per_row, double l1, double l2,
                         double l2_diag, double pert, bool non_negative) {
    if (k > KS || l > LMAX || l < 1 || rows < 1) return false;
    const T* Hx = hx_per_row ? static_cast<const T*>(Hx_any) : nullptr;                 // per-row Hessians: compute dtype
    const double* Hx64 = hx_per_row ? nullptr : static_cast<const double*>(Hx_any);     // shared Hessian: float64
    const int kt = wsolve::pick_kt(int(k));
    const size_t smem = sizeof(double) * (size_t(WARPS) * TILE + size_t(kt) * (kt + 1)) + sizeof(T) * (size_t(l) * (kt + 4) + 32);
    if (smem > size_t(ctx->max_smem_optin)) return false;
    // Definiteness shortcut: H_j = Hx(_j) + wy Z^T D Z + l2 I with the label term PSD when wy >= 0.
    //   l2 >= pert                      -> every H_j has lambda_min >= pert (pd_mode 2, if Hx is PSD: weights >= 0)
    //   shared Hx: test Hx + l2 I once  -> verdict read by every row (pd_mode 1)
    int pd_mode = 0;
    int* flag = nullptr;
    if (wy >= 0.0 && ctx->chol_fastpath) {
        if (!hx_per_row) {
            flag = static_cast<int*>(scratch(ctx, 2, 256));
            if (kt == 8) pd_flag_kernel<double, 8><<<1, 32, 0, ctx->stream>>>(int(k), Hx64, 1.0, l2_diag, pert, flag);
            else if (kt == 16) pd_flag_kernel<double, 16><<<1, 32, 0, ctx->stream>>>(int(k), Hx64, 1.0, l2_diag, pert, flag);
            else pd_flag_kernel<double, 32><<<1, 32, 0, ctx->stream>>>(int(k), Hx64, 1.0, l2_diag, pert, flag);
            PYCMF_LAUNCH_CHECK(ctx);
            pd_mode = 1;
        } else if (l2_diag >= pert) {
            pd_mode = 2;
        }
    }
    // two to three resident CTAs per SM (registers); every warp walks its rows with a grid stride, so the prologue
    // (Z and the shared Hessian into shared memory) is paid once per CTA, not once per four rows
    const int minb = ctx->finish_minblocks >= 4 ? 4 : (ctx->finish_minblocks == 3 ? 3 : 2);
    const int64_t grid = std::min<int64_t>(ceil_div(rows, WARPS), int64_t(minb) * ctx->num_sms);
    Timed timer(ctx, "newton_finish_small");
#define LAUNCH(KT)                                                                                                      \
    do {                                                                                                                \
        auto kern = minb == 4 ? newton_finish_small_kernel<T, KT, 4>                                                    \
                              : (minb == 3 ? newton_finish_small_kernel<T, KT, 3> : newton_finish_small_kernel<T, KT, 2>); \
        PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));                 \
        kern<<<(unsigned)grid, WARPS * 32, smem, ctx->stream>>>(rows, int(l), int(k), F, Z, Y, ldy, y_link, T(wy), gx,  \
                                                               Hx, Hx64, hx_per_row ? k * k : 0, l1, l2, l2_diag, pert, \
                                                               non_negative, ctx->chol_fastpath != 0, pd_mode, flag);  \
    } while (0)
    if (kt == 8) LAUNCH(8);
    else if (kt == 16) LAUNCH(16);
    else LAUNCH(32);
#undef LAUNCH
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
                                         int64_t, int, double, const float*, const void*, bool, double, double, double,
                                         double, bool);
template bool newton_finish_small<double>(pycmf_ctx*, int64_t, int64_t, int64_t, double*, const double*, const double*,
                                          int64_t, int, double, const double*, const void*, bool, double, double,
                                          double, double, bool);

}  // namespace pycmf

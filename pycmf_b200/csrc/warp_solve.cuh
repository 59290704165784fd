// Warp-level eigenvalue-clamped solve for small symmetric systems (k <= 32), float64.
//   x = S(H) g,  S(H) = Q diag(1 / max(|lambda|, pert)) Q^T          (reference _safe_invert, cmf_solvers.py:346-356)
// One warp per matrix; the matrix lives in a per-warp shared-memory tile (lane == row, leading dimension 33 so a
// column access is bank-conflict free).  Fast path: Cholesky of H - pert I succeeds  <=>  lambda_min(H) > pert  <=>
// the clamp is inactive and S(H) = H^-1, solved with a Cholesky of H.  When the caller can prove lambda_min >= pert
// (Hessian = PSD term + l2 I with l2 >= pert, or a shared lower bound tested once) the test factorisation is skipped.
// Otherwise: one-sided Jacobi in the same tile.  Runtime loops only (small code: the I-cache matters here).
#pragma once
#include "common.cuh"

namespace pycmf {
namespace wsolve {

constexpr int KS = 32;
constexpr int WLD = KS + 1;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// In-place Cholesky of the lower triangle of the tile.  On success W[r][c] (c < r) = L[r][c] and the DIAGONAL holds
// 1 / L[r][r].  Uniform return value (false: a pivot was <= floor).
__device__ __forceinline__ bool chol_tile(double* W, int k, int lane, double floor) {
    for (int j = 0; j < k; j++) {
        const double piv = W[j * WLD + j];
        if (!(piv > floor)) return false;
        const double inv = rsqrt(piv);
        double lrj = 0.0;
        if (lane > j && lane < k) {
            lrj = W[lane * WLD + j] * inv;
            W[lane * WLD + j] = lrj;
        }
        if (lane == j) W[j * WLD + j] = inv;
        __syncwarp();
        // trailing update of this lane's row: W[r][c] -= L[r][j] L[c][j] for j < c <= r (4 independent columns a time)
        for (int c0 = j + 1; c0 < k; c0 += 4) {
            double l4[4], w4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = c0 + u;
                const bool on = c < k && c <= lane && lane < k;
                // L[c][j] for c == lane is lrj itself; for c < lane it is row c, column j (broadcast read)
                l4[u] = on ? (c == lane ? lrj : W[c * WLD + j]) : 0.0;
                w4[u] = on ? W[lane * WLD + c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = c0 + u;
                if (c < k && c <= lane && lane < k) W[lane * WLD + c] = fma(-lrj, l4[u], w4[u]);
            }
        }
        __syncwarp();
    }
    return true;
}

// Solve L L^T x = b with the factor left by chol_tile; lane r holds b_r on entry and x_r on return.
__device__ __forceinline__ double chol_solve_tile(const double* W, int k, int lane, double b) {
    for (int j = 0; j < k; j++) {
        const double y = shfl_d(b, j) * W[j * WLD + j];
        if (lane == j) b = y;
        if (lane > j && lane < k) b = fma(-W[lane * WLD + j], y, b);
    }
    for (int j = k - 1; j >= 0; j--) {
        const double x = shfl_d(b, j) * W[j * WLD + j];
        if (lane == j) b = x;
        if (lane < j) b = fma(-W[j * WLD + lane], x, b);
    }
    return b;
}

// One-sided (Hestenes) Jacobi sweeps on the columns of the symmetric tile W (lane == row), executed by ONE warp.
// On exit the columns are mutually orthogonal: W = Q diag(lambda) up to column order / sign.
__device__ __forceinline__ void jacobi_sweeps_tile(double* W, int k, int lane, double pert) {
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
}

// x = S(H) g from the orthogonalised tile:  x = g/p + sum_{sigma_i >= p} (1/sigma_i - 1/p) (w_i . g) / sigma_i^2  w_i
__device__ __forceinline__ double jacobi_apply_tile(const double* W, int k, int lane, double g, double pert) {
    const bool act = lane < k;
    double x = g / pert;
    for (int i = 0; i < k; i++) {
        const double w = act ? W[lane * WLD + i] : 0.0;
        const double al = warp_sum(w * w), dg = warp_sum(w * (act ? g : 0.0));
        const double sigma = sqrt(al);
        if (sigma >= pert) x += (1.0 / sigma - 1.0 / pert) * dg / al * w;
    }
    return x;
}

__device__ __forceinline__ double jacobi_solve_tile(double* W, int k, int lane, double g, double pert) {
    jacobi_sweeps_tile(W, k, lane, pert);
    return jacobi_apply_tile(W, k, lane, g, pert);
}

// Writes row `lane` (lower part c <= lane) of the matrix into the tile, shifting the diagonal by `shift`.
template <typename R>
__device__ __forceinline__ void store_row_lower(double* W, const R (&H_row)[KS], int k, int lane, double shift) {
#pragma unroll
    for (int c = 0; c < KS; c++)
        if (c < k && c <= lane && lane < k) W[lane * WLD + c] = double(H_row[c]) + (c == lane ? shift : 0.0);
}

// The full clamped solve.  H_row: row `lane` of the symmetric matrix in registers (entries c <= lane are used, like
// eigh(lower=True)); g: this lane's right-hand-side entry; W: per-warp tile of KS * WLD doubles.
// known_pd: the caller guarantees lambda_min(H) >= pert (skips the test factorisation).
template <typename R>
__device__ __forceinline__ double safe_solve_warp(const R (&H_row)[KS], int k, int lane, double g, double pert,
                                                  bool chol_fastpath, bool known_pd, double* W) {
    const bool act = lane < k;
    if (chol_fastpath) {
        bool ok = known_pd;
        if (!ok) {
            double tr = 0.0;
#pragma unroll
            for (int c = 0; c < KS; c++)
                if (c == lane && act) tr = fabs(double(H_row[c]) - pert);
            tr = warp_sum(tr);
            store_row_lower(W, H_row, k, lane, -pert);
            __syncwarp();
            ok = chol_tile(W, k, lane, 1e-13 * (tr + pert));
            __syncwarp();
        }
        if (ok) {
            store_row_lower(W, H_row, k, lane, 0.0);
            __syncwarp();
            if (chol_tile(W, k, lane, 0.0)) {
                const double x = chol_solve_tile(W, k, lane, act ? g : 0.0);
                __syncwarp();
                return x;
            }
            __syncwarp();
        }
    }
    // eigenvalue clamp active (or fast path disabled): Jacobi on the symmetric tile built from the lower triangle
    store_row_lower(W, H_row, k, lane, 0.0);
    __syncwarp();
    for (int c = lane + 1; c < k; c++)
        if (act) W[lane * WLD + c] = W[c * WLD + lane];
    __syncwarp();
    const double x = jacobi_solve_tile(W, k, lane, act ? g : 0.0, pert);
    __syncwarp();
    return x;
}

}  // namespace wsolve
}  // namespace pycmf

// Warp-level eigenvalue-clamped solve for small symmetric systems (k <= 32), all in float64.
//   x = S(H) g,  S(H) = Q diag(1 / max(|lambda|, pert)) Q^T          (reference _safe_invert, cmf_solvers.py:346-356)
// Lane r owns row r of H in REGISTERS (double A[32]); the Cholesky factorisation, both triangular solves and
// the definiteness test (Cholesky of H - pert I succeeds <=> lambda_min(H) > pert <=> the clamp is inactive and
// S(H) = H^-1) exchange data by warp shuffles only -- no shared-memory latency chains, no block barriers.
// If the test fails the warp falls back to a one-sided Jacobi in a per-warp shared-memory tile.
#pragma once
#include "common.cuh"

namespace pycmf {
namespace wsolve {

constexpr int KS = 32;
constexpr int WLD = KS + 1;   // padded leading dimension of the Jacobi tile (doubles)

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// In-register Cholesky. On success A[c] (c <= lane) holds L[lane][c] and *dinv = 1 / L[lane][lane].
// Uniform return value (false: a pivot was <= floor).
__device__ __forceinline__ bool chol_reg(double (&A)[KS], int k, int lane, double floor, double* dinv) {
    double my_inv = 0.0;
#pragma unroll
    for (int j = 0; j < KS; j++) {
        if (j < k) {
            const double piv = shfl_d(A[j], j);
            if (!(piv > floor)) return false;
            const double inv = rsqrt(piv);
            const double lrj = A[j] * inv;
            A[j] = lrj;
            if (lane == j) my_inv = inv;
#pragma unroll
            for (int c = j + 1; c < KS; c++) {
                if (c < k) {
                    const double lcj = shfl_d(lrj, c);
                    A[c] = fma(-lrj, lcj, A[c]);
                }
            }
        }
    }
    *dinv = my_inv;
    return true;
}

// Solve L L^T x = b with L in registers (see chol_reg); lane r holds b_r on entry and x_r on return.
__device__ __forceinline__ double chol_solve_reg(const double (&A)[KS], int k, int lane, double dinv, double b) {
#pragma unroll
    for (int j = 0; j < KS; j++) {
        if (j < k) {
            const double y = shfl_d(b * dinv, j);
            if (lane == j) b = y;
            if (lane > j) b = fma(-A[j], y, b);
        }
    }
#pragma unroll
    for (int j = KS - 1; j >= 0; j--) {
        if (j < k) {
            const double x = shfl_d(b * dinv, j);
            if (lane == j) b = x;
            // b_r -= L[j][r] x for r < j : L[j][r] is register r of lane j
#pragma unroll
            for (int r = 0; r < KS; r++) {
                if (r < j) {
                    const double ljr = shfl_d(A[r], j);
                    if (lane == r) b = fma(-ljr, x, b);
                }
            }
        }
    }
    return b;
}

// One-sided (Hestenes) Jacobi on the columns of the symmetric tile W (lane == row).  Returns x_lane.
__device__ __forceinline__ double jacobi_solve_tile(double* W, int k, int lane, double g, double pert) {
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

// The full clamped solve.  H_row: row `lane` of the symmetric matrix (entries c <= lane are used, like
// eigh(lower=True)); g: this lane's right-hand-side entry; W: per-warp tile of KS * WLD doubles (fallback only).
__device__ __forceinline__ double safe_solve_warp(const double (&H_row)[KS], int k, int lane, double g, double pert,
                                                  bool chol_fastpath, double* W) {
    const bool act = lane < k;
    if (chol_fastpath) {
        double A[KS];
        double tr = 0.0;
#pragma unroll
        for (int c = 0; c < KS; c++) {
            A[c] = H_row[c] - ((c == lane) ? pert : 0.0);
            if (c == lane && act) tr = fabs(A[c]);
        }
        tr = warp_sum(tr);
        double dinv;
        if (chol_reg(A, k, lane, 1e-13 * (tr + pert), &dinv)) {
#pragma unroll
            for (int c = 0; c < KS; c++) A[c] = H_row[c];
            if (chol_reg(A, k, lane, 0.0, &dinv)) return chol_solve_reg(A, k, lane, dinv, act ? g : 0.0);
        }
    }
    // eigenvalue clamp active (or fast path disabled): Jacobi on the symmetric tile built from the lower triangle
#pragma unroll
    for (int c = 0; c < KS; c++)
        if (c < k && act && c <= lane) W[lane * WLD + c] = H_row[c];
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

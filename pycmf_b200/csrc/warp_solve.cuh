// Warp-level eigenvalue-clamped solve for small symmetric systems (k <= 32), float64.
//   x = S(H) g,  S(H) = Q diag(1 / max(|lambda|, pert)) Q^T          (reference _safe_invert, cmf_solvers.py:346-356)
// One warp per matrix.  Fast path: Cholesky of H - pert I succeeds  <=>  lambda_min(H) > pert  <=>  the clamp is
// inactive and S(H) = H^-1, solved with a Cholesky of H.  When the caller can prove lambda_min >= pert (Hessian =
// PSD term + l2 I with l2 >= pert, or a shared lower bound tested once) the test factorisation is skipped.
//
// The factorisation is register-resident: lane r holds row r of the trailing matrix in KT registers (KT = 8 / 16 / 32,
// compile-time indices), every step publishes one column of L to a per-warp shared-memory tile and the trailing
// update reads it back as 128-bit broadcasts: ~1.9 k warp instructions for KT = 32 (the previous shared-memory
// tile version with runtime loops and per-element predicates executed ~12 k, profiles/r01: 58 % of the V finish).
// Fallback (eigenvalue clamp active): one-sided Jacobi in the same tile.
#pragma once
#include "common.cuh"

namespace pycmf {
namespace wsolve {

constexpr int KS = 32;
constexpr int WLD = KS + 2;     // even: a pair of doubles at an even column is 16-byte aligned
constexpr int TILE = KS * WLD;  // doubles per warp

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__host__ __device__ constexpr int pick_kt(int k) { return k <= 8 ? 8 : (k <= 16 ? 16 : 32); }

// 1 / sqrt(x) for x in [1e-30, 1e30]: single-precision seed (MUFU.RSQ) and two Newton steps in double: branch free,
// ~10 instructions (the library rsqrt() carries a slow path for denormals / infinities).
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y = double(rsqrtf(float(x)));
    const double h = 0.5 * x;
    y = fma(y, fma(-h * y, y, 0.5), y);
    y = fma(y, fma(-h * y, y, 0.5), y);
    return y;
}

// In-place Cholesky, rows in registers.  On entry a[c] = A[lane][c] for c <= lane (entries c > lane are ignored and
// clobbered).  On success a[c] = L[lane][c] for c < lane and 0 for c >= lane, *dinv = 1 / L[lane][lane], and
// Lc[j * WLD + r] = L[r][j] for r > j, 0 for r <= j (column j of L without its diagonal).  The zeros make the
// triangular solves select-free.  Uniform return value (false: a pivot was <= floor or outside [1e-30, 1e30]).
template <int KT>
__device__ __forceinline__ bool chol_reg(double (&a)[KT], int lane, double floor, double* __restrict__ Lc, double* dinv) {
    // No early exit: a failed pivot is replaced by 1 and only remembered, so that the whole factorisation is one
    // basic block (the scheduler overlaps the tail of one trailing update with the next pivot's rsqrt chain).
    bool ok = true;
    double di = 1.0;
#pragma unroll
    for (int j = 0; j < KT; j++) {
        double piv = shfl_d(a[j], j);
        const bool good = piv > floor && piv > 1e-30 && piv < 1e30;
        ok = ok && good;
        piv = good ? piv : 1.0;
        const double inv = rsqrt_fast(piv);
        const double l = (lane > j) ? a[j] * inv : 0.0;
        di = (lane == j) ? inv : di;
        a[j] = l;
        Lc[j * WLD + lane] = l;
        if (j + 1 < KT) {
            __syncwarp();
            // trailing update of this lane's row: a[c] -= L[lane][j] L[c][j], c > j
            int c = j + 1;
            if (c & 1) {
                a[c] = fma(-l, Lc[j * WLD + c], a[c]);
                c++;
            }
#pragma unroll
            for (; c + 1 < KT; c += 2) {
                const double2 p = *reinterpret_cast<const double2*>(Lc + j * WLD + c);
                a[c] = fma(-l, p.x, a[c]);
                a[c + 1] = fma(-l, p.y, a[c + 1]);
            }
        }
    }
    __syncwarp();
    *dinv = di;
    return ok;
}

// Solve L L^T x = b with the factor left by chol_reg; lane r holds b_r on entry and x_r on return.
template <int KT>
__device__ __forceinline__ double chol_solve_reg(const double (&a)[KT], double dinv, int lane, double b,
                                                 const double* __restrict__ Lc) {
#pragma unroll
    for (int j = 0; j < KT; j++) {                       // L y = b : lanes > j subtract L[lane][j] y_j (a[j] = 0 elsewhere)
        const double y = shfl_d(b * dinv, j);
        b = fma(-a[j], y, b);
    }
    b *= dinv;
#pragma unroll
    for (int j = KT - 1; j >= 0; j--) {                  // L^T x = y : lanes < j subtract L[j][lane] x_j (0 elsewhere)
        const double x = shfl_d(b * dinv, j);
        b = fma(-Lc[lane * WLD + j], x, b);
    }
    return b * dinv;
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

// Factorisation half of the clamped solve.  H_row: row `lane` of the symmetric matrix in registers (entries c <= lane
// are used, like eigh(lower=True); rows / columns >= k are ignored); diag is added to the diagonal; W: per-warp tile of
// TILE doubles.  known_pd: the caller guarantees lambda_min(H + diag I) >= pert.
// Returns true when the Cholesky path ran (factor in `a` / *dinv / W, apply with chol_solve_reg) and false when the eigenvalue
// clamp is active (W then holds the Jacobi-orthogonalised tile, apply with jacobi_apply_tile).  Uniform.
template <int KT, typename R>
__device__ __forceinline__ bool safe_factor_warp(const R (&H_row)[KT], double diag, int k, int lane, double pert,
                                                 bool chol_fastpath, bool known_pd, double* W, double (&a)[KT],
                                                 double* dinv, const double* base_row = nullptr) {
    // base_row: row `lane` (KT doubles) of a float64 matrix added to H in double (the shared part of the Hessian)
    const bool act = lane < k;
    if (chol_fastpath) {
        bool ok = true;
        // pass 0: test factorisation of H - pert I (skipped when known_pd); pass 1: the factorisation that is used.
        // One runtime loop so that the unrolled factorisation exists once in the instruction stream.
        for (int pass = known_pd ? 1 : 0; pass < 2 && ok; pass++) {
            const double shift = diag - (pass == 0 ? pert : 0.0);
            double tr = 0.0;
#pragma unroll
            for (int c = 0; c < KT; c++) {
                // rows / columns beyond k: identity padding
                a[c] = (act && c < k) ? double(H_row[c]) + (base_row != nullptr ? base_row[c] : 0.0) + (c == lane ? shift : 0.0)
                                      : (c == lane ? 1.0 : 0.0);
                if (c == lane && act) tr = fabs(a[c]);
            }
            double floor = 0.0;
            if (pass == 0) floor = 1e-13 * (warp_sum(tr) + pert);
            ok = chol_reg<KT>(a, lane, floor, W, dinv);
        }
        if (ok) return true;
    }
    // eigenvalue clamp active (or fast path disabled): Jacobi on the symmetric tile built from the lower triangle
#pragma unroll
    for (int c = 0; c < KT; c++)
        if (c < k && c <= lane && act)
            W[lane * WLD + c] = double(H_row[c]) + (base_row != nullptr ? base_row[c] : 0.0) + (c == lane ? diag : 0.0);
    __syncwarp();
    for (int c = lane + 1; c < k; c++)
        if (act) W[lane * WLD + c] = W[c * WLD + lane];
    __syncwarp();
    jacobi_sweeps_tile(W, k, lane, pert);
    return false;
}

// x = S(H) g for one right-hand side after safe_factor_warp (lane r holds g_r / returns x_r).
template <int KT>
__device__ __forceinline__ double safe_apply_warp(bool factored, const double (&a)[KT], double dinv, int k, int lane,
                                                  double g, double pert, const double* W) {
    const double gr = lane < k ? g : 0.0;
    if (factored) return chol_solve_reg<KT>(a, dinv, lane, gr, W);
    return jacobi_apply_tile(W, k, lane, gr, pert);
}

}  // namespace wsolve
}  // namespace pycmf

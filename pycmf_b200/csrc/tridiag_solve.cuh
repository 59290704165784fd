// Eigenvalue-clamped solve WITHOUT a full eigendecomposition, one CTA per matrix (float64):
//     x = S(H) g,   S(H) = Q diag(1 / max(|lambda|, p)) Q^T                       (reference _safe_invert, cmf_solvers.py:346-356)
//       = g / p + sum_{|lambda_i| >= p} (1 / |lambda_i| - 1 / p) (q_i . g) q_i
// Only the eigenpairs ABOVE the clamp level enter, and the clamped part of the spectrum (the rank-deficient bulk of a
// sampled weighted Gram) never has to be resolved.  Stages:
//   1. Householder tridiagonalisation H = P T P^T in shared memory (reflectors kept in the rows they annihilate), with
//      y = P^T g applied on the fly.  4/3 k^3 flop, column-owned updates (conflict-free), ~7 barriers per step.
//   2. Sturm counts at -p and +p give the indices of the eigenvalues that matter; each gets its own thread and is bisected
//      (three-term recurrence of the leading principal minors with rescaling: 2 dependent FMAs per element, no division).
//   3. One thread per wanted eigenvector: inverse iteration with a pivoted LU of the shifted tridiagonal (EISPACK tinvit /
//      LAPACK dstein scheme), three solves, vectors and LU factors in a per-CTA global scratch laid out [vector][element]
//      (each thread streams its own rows through L1).
//   4. Vectors whose eigenvalues are closer than 1e-5 ||T|| are orthonormalised (modified Gram-Schmidt, CTA-wide dots): a
//      multiple eigenvalue needs an orthonormal basis of its eigenspace, nothing more.
//   5. z = y / p + sum coef_i z_i, x = P z (reflectors applied in reverse).
// ~16x fewer flops than the ten one-sided Jacobi sweeps this replaces at k = 128.  Any sign of trouble (a vector that does
// not survive the orthogonalisation) returns false and the caller falls back to Jacobi.  Prototype with the same arithmetic:
// scripts/tridiag_clamped_solve.py (1e-13 against eigh on rank-deficient, indefinite, clustered and clamp-level spectra).
#pragma once
#include "common.cuh"

namespace pycmf {
namespace tri {

// both sums to every thread; red: 64 doubles of shared memory
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < nwarps; w++) { sa += red[w]; sb += red[32 + w]; }
    a = sa;
    b = sb;
}

// number of eigenvalues of the (scaled) tridiagonal that are < x: sign changes of the leading principal minors
__device__ __forceinline__ int sturm_count(const double* __restrict__ ds, const double* __restrict__ es2, int k, double x) {
    double p0 = 1.0, p1 = ds[0] - x;
    bool sp = p1 < 0.0;                          // p0 > 0; an exact zero takes the sign opposite to its predecessor
    if (p1 == 0.0) sp = true;
    int cnt = sp ? 1 : 0;
    for (int i = 1; i < k; i++) {
        const double p2 = fma(ds[i] - x, p1, -es2[i] * p0);
        p0 = p1;
        p1 = p2;
        const bool s = p1 < 0.0 || (p1 == 0.0 && !sp);
        cnt += (s != sp) ? 1 : 0;
        sp = s;
        if ((i & 7) == 7) {                      // entries are <= 1 in magnitude: growth <= 3.3 per step, decay unbounded
            const double mx = fmax(fabs(p0), fabs(p1));
            if (mx > 1e100) { p0 *= 1e-100; p1 *= 1e-100; }
            else if (mx < 1e-100 && mx > 0.0) { p0 *= 1e100; p1 *= 1e100; }
        }
    }
    return cnt;
}

// work (shared memory, doubles): 16 k + 328.  zg (global, doubles): 6 k (k + 16) for this CTA.
__host__ __device__ constexpr size_t work_doubles(int k) { return size_t(16) * k + 72 + 256; }
// per-CTA global scratch of the inverse iteration: six arrays [vector][element], row pitch k + 16 doubles (an odd number of
// 128-byte lines, so that the rows of different vectors spread over all L1 sets)
__host__ __device__ constexpr int vec_pitch(int k) { return k + 16; }
__host__ __device__ constexpr size_t scratch_doubles(int k) { return size_t(6) * k * vec_pitch(k); }


// request the line of p into L1 (the inverse-iteration sweeps walk per-thread rows of the global scratch)
__device__ __forceinline__ void pf_l1(const double* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// shared-memory store of v when x == y.  Inline PTX on purpose: written as `if (u == sel) val = a[u]` over the unrolled
// register array, the compiler turns the select chain into an indexed load and moves the whole array to local memory.
__device__ __forceinline__ void st_shared_if_eq(uint32_t saddr, int x, int y, double v) {
    asm volatile("{\n .reg .pred p;\n setp.eq.s32 p, %1, %2;\n @p st.shared.f64 [%0], %3;\n}"
                 :: "r"(saddr), "r"(x), "r"(y), "d"(v) : "memory");
}

// ---- 1'. tridiagonalisation with the trailing matrix in REGISTERS (k = KR = 64 / 128, blockDim.x = 2 KR) --------------------
// The shared-memory form above is latency-bound: 8 warps walk dependent shared-memory loads and ~10 barriers per Householder
// step (ncu: 57 % of the kernel's samples, issue slots 22 % busy).  Here the symmetric matrix lives in registers, column-owned:
// the lane pair (2c, 2c + 1) holds column c, lane h = tid & 1 its rows r = 2 i + h as a[i] (KR / 2 doubles per thread, every
// index a compile-time constant).  Per step: row j (published to shared memory by the previous step's update) gives the
// reflector; p = S v is KR / 2 register FMAs per thread against broadcast reads of v, the pair combines by one shuffle;
// S -= v q^T + q v^T is two FMAs per element against broadcast reads of v and q.  Four barriers per step, no traffic on the
// matrix.  Blocks of 8 register elements whose rows are all <= j and warps whose columns are all <= j are skipped (uniform
// branches).  Same arithmetic as the loop above (reflectors in the rows of W they annihilate, y = P^T g on the fly).
// buf: 3 KR + 8 doubles of shared memory, 16-byte aligned.  red: 64 doubles.
template <int KR>
__device__ __forceinline__ void tridiag_reg(double* __restrict__ W, double* __restrict__ d, double* __restrict__ e,
                                            double* __restrict__ beta, double* __restrict__ y, double* __restrict__ buf,
                                            double* __restrict__ red) {
    constexpr int k = KR, HN = KR / 2, VP = HN + 2, NB = HN / 8, NW = KR / 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid >> 1, h = tid & 1;
    double* colbuf = buf;                 // KR      row j of the current matrix
    double* vd = buf + KR;                // 2 x VP  reflector, de-interleaved: v[r] at vd[(r & 1) * VP + (r >> 1)]
    double* qd = vd + 2 * VP;             // 2 x VP  q, same layout
    const uint32_t col_c = uint32_t(__cvta_generic_to_shared(colbuf + c));
    double a[HN];
#pragma unroll
    for (int i = 0; i < HN; i++) a[i] = W[(2 * i + h) * k + c];
    if (h == 0) colbuf[c] = a[0];
    // The step loop is unrolled over blocks of 16 steps (jb) so that every register index below is a compile-time constant:
    // during steps 16 jb .. 16 jb + 15 the register blocks b < jb are dead, and the row to publish sits in block jb (or is
    // element 0 of block jb + 1).
#pragma unroll
    for (int jb = 0; jb < NB; jb++) {
        const int jend = jb == NB - 1 ? 14 : 16;                           // steps j = 0 .. k - 3
        for (int jj = 0; jj < jend; jj++) {
            const int j = 16 * jb + jj;
            __syncthreads();                                               // (A) colbuf = row j
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < KR / 32; u++) {
                const int r = lane + 32 * u;
                const double x = colbuf[r];
                s = r > j ? fma(x, x, s) : s;
            }
            s = warp_sum(s);                                               // same operations in every warp: uniform value
            const double x0 = colbuf[j + 1], dj = colbuf[j];
            const double tail2 = s - x0 * x0;
            const int selx = h == ((j + 1) & 1) ? (j + 1) >> 1 : -1;       // row j + 1 is a[(j + 1) / 2] of the lanes h == (j + 1) % 2
            if (!(tail2 > 0.0)) {                                          // nothing to annihilate (uniform decision)
                if (tid == 0) { d[j] = dj; e[j + 1] = x0; beta[j] = 0.0; }
                __syncthreads();                                           // everyone is done with colbuf
#pragma unroll
                for (int u = 0; u < 8; u++) st_shared_if_eq(col_c, selx, 8 * jb + u, a[8 * jb + u]);
                if (jb + 1 < NB) st_shared_if_eq(col_c, selx, 8 * (jb + 1), a[jb + 1 < NB ? 8 * (jb + 1) : 0]);
                continue;
            }
            const double alpha = x0 >= 0.0 ? -sqrt(s) : sqrt(s);
            const double v0 = x0 - alpha;
            const double bq = 2.0 / (tail2 + v0 * v0);
            if (tid < k) {
                const int r = tid;
                const double vr = r <= j ? 0.0 : (r == j + 1 ? v0 : colbuf[r]);
                vd[(r & 1) * VP + (r >> 1)] = vr;
                if (r > j) W[j * k + r] = vr;                              // the reflector lives in the row it annihilated
            }
            if (tid == 0) { d[j] = dj; e[j + 1] = alpha; beta[j] = bq; }
            __syncthreads();                                               // (B) v
            const bool live = 16 * warp + 15 > j;                          // this warp still owns a column of the trailing block
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            if (live) {
#pragma unroll
                for (int b = jb; b < NB; b++) {                            // rows <= j inside block jb meet v = 0
                    const double2* vv = reinterpret_cast<const double2*>(vd + h * VP + 8 * b);
                    const double2 v01 = vv[0], v23 = vv[1], v45 = vv[2], v67 = vv[3];
                    p0 = fma(a[8 * b + 0], v01.x, p0); p1 = fma(a[8 * b + 1], v01.y, p1);
                    p2 = fma(a[8 * b + 2], v23.x, p2); p3 = fma(a[8 * b + 3], v23.y, p3);
                    p0 = fma(a[8 * b + 4], v45.x, p0); p1 = fma(a[8 * b + 5], v45.y, p1);
                    p2 = fma(a[8 * b + 6], v67.x, p2); p3 = fma(a[8 * b + 7], v67.y, p3);
                }
            }
            double pc = (p0 + p1) + (p2 + p3);
            pc += __shfl_xor_sync(0xffffffffu, pc, 1);                     // both lanes of the pair: p_c = (S v)_c
            const double vc = vd[(c & 1) * VP + (c >> 1)];                 // 0 for c <= j
            double qc = c > j ? bq * pc : 0.0;
            double vp = h == 0 ? vc * qc : 0.0;
            double vy = h == 0 ? vc * y[c] : 0.0;
            vp = warp_sum(vp);
            vy = warp_sum(vy);
            if (lane == 0) { red[warp] = vp; red[32 + warp] = vy; }
            __syncthreads();                                               // (C) partial sums
            double svp = 0.0, svy = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) { svp += red[w]; svy += red[32 + w]; }
            const double Kc = 0.5 * bq * svp;
            qc -= Kc * vc;
            if (h == 0) {
                qd[(c & 1) * VP + (c >> 1)] = qc;
                y[c] -= bq * svy * vc;                                     // y <- H_j y
            }
            __syncthreads();                                               // (D) q
            if (live) {
                // S -= v q^T + q v^T on the elements this thread owns
#pragma unroll
                for (int b = jb; b < NB; b++) {
                    const double2* vv = reinterpret_cast<const double2*>(vd + h * VP + 8 * b);
                    const double2* qq = reinterpret_cast<const double2*>(qd + h * VP + 8 * b);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double2 v2 = vv[u], q2 = qq[u];
                        a[8 * b + 2 * u] = fma(-q2.x, vc, fma(-v2.x, qc, a[8 * b + 2 * u]));
                        a[8 * b + 2 * u + 1] = fma(-q2.y, vc, fma(-v2.y, qc, a[8 * b + 2 * u + 1]));
                    }
                }
                // publish row j + 1 of the updated matrix for the next step
#pragma unroll
                for (int u = 0; u < 8; u++) st_shared_if_eq(col_c, selx, 8 * jb + u, a[8 * jb + u]);
                if (jb + 1 < NB) st_shared_if_eq(col_c, selx, 8 * (jb + 1), a[jb + 1 < NB ? 8 * (jb + 1) : 0]);
            }
        }
    }
    __syncthreads();
    W[(k - 2 + h) * k + c] = a[HN - 1];                                    // rows k - 2, k - 1: the last 2 x 2 block
    __syncthreads();
}

// KR > 0: k == KR and blockDim.x == 2 KR, tridiagonalisation in registers (tridiag_reg)
template <int KR = 0>
__device__ bool clamped_solve(double* __restrict__ W, int k, const double* __restrict__ g, double* __restrict__ x,
                              double* __restrict__ work, double* __restrict__ zg, double pert) {
    const int tid = threadIdx.x, NT = blockDim.x;
    double* d = work;                 // k   diagonal of T
    double* e = d + k;                // k   e[i] couples i - 1 and i (e[0] = 0)
    double* beta = e + k;             // k   reflector scalars
    double* v = beta + k;             // k   current reflector / later: scaled diagonal
    double* q = v + k;                // k   / later: squared scaled off-diagonal
    double* pp = q + k;               // 4 k partial mat-vec sums / later: scaled off-diagonal
    double* y = pp + 4 * k;           // k   P^T g, then the combined vector
    double* lam = y + k;              // k   wanted eigenvalues (scaled)
    double* coef = lam + k;           // k
    int* cstart = reinterpret_cast<int*>(coef + k);   // 6 k + 2 ints in 3 k + 1 doubles: cluster starts, block of a vector,
    int* vblk = cstart + k;                           //   block starts, per-block counts below -p / below +p, wanted offsets
    int* bstart = vblk + k;                           // k + 1
    int* bneg = bstart + k + 1;
    int* blt = bneg + k;
    int* woff = blt + k;                              // k + 1
    double* red = coef + 4 * k + 1;   // 64
    __shared__ int s_fail, s_nblk;
    if (tid == 0) s_fail = 0;
    for (int r = tid; r < k; r += NT) y[r] = g[r];
    __syncthreads();

    // ---- 1. tridiagonalisation ----------------------------------------------------------------------------------------
    if constexpr (KR > 0) {
        tridiag_reg<KR>(W, d, e, beta, y, pp, red);
    } else
    for (int j = 0; j < k - 2; j++) {
        const int m = k - j - 1;
        double* xr = W + size_t(j) * k + j + 1;              // row j right of the diagonal (= column j below it)
        double s = 0.0, dummy = 0.0;
        for (int c = tid; c < m; c += NT) s = fma(xr[c], xr[c], s);
        block_sum2(s, dummy, red);
        const double x0 = xr[0];
        const double tail2 = s - x0 * x0;
        if (!(tail2 > 0.0)) {                               // nothing to annihilate (uniform decision)
            if (tid == 0) { d[j] = W[size_t(j) * k + j]; e[j + 1] = x0; beta[j] = 0.0; }
            continue;
        }
        const double alpha = x0 >= 0.0 ? -sqrt(s) : sqrt(s);
        const double v0 = x0 - alpha;
        const double b = 2.0 / (tail2 + v0 * v0);
        for (int c = tid; c < m; c += NT) v[c] = c == 0 ? v0 : xr[c];
        __syncthreads();
        // p = b S v, S = trailing m x m block (symmetric, kept in full): thread owns a column, row segments interleaved
        const int mp = (m + 31) & ~31;
        const int tpc = max(1, min(4, NT / mp));
        // (32-bit indices -- k^2 <= 65536 -- and four independent loads per trip: with 8 warps per SM these loops are latency-
        //  bound, ncu put 57 % of the kernel here while they walked one dependent generic load at a time)
        if (tid < tpc * mp) {
            const int c = tid % mp, seg = tid / mp;
            if (c < m) {
                const double* Sc = W + (j + 1) * k + j + 1 + c;
                const int st = tpc * k;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int r = seg, o = seg * k;
                for (; r + 3 * tpc < m; r += 4 * tpc, o += 4 * st) {
                    const double s0 = Sc[o], s1 = Sc[o + st], s2 = Sc[o + 2 * st], s3 = Sc[o + 3 * st];
                    a0 = fma(s0, v[r], a0);
                    a1 = fma(s1, v[r + tpc], a1);
                    a2 = fma(s2, v[r + 2 * tpc], a2);
                    a3 = fma(s3, v[r + 3 * tpc], a3);
                }
                for (; r < m; r += tpc, o += st) a0 = fma(Sc[o], v[r], a0);
                pp[seg * k + c] = (a0 + a1) + (a2 + a3);
            }
        }
        __syncthreads();
        double vp = 0.0, vy = 0.0;
        for (int c = tid; c < m; c += NT) {
            double pc = 0.0;
            for (int sg = 0; sg < tpc; sg++) pc += pp[sg * k + c];
            pc *= b;
            q[c] = pc;
            vp = fma(v[c], pc, vp);
            vy = fma(v[c], y[j + 1 + c], vy);
        }
        block_sum2(vp, vy, red);
        const double K = 0.5 * b * vp;
        for (int c = tid; c < m; c += NT) {
            q[c] -= K * v[c];
            y[j + 1 + c] -= b * vy * v[c];                   // y <- H_j y
        }
        __syncthreads();
        // S -= v q^T + q v^T
        if (tid < tpc * mp) {
            const int c = tid % mp, seg = tid / mp;
            if (c < m) {
                const double vc = v[c], qc = q[c];
                double* Sc = W + (j + 1) * k + j + 1 + c;
                const int st = tpc * k;
                int r = seg, o = seg * k;
                for (; r + 3 * tpc < m; r += 4 * tpc, o += 4 * st) {
                    double s0 = Sc[o], s1 = Sc[o + st], s2 = Sc[o + 2 * st], s3 = Sc[o + 3 * st];
                    s0 -= fma(v[r], qc, q[r] * vc);
                    s1 -= fma(v[r + tpc], qc, q[r + tpc] * vc);
                    s2 -= fma(v[r + 2 * tpc], qc, q[r + 2 * tpc] * vc);
                    s3 -= fma(v[r + 3 * tpc], qc, q[r + 3 * tpc] * vc);
                    Sc[o] = s0; Sc[o + st] = s1; Sc[o + 2 * st] = s2; Sc[o + 3 * st] = s3;
                }
                for (; r < m; r += tpc, o += st) Sc[o] -= fma(v[r], qc, q[r] * vc);
            }
        }
        for (int c = tid; c < m; c += NT) xr[c] = v[c];      // the reflector lives in the row it annihilated
        if (tid == 0) { d[j] = W[size_t(j) * k + j]; e[j + 1] = alpha; beta[j] = b; }
        __syncthreads();
    }
    if (tid == 0) {
        e[0] = 0.0;
        if (k >= 2) {
            d[k - 2] = W[size_t(k - 2) * k + k - 2];
            e[k - 1] = W[size_t(k - 2) * k + k - 1];
            beta[k - 2] = 0.0;
        }
        d[k - 1] = W[size_t(k - 1) * k + k - 1];
        beta[k - 1] = 0.0;
    }
    __syncthreads();

    // ---- 2. the eigenvalues above the clamp level ------------------------------------------------------------------------
    double* ds = v;
    double* es2 = q;
    double* es = pp;
    double tn = 0.0;
    for (int i = 0; i < k; i++) tn = fmax(tn, fabs(d[i]) + fabs(e[i]) + (i + 1 < k ? fabs(e[i + 1]) : 0.0));   // every thread
    __syncthreads();
    if (!(tn > 0.0) || !(tn < 1e300)) {                       // zero matrix (everything clamped) or garbage
        for (int r = tid; r < k; r += NT) y[r] = y[r] / pert;
        __syncthreads();
    } else {
        const double inv = 1.0 / tn;
        for (int i = tid; i < k; i += NT) {
            ds[i] = d[i] * inv;
            es[i] = e[i] * inv;
            es2[i] = (e[i] * inv) * (e[i] * inv);
        }
        __syncthreads();
        const double ps = pert * inv;
        // ---- 2a. split T where an off-diagonal is negligible (LAPACK dstebz criterion): eigenvectors live on their block, so
        //          a multiple eigenvalue of H (rank-deficient Gram + l2 I, ...) becomes one simple eigenvalue per block and the
        //          vectors of different blocks are orthogonal by support
        if (tid == 0) {
            int nb = 0;
            bstart[0] = 0;
            for (int i = 1; i < k; i++)
                if (es2[i] <= 4.930380657631324e-32 * fabs(ds[i - 1] * ds[i]) + 1e-300) { es[i] = 0.0; es2[i] = 0.0; bstart[++nb] = i; }
            bstart[++nb] = k;
            s_nblk = nb;
        }
        __syncthreads();
        // reciprocals of the off-diagonals that survived the split: the LU factorisations below divide by them when they pivot
        double* esi = pp + k;
        for (int i = tid; i < k; i += NT) esi[i] = es[i] != 0.0 ? 1.0 / es[i] : 0.0;
        const int nblk = s_nblk;
        for (int bI = tid; bI < nblk; bI += NT) {             // wanted eigenvalues of every block: below -p and above +p
            const int s0 = bstart[bI], sz = bstart[bI + 1] - s0;
            const int nn = sturm_count(ds + s0, es2 + s0, sz, -ps), nl = sturm_count(ds + s0, es2 + s0, sz, ps);
            bneg[bI] = nn;
            blt[bI] = nl;
            woff[bI + 1] = nn + (sz - nl);
        }
        __syncthreads();
        if (tid == 0) {
            woff[0] = 0;
            for (int bI = 0; bI < nblk; bI++) woff[bI + 1] += woff[bI];
        }
        __syncthreads();
        const int nw = woff[nblk];                            // wanted (uniform); <= k <= NT
        // ---- 2b / 3. one thread per wanted eigenpair: bisection inside its block, then inverse iteration ---------------------
        // arrays [vector * ZP + element]: a thread walks ITS vector through consecutive addresses (one L2 round trip per 16
        // elements, then L1 hits).  The [element][vector] layout this replaces was coalesced across the threads, but with at most
        // k threads per SM in these chains every element paid the L2 latency: ncu put 26 % of the kernel's samples on the
        // barrier behind the sweeps.
        const int ZP = vec_pitch(k);
        double* Z = zg;
        double* U0 = Z + size_t(k) * ZP;       // reciprocal pivots
        double* U1 = U0 + size_t(k) * ZP;
        double* U2 = U1 + size_t(k) * ZP;
        double* L = U2 + size_t(k) * ZP;
        double* PV = L + size_t(k) * ZP;       // 1.0 where rows were swapped
        // 2b. multi-section: the CTA's threads are dealt out G per wanted eigenvalue and cut its bracket into G + 1 parts per
        //     round (log2(G + 1) bits per Sturm evaluation instead of 1: with 16 wanted eigenvalues and 256 threads, 14 rounds
        //     instead of 55).  G is a power of two <= 32, so a group sits inside one warp: the G answers are combined by a
        //     ballot, every lane of the group keeps the bracket in registers and the rounds need no barrier (the two barriers
        //     and the leader's serial scan of a flag array per round were 20 % of the kernel's samples).  The final brackets
        //     go to the (now dead) unscaled d / e arrays.
        {
            double* blo = d;
            double* bhi = e;
            int G = 1;
            while (2 * G <= min(32, NT / max(nw, 1))) G *= 2;
            const int rounds = nw > 0 ? int(56.0f / log2f(float(G + 1))) + 2 : 0;     // bracket 2 -> 2^-54
            const int tq = tid / G, gq = tid % G;
            const int lane = tid & 31;
            const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
            const int gshift = lane & ~(G - 1);
            const double inv = 1.0 / double(G + 1);
            int qs0 = 0, qsz = 0, qidx = 0;
            double lo = 0.0, hi = 0.0;
            if (tq < nw) {
                int bI = 0;
                while (woff[bI + 1] <= tq) bI++;
                qs0 = bstart[bI];
                qsz = bstart[bI + 1] - qs0;
                const int jl = tq - woff[bI];
                qidx = jl < bneg[bI] ? jl : blt[bI] + (jl - bneg[bI]);
                lo = qsz == 1 ? ds[qs0] : -1.0009765625;
                hi = qsz == 1 ? ds[qs0] : 1.0009765625;
                if (gq == 0) vblk[tq] = bI;
            }
            __syncthreads();                                          // (d / e are read above by nobody any more)
            for (int rd = 0; rd < rounds; rd++) {
                // interior point m = gq + 1 of the bracket: x_m = lo + (hi - lo) (m / (G + 1)), the same expression below
                const double xq = lo + (hi - lo) * (double(gq + 1) * inv);
                int above = 1;                                        // "more than idx eigenvalues below xq"
                if (tq < nw && qsz > 1 && hi > lo) above = sturm_count(ds + qs0, es2 + qs0, qsz, xq) > qidx ? 1 : 0;
                const unsigned grp = (__ballot_sync(0xffffffffu, above != 0) >> gshift) & gmask;
                const int j = grp != 0u ? __ffs(int(grp)) - 1 : G;    // first interior point with the eigenvalue below it
                const double nlo = j > 0 ? lo + (hi - lo) * (double(j) * inv) : lo;
                const double nhi = j < G ? lo + (hi - lo) * (double(j + 1) * inv) : hi;
                const double l2 = fmax(lo, fmin(nlo, hi)), h2 = fmin(hi, fmax(nhi, lo));
                lo = l2;
                hi = h2;
            }
            if (tq < nw && gq == 0) { blo[tq] = lo; bhi[tq] = hi; }
            __syncthreads();
        }
        const int t = tid;
        int s0 = 0, sz = 0;
        double *Zb = Z, *U0b = U0, *U1b = U1, *U2b = U2, *Lb = L, *PVb = PV;
        if (tid < nw) {
            const int bI = vblk[t];
            s0 = bstart[bI];
            sz = bstart[bI + 1] - s0;
            const double* bd = ds + s0;
            const double* be = es + s0;        // be[i] couples i - 1 and i inside the block (be[0] is never used)
            const double* bei = esi + s0;      // 1 / be[i]
            const double lm = 0.5 * (d[t] + e[t]);
            lam[t] = lm;
            for (int i = 0; i < s0; i++) Z[size_t(t) * ZP + i] = 0.0;
            for (int i = s0 + sz; i < k; i++) Z[size_t(t) * ZP + i] = 0.0;
            Zb = Z + size_t(t) * ZP + s0;                  // element i of the block at Zb[i]
            U0b = U0 + size_t(t) * ZP + s0; U1b = U1 + size_t(t) * ZP + s0; U2b = U2 + size_t(t) * ZP + s0;
            Lb = L + size_t(t) * ZP + s0; PVb = PV + size_t(t) * ZP + s0;
            if (sz == 1) {
                Zb[0] = 1.0;
            } else {
                const double tiny = 2.220446049250313e-16;
                double r0 = bd[0] - lm, r1 = be[1], r2 = 0.0;
                // One division per row at most, and only when the running pivot is kept (the reciprocal of an off-diagonal pivot
                // comes from shared memory); U is stored with its rows already divided by the pivot, so that the back substitution
                // carries ONE dependent FMA per element.  (Two divisions per row, one of them on the dependent chain, made this
                // loop 17 % of the kernel's samples.)
                for (int i = 0; i < sz - 1; i++) {
                    const double q0 = be[i + 1], q1 = bd[i + 1] - lm, q2 = i + 2 < sz ? be[i + 2] : 0.0;
                    double ui, u1, u2, l, pv;
                    if (fabs(q0) > fabs(r0)) {
                        pv = 1.0; ui = bei[i + 1]; u1 = q1; u2 = q2;
                        l = r0 * ui;
                        r0 = r1 - l * q1; r1 = r2 - l * q2; r2 = 0.0;
                    } else {
                        if (r0 == 0.0) r0 = tiny;
                        pv = 0.0; ui = 1.0 / r0; u1 = r1; u2 = r2;
                        l = q0 * ui;
                        r0 = q1 - l * r1; r1 = q2 - l * r2; r2 = 0.0;
                    }
                    U0b[i] = ui; U1b[i] = u1 * ui; U2b[i] = u2 * ui;
                    Lb[i] = l; PVb[i] = pv;
                }
                if (r0 == 0.0) r0 = tiny;
                U0b[sz - 1] = 1.0 / r0; U1b[sz - 1] = 0.0; U2b[sz - 1] = 0.0;
                uint32_t st = 0x9e3779b9u * uint32_t(t + 1) + 0x7f4a7c15u;     // start vector: positive pseudo-random entries
                for (int i = 0; i < sz; i++) {
                    st = st * 1664525u + 1013904223u;
                    Zb[i] = 0.5 + double(st >> 8) * (1.0 / 16777216.0);
                }
            }
        }
        __syncthreads();
        // clusters of close eigenvalues of the SAME block: their vectors are re-orthogonalised after EVERY sweep (as dstein does;
        // orthogonalising only at the end lets the amplification ratios inside a multiple eigenvalue compound: 1e-9 instead of 1e-12)
        if (tid == 0) {
            int start = 0;
            for (int u = 0; u < nw; u++) {
                // separately computed vectors of eigenvalues a gap delta ||T|| apart overlap by ~eps / delta (2e-11 here)
                if (u > 0 && (vblk[u] != vblk[u - 1] || !(lam[u] - lam[u - 1] < 1e-5))) start = u;
                cstart[u] = start;
            }
        }
        __syncthreads();
        for (int it = 0; it < 3; it++) {
            if (tid < nw && sz > 1) {
                if (it > 0) {                                  // apply L^-1 P; the running element stays in a register, so the
                    double cur = Zb[0];                        // loads of one step do not wait for the stores of the previous
                    if (sz > 16) { pf_l1(Zb + 16); pf_l1(Lb + 16); pf_l1(PVb + 16); }
#pragma unroll 8
                    for (int i = 0; i < sz - 1; i++) {
                        // the rows are walked at one dependent FMA per element: the next lines are requested two ahead
                        if ((i & 15) == 0 && i + 32 < sz) { pf_l1(Zb + i + 32); pf_l1(Lb + i + 32); pf_l1(PVb + i + 32); }
                        const double xn = Zb[i + 1], l = Lb[i], pv = PVb[i];
                        if (pv != 0.0) { Zb[i] = xn; cur = cur - l * xn; }
                        else { Zb[i] = cur; cur = xn - l * cur; }
                    }
                    Zb[sz - 1] = cur;
                }
                double x1 = 0.0, x2 = 0.0, nrm = 0.0, s2 = 0.0; // back substitution with U (bandwidth 2, rows divided by the pivot)
                if (sz > 16) { pf_l1(Zb + sz - 17); pf_l1(U0b + sz - 17); pf_l1(U1b + sz - 17); pf_l1(U2b + sz - 17); }
#pragma unroll 8
                for (int i = sz - 1; i >= 0; i--) {
                    if (((sz - 1 - i) & 15) == 0 && i >= 32) {
                        pf_l1(Zb + i - 32); pf_l1(U0b + i - 32); pf_l1(U1b + i - 32); pf_l1(U2b + i - 32);
                    }
                    const double zi = Zb[i], a1 = U1b[i], a2 = U2b[i], a0 = U0b[i];
                    const double tv = fma(-a1, x1, fma(-a2, x2, zi * a0));   // the inner FMA does not wait for x1
                    Zb[i] = tv;
                    x2 = x1; x1 = tv;
                    nrm = fmax(nrm, fabs(tv));
                    s2 = fma(tv, tv, s2);
                }
                // to unit 2-norm.  The solve amplifies by up to ~1e16 per tiny pivot: when the plain sum of squares left the
                // representable range, rescale by the max norm first
                double sc2;
                if (s2 > 1e-280 && s2 < 1e280) {
                    sc2 = 1.0 / sqrt(s2);
                } else {
                    const double sc = nrm > 0.0 ? 1.0 / nrm : 1.0;
                    s2 = 0.0;
                    for (int i = 0; i < sz; i++) { const double tv = Zb[i] * sc; s2 = fma(tv, tv, s2); }
                    sc2 = sc / sqrt(s2);
                }
#pragma unroll 8
                for (int i = 0; i < sz; i++) {
                    if ((i & 15) == 0 && i + 32 < sz) pf_l1(Zb + i + 32);
                    Zb[i] *= sc2;
                }
            }
            __syncthreads();
            for (int u = 1; u < nw; u++) {
                const int c0 = cstart[u];
                if (c0 == u) continue;                        // uniform
                for (int i = c0; i < u; i++) {
                    double dot = 0.0, dummy = 0.0;
                    for (int el = tid; el < k; el += NT) dot = fma(Z[size_t(i) * ZP + el], Z[size_t(u) * ZP + el], dot);
                    block_sum2(dot, dummy, red);
                    for (int el = tid; el < k; el += NT) Z[size_t(u) * ZP + el] -= dot * Z[size_t(i) * ZP + el];
                    __syncthreads();
                }
                double n2 = 0.0, dummy = 0.0;
                for (int el = tid; el < k; el += NT) n2 = fma(Z[size_t(u) * ZP + el], Z[size_t(u) * ZP + el], n2);
                block_sum2(n2, dummy, red);
                if (!(n2 > 1e-8)) { if (tid == 0 && it == 2) s_fail = 1; if (!(n2 > 0.0)) n2 = 1.0; }   // a copy of an earlier vector
                const double sc = 1.0 / sqrt(n2);
                for (int el = tid; el < k; el += NT) Z[size_t(u) * ZP + el] *= sc;
                __syncthreads();
            }
        }
        // ---- 5. combine --------------------------------------------------------------------------------------------------
        if (tid < nw) {
            double c = 0.0;
            for (int i = 0; i < k; i++) c = fma(Z[size_t(tid) * ZP + i], y[i], c);
            coef[tid] = (1.0 / (fabs(lam[tid]) * tn) - 1.0 / pert) * c;
        }
        __syncthreads();
        for (int el = tid; el < k; el += NT) {
            double acc = y[el] / pert;
            for (int t = 0; t < nw; t++) acc = fma(coef[t], Z[size_t(t) * ZP + el], acc);
            x[el] = acc;
        }
        __syncthreads();
        for (int el = tid; el < k; el += NT) y[el] = x[el];
        __syncthreads();
    }
    // ---- x = P y: reflectors in reverse -------------------------------------------------------------------------------------
    for (int j = k - 3; j >= 0; j--) {
        const double b = beta[j];
        if (b == 0.0) continue;                               // uniform
        const int m = k - j - 1;
        const double* vj = W + size_t(j) * k + j + 1;
        double s = 0.0, dummy = 0.0;
        for (int c = tid; c < m; c += NT) s = fma(vj[c], y[j + 1 + c], s);
        block_sum2(s, dummy, red);
        for (int c = tid; c < m; c += NT) y[j + 1 + c] -= b * s * vj[c];
        __syncthreads();                                      // the next step's partial sums read the updated y
    }
    __syncthreads();
    for (int el = tid; el < k; el += NT) x[el] = y[el];
    __syncthreads();
    return s_fail == 0;
}

}  // namespace tri
}  // namespace pycmf

// Newton-step kernels:
//   row_grad_hess     per-row gradient / weighted-Gram Hessian over (sampled) rows of the other factor
//   safe_solve        batched k x k eigenvalue-clamped solve  x = S(H) g   (reference _safe_invert,
//                     cmf_solvers.py:346-356) : Cholesky fast path when lambda_min(H) >= pert,
//                     one-sided Jacobi otherwise; always float64
//   newton_solve_rows F_i <- F_i - (g_i + l1 sign F_i + l2 F_i) S(H_i + l2 I), optional clamp
//   sample_indices    on-device per-row sampling without replacement (Feistel permutation prefix)
#include <type_traits>

#include "common.cuh"
#include "tridiag_solve.cuh"

namespace pycmf {
namespace {

// =============================================================================================
// row_grad_hess
// =============================================================================================
constexpr int TJ = 32;  // sampled rows of B staged per tile

template <typename T, int HB>
__global__ void __launch_bounds__(256)
row_grad_hess_kernel(int64_t rows, int64_t m, int k, const T* __restrict__ A, const T* __restrict__ B,
                     const T* __restrict__ Tgt, int64_t ldt, bool trans_t,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                     const T* __restrict__ vals, int link, T w,
                     const int32_t* __restrict__ idx, int64_t n_sample,
                     T* __restrict__ g, T* __restrict__ H, bool accumulate, int a_off, int b_off,
                     int64_t samples_per_split, int64_t g_split_stride, int64_t h_split_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int kp = k + 1;
    T* a_s = reinterpret_cast<T*>(smem_raw);          // k
    T* B_s = a_s + ((k + 3) & ~3);                    // TJ x kp
    T* r_s = B_s + TJ * kp;                           // TJ   (w * residual)
    T* w_s = r_s + TJ;                                // TJ   (w * f')
    T* t_s = w_s + TJ;                                // TJ   targets of the staged samples
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = tid >> 4, tb = tid & 15;
    const bool want_g = g != nullptr, want_h = H != nullptr;
    const bool sym = a_off == b_off;      // the CTA covers a diagonal block of H (always, unless k needs several register passes)

    for (int c = tid; c < k; c += 256) a_s[c] = A[i * k + c];
    T gacc = T(0);
    T hacc[HB][HB];
#pragma unroll
    for (int x = 0; x < HB; x++)
#pragma unroll
        for (int y = 0; y < HB; y++) hacc[x][y] = T(0);

    const int64_t total_all = idx != nullptr ? n_sample : m;
    const int64_t t_first = int64_t(blockIdx.y) * samples_per_split;
    const int64_t total = t_first + samples_per_split < total_all ? t_first + samples_per_split : total_all;
    if (g != nullptr) g += int64_t(blockIdx.y) * g_split_stride;
    if (H != nullptr) H += int64_t(blockIdx.y) * h_split_stride;
    int lo = 0, hi = 0;
    if (rowptr != nullptr) { lo = rowptr[i]; hi = rowptr[i + 1]; }

    for (int64_t t0 = t_first; t0 < total; t0 += TJ) {
        const int64_t rem_t = total - t0;
        const int cnt = rem_t < TJ ? int(rem_t) : TJ;
        __syncthreads();
        // stage the sampled rows of B
        for (int e = tid; e < TJ * k; e += 256) {
            int jj = e / k, c = e % k;
            T v = T(0);
            if (jj < cnt) {
                int64_t j = idx != nullptr ? int64_t(idx[i * n_sample + t0 + jj]) : t0 + jj;
                if (j >= 0) v = B[j * k + c];   // j < 0: sample lives on another shard
            }
            B_s[jj * kp + c] = v;
        }
        // targets of the tile, fetched in parallel (one thread per sample) instead of serially by lane 0 below
        if (want_g && tid < TJ) {
            T tg = T(0);
            if (tid < cnt) {
                const int64_t j = idx != nullptr ? int64_t(idx[i * n_sample + t0 + tid]) : t0 + tid;
                if (j >= 0) {
                    if (Tgt != nullptr) {
                        tg = trans_t ? Tgt[j * ldt + i] : Tgt[i * ldt + j];
                    } else if (rowptr != nullptr) {
                        int l = lo, h = hi;
                        while (l < h) {
                            int mid = (l + h) >> 1;
                            if (colidx[mid] < int(j)) l = mid + 1; else h = mid;
                        }
                        if (l < hi && colidx[l] == int(j)) tg = vals[l];
                    }
                }
            }
            t_s[tid] = tg;
        }
        __syncthreads();
        // estimates: each warp takes TJ / 8 sampled rows
        for (int jj = warp; jj < TJ; jj += 8) {
            T d = T(0);
            for (int c = lane; c < k; c += 32) d = fma(a_s[c], B_s[jj * kp + c], d);
            d = warp_sum(d);
            if (lane == 0) {
                T rr = T(0), ww = T(0);
                int64_t j = -1;
                if (jj < cnt) j = idx != nullptr ? int64_t(idx[i * n_sample + t0 + jj]) : t0 + jj;
                if (j >= 0) {
                    T est = d, fp = T(1);
                    if (link == PYCMF_LOGIT) { est = sigmoid_<T>(d); fp = est * (T(1) - est); }
                    const T tg = want_g ? t_s[jj] : T(0);
                    rr = w * (est - tg);
                    ww = w * fp;
                }
                r_s[jj] = rr;
                w_s[jj] = ww;
            }
        }
        __syncthreads();
        if (want_g && tid < k) {
#pragma unroll 8
            for (int jj = 0; jj < TJ; jj++) gacc = fma(r_s[jj], B_s[jj * kp + tid], gacc);
        }
        if (want_h) {
            for (int jj = 0; jj < cnt; jj++) {
                T wa[HB], bb[HB];
                const T wj = w_s[jj];
#pragma unroll
                for (int x = 0; x < HB; x++) {
                    int a = a_off + ta + 16 * x;
                    wa[x] = a < k ? wj * B_s[jj * kp + a] : T(0);
                }
#pragma unroll
                for (int y = 0; y < HB; y++) {
                    int b = b_off + tb + 16 * y;
                    bb[y] = b < k ? B_s[jj * kp + b] : T(0);
                }
                // H_i is symmetric: only the register tiles on and below the diagonal are accumulated (x > y: every element of
                // the tile has a > b when a_off == b_off), 36 of 64 at HB = 8; the mirror images are written out at the end
#pragma unroll
                for (int x = 0; x < HB; x++)
#pragma unroll
                    for (int y = 0; y < HB; y++)
                        if (x >= y || !sym) hacc[x][y] = fma(wa[x], bb[y], hacc[x][y]);
            }
        }
    }
    if (want_g && tid < k) {
        T prev = accumulate ? g[i * k + tid] : T(0);
        g[i * k + tid] = prev + gacc;
    }
    if (want_h) {
        T* Hi = H + i * int64_t(k) * k;
#pragma unroll
        for (int x = 0; x < HB; x++) {
            int a = a_off + ta + 16 * x;
            if (a >= k) continue;
#pragma unroll
            for (int y = 0; y < HB; y++) {
                int b = b_off + tb + 16 * y;
                if (b >= k) continue;
                if (sym && x < y) continue;                      // written as the mirror image of tile (y, x)
                T prev = accumulate ? Hi[a * k + b] : T(0);
                Hi[a * k + b] = prev + hacc[x][y];
                if (sym && x > y) {
                    prev = accumulate ? Hi[b * k + a] : T(0);
                    Hi[b * k + a] = prev + hacc[x][y];
                }
            }
        }
    }
}

// =============================================================================================
// row_grad_hess on the tensor cores (fp32, n_components 64 / 128, non-negative weight)
// =============================================================================================
// Same contract as row_grad_hess_kernel; the weighted Gram  H_i += sum_j ww_j b_j b_j^T  (2 s k^2 flop per row: 2.65e15 per
// iteration at C4's full size, the wall of the sampled logit Newton step) runs as warp-level MMAs.  The sample sets differ from
// row to row, so the operand is built per CTA: the tile of TJ gathered factor rows is scaled by sqrt(ww_j) and split ONCE into
// tf32 hi / lo parts in shared memory (C = C_hi + C_lo); then H += C_hi^T C_hi + C_hi^T C_lo + C_lo^T C_hi (3xTF32: fp32
// products) with mma.sync.m16n8k8 -- A and B fragments both come from the same two tiles, pitch = 8 (mod 32) words so that
// the fragment loads are conflict-free.  Eight warps, each 1/4 of the rows x 1/2 of the columns of H in registers.
// (Legacy warp-level path on purpose: 279 TFLOP/s TF32 measured on B200, scripts/hmma_rate.cu; tcgen05 wants its operands
// behind shared-memory descriptors with a 128-row M, which a per-row gathered 32-sample tile does not fill.)
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int RCM = 1024;  // CSR nonzeros of the row kept in shared memory for the target lookups
constexpr int TJM = 32;    // sampled rows per tile of the tensor-core kernel (64 measured slower: 51 vs 31.5 ms on the C4 slice)

template <int K>
__global__ void __launch_bounds__(256)
row_grad_hess_mma_kernel(int64_t rows, int64_t m, const float* __restrict__ A, const float* __restrict__ B,
                         const float* __restrict__ Tgt, int64_t ldt, bool trans_t,
                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                         const float* __restrict__ vals, int link, float w,
                         const int32_t* __restrict__ idx, int64_t n_sample,
                         float* __restrict__ g, float* __restrict__ H, bool accumulate,
                         int64_t samples_per_split, int64_t g_split_stride, int64_t h_split_stride) {
    constexpr int k = K;
    constexpr int KP = K + 8;                 // = 8 (mod 32) words
    constexpr int MT = K / 64;                // 16-row MMA tiles per warp (warp owns K / 4 rows)
    constexpr int NT = K / 16;                // 8-column MMA tiles per warp (warp owns K / 2 columns)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* a_s = reinterpret_cast<float*>(smem_raw);     // K
    float* B_s = a_s + K;                                 // TJM x KP  gathered rows (fp32)
    uint32_t* Chi = reinterpret_cast<uint32_t*>(B_s + TJM * KP);   // TJM x KP  tf32(sqrt(ww) b)
    uint32_t* Clo = Chi + TJM * KP;                        // TJM x KP  tf32(remainder)
    float* r_s = reinterpret_cast<float*>(Clo + TJM * KP); // TJM   (w * residual)
    float* w_s = r_s + TJM;                                // TJM   sqrt(w * f')
    float* t_s = w_s + TJM;                                // TJM   targets of the staged samples
    int* idx_s = reinterpret_cast<int*>(t_s + TJM);        // TJM   sample indices of the tile (-1 = none / other shard)
    int* rc_col = idx_s + TJM;                             // RCM   this row's CSR column indices ...
    float* rc_val = reinterpret_cast<float*>(rc_col + RCM);// RCM   ... and values (target lookups stay on chip)
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int m0 = (warp >> 1) * (K / 4), n0 = (warp & 1) * (K / 2);
    const bool want_g = g != nullptr;

    for (int c = tid; c < k; c += 256) a_s[c] = A[i * k + c];
    float gacc = 0.f;
    float acc[MT][NT][4];
#pragma unroll
    for (int x = 0; x < MT; x++)
#pragma unroll
        for (int y = 0; y < NT; y++)
#pragma unroll
            for (int z = 0; z < 4; z++) acc[x][y][z] = 0.f;

    const int64_t total_all = idx != nullptr ? n_sample : m;
    const int64_t t_first = int64_t(blockIdx.y) * samples_per_split;
    const int64_t total = t_first + samples_per_split < total_all ? t_first + samples_per_split : total_all;
    if (g != nullptr) g += int64_t(blockIdx.y) * g_split_stride;
    H += int64_t(blockIdx.y) * h_split_stride;
    int lo = 0, hi = 0;
    if (rowptr != nullptr) { lo = rowptr[i]; hi = rowptr[i + 1]; }
    // the row's nonzeros on chip (when they fit): every tile looks its 32 targets up by binary search, and seven dependent
    // global loads per lookup were a third of a tile's latency chain
    const bool row_cached = want_g && rowptr != nullptr && hi - lo <= RCM;
    if (row_cached)
        for (int e = tid; e < hi - lo; e += 256) { rc_col[e] = colidx[lo + e]; rc_val[e] = vals[lo + e]; }

    for (int64_t t0 = t_first; t0 < total; t0 += TJM) {
        const int64_t rem_t = total - t0;
        const int cnt = rem_t < TJM ? int(rem_t) : TJM;
        __syncthreads();
        // the tile's sample indices, read from global memory ONCE (they used to be re-read by the staging loop, the target
        // lookup and lane 0 of every estimate: three dependent round trips per tile), and the targets of the samples
        if (tid < TJM) {
            int j = -1;
            if (tid < cnt) j = idx != nullptr ? idx[i * n_sample + t0 + tid] : int(t0 + tid);
            idx_s[tid] = j;
            if (want_g) {
                float tg = 0.f;
                if (j >= 0) {
                    if (Tgt != nullptr) {
                        tg = trans_t ? Tgt[int64_t(j) * ldt + i] : Tgt[i * ldt + j];
                    } else if (row_cached) {
                        int l = 0, h = hi - lo;
                        while (l < h) {
                            const int mid = (l + h) >> 1;
                            if (rc_col[mid] < j) l = mid + 1; else h = mid;
                        }
                        if (l < hi - lo && rc_col[l] == j) tg = rc_val[l];
                    } else if (rowptr != nullptr) {
                        int l = lo, h = hi;
                        while (l < h) {
                            const int mid = (l + h) >> 1;
                            if (colidx[mid] < j) l = mid + 1; else h = mid;
                        }
                        if (l < hi && colidx[l] == j) tg = vals[l];
                    }
                }
                t_s[tid] = tg;
            }
        }
        __syncthreads();
        // stage the sampled rows of B (16-byte loads: 4 columns per thread and trip)
        for (int e = tid; e < TJM * (K / 4); e += 256) {
            const int jj = e / (K / 4), c4 = (e % (K / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int j = idx_s[jj];
            if (j >= 0) v = *reinterpret_cast<const float4*>(B + int64_t(j) * k + c4);   // j < 0: none / sample on another shard
            *reinterpret_cast<float4*>(B_s + jj * KP + c4) = v;
        }
        __syncthreads();
        // estimates: each warp takes TJM / 8 sampled rows
        for (int jj = warp; jj < TJM; jj += 8) {
            float d = 0.f;
            for (int c = lane; c < k; c += 32) d = fmaf(a_s[c], B_s[jj * KP + c], d);
            d = warp_sum(d);
            if (lane == 0) {
                float rr = 0.f, ww = 0.f;
                if (idx_s[jj] >= 0) {
                    float est = d, fp = 1.f;
                    if (link == PYCMF_LOGIT) { est = sigmoid_<float>(d); fp = est * (1.f - est); }
                    const float tg = want_g ? t_s[jj] : 0.f;
                    rr = w * (est - tg);
                    ww = w * fp;
                }
                r_s[jj] = rr;
                w_s[jj] = sqrtf(fmaxf(ww, 0.f));
            }
        }
        __syncthreads();
        if (want_g && tid < k) {
#pragma unroll 8
            for (int jj = 0; jj < TJM; jj++) gacc = fmaf(r_s[jj], B_s[jj * KP + tid], gacc);
        }
        // C = sqrt(ww) b, split into tf32 hi / lo once per element
        for (int e = tid; e < TJM * K; e += 256) {
            const int jj = e / K, c = e % K;
            const float x = w_s[jj] * B_s[jj * KP + c];
            const uint32_t h32 = to_tf32(x);
            Chi[jj * KP + c] = h32;
            Clo[jj * KP + c] = to_tf32(x - __uint_as_float(h32));
        }
        __syncthreads();
        // H += C_hi^T C_hi + C_hi^T C_lo + C_lo^T C_hi   (contraction over the TJM samples, 8 per MMA)
#pragma unroll
        for (int kb = 0; kb < TJM; kb += 8) {
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int x = 0; x < MT; x++) {
                const int r0 = (kb + tq) * KP + m0 + x * 16 + gq, r1 = (kb + tq + 4) * KP + m0 + x * 16 + gq;
                ah[x][0] = Chi[r0]; ah[x][1] = Chi[r0 + 8]; ah[x][2] = Chi[r1]; ah[x][3] = Chi[r1 + 8];
                al[x][0] = Clo[r0]; al[x][1] = Clo[r0 + 8]; al[x][2] = Clo[r1]; al[x][3] = Clo[r1 + 8];
            }
#pragma unroll
            for (int y = 0; y < NT; y++) {
                const int c0 = (kb + tq) * KP + n0 + y * 8 + gq, c1 = (kb + tq + 4) * KP + n0 + y * 8 + gq;
                const uint32_t bh0 = Chi[c0], bh1 = Chi[c1], bl0 = Clo[c0], bl1 = Clo[c1];
#pragma unroll
                for (int x = 0; x < MT; x++) {
                    mma_tf32(acc[x][y], al[x], bh0, bh1);        // corrections first, the large term last
                    mma_tf32(acc[x][y], ah[x], bl0, bl1);
                    mma_tf32(acc[x][y], ah[x], bh0, bh1);
                }
            }
        }
    }
    if (want_g && tid < k) {
        const float prev = accumulate ? g[i * k + tid] : 0.f;
        g[i * k + tid] = prev + gacc;
    }
    float* Hi = H + i * int64_t(k) * k;
#pragma unroll
    for (int x = 0; x < MT; x++)
#pragma unroll
        for (int y = 0; y < NT; y++) {
            const int r = m0 + x * 16 + gq, c = n0 + y * 8 + 2 * tq;
            float2* p0 = reinterpret_cast<float2*>(Hi + r * k + c);
            float2* p1 = reinterpret_cast<float2*>(Hi + (r + 8) * k + c);
            float2 v0 = make_float2(acc[x][y][0], acc[x][y][1]), v1 = make_float2(acc[x][y][2], acc[x][y][3]);
            if (accumulate) { const float2 q0 = *p0, q1 = *p1; v0.x += q0.x; v0.y += q0.y; v1.x += q1.x; v1.y += q1.y; }
            *p0 = v0;
            *p1 = v1;
        }
}

// =============================================================================================
// few rows, many samples (the Z update: l label rows against all d rows of V), unsampled, Hessian only.
//   H_i = w * sum_j f'(a_i . b_j) b_j b_j^T      for a GROUP of 16 rows i that share every b_j b_j^T:
// the outer product of a sample is formed once and accumulated into 16 rows' Hessians (64 FMAs per 6 shared-memory
// loads per thread), instead of once per row as in row_grad_hess_kernel (measured instruction-bound there).
// CTA = (row group, sample chunk); thread owns 4 consecutive entries (a, b..b+3) of the 32 x 32 matrices.
// =============================================================================================
constexpr int FR_RG = 16;     // rows per group
constexpr int FR_TS = 64;     // samples per staged tile
constexpr int FR_K = 32;      // max n_components

template <typename T>
__global__ void __launch_bounds__(256)
few_rows_hess_kernel(int64_t rows, int64_t m, int k, const T* __restrict__ A, const T* __restrict__ B, int link, T w,
                     T* __restrict__ Hpart, int64_t samples_per_split) {
    __shared__ __align__(16) T A_s[FR_RG][FR_K + 1];
    __shared__ __align__(16) T B_s[FR_TS][FR_K + 4];     // row pitch 36 floats: 16-byte aligned rows for vector loads
    __shared__ __align__(16) T W_s[FR_TS][FR_RG];
    const int tid = threadIdx.x;
    const int64_t r0 = int64_t(blockIdx.x) * FR_RG;
    const int64_t s_begin = int64_t(blockIdx.y) * samples_per_split;
    const int64_t s_end = s_begin + samples_per_split < m ? s_begin + samples_per_split : m;
    for (int e = tid; e < FR_RG * FR_K; e += 256) {
        int r = e / FR_K, c = e % FR_K;
        A_s[r][c] = (r0 + r < rows && c < k) ? A[(r0 + r) * k + c] : T(0);
    }
    const int ea = tid >> 3, eb = (tid & 7) * 4;          // entries (ea, eb .. eb + 3)
    T acc[4][FR_RG];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int r = 0; r < FR_RG; r++) acc[x][r] = T(0);
    for (int64_t t0 = s_begin; t0 < s_end; t0 += FR_TS) {
        const int cnt = int(s_end - t0 < FR_TS ? s_end - t0 : FR_TS);
        __syncthreads();
        for (int e = tid; e < FR_TS * FR_K; e += 256) {
            int j = e / FR_K, c = e % FR_K;
            B_s[j][c] = (j < cnt && c < k) ? B[(t0 + j) * k + c] : T(0);
        }
        __syncthreads();
        {   // weights: thread -> sample j = tid % 64, rows 4 * (tid / 64) .. + 3
            const int j = tid & 63, rg = (tid >> 6) * 4;
            T d[4] = {T(0), T(0), T(0), T(0)};
            for (int c = 0; c < FR_K; c++) {
                const T b = B_s[j][c];
#pragma unroll
                for (int x = 0; x < 4; x++) d[x] = fma(A_s[rg + x][c], b, d[x]);
            }
#pragma unroll
            for (int x = 0; x < 4; x++) {
                T fp = T(1);
                if (link == PYCMF_LOGIT) { const T sg = sigmoid_<T>(d[x]); fp = sg * (T(1) - sg); }
                W_s[j][rg + x] = (j < cnt && r0 + rg + x < rows) ? w * fp : T(0);
            }
        }
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            const T va = B_s[j][ea];
            T p[4];
#pragma unroll
            for (int x = 0; x < 4; x++) p[x] = va * B_s[j][eb + x];
#pragma unroll
            for (int r = 0; r < FR_RG; r++) {
                const T wr = W_s[j][r];
#pragma unroll
                for (int x = 0; x < 4; x++) acc[x][r] = fma(wr, p[x], acc[x][r]);
            }
        }
    }
    // partial Hessians: Hpart[split][row][a][b]
    T* out = Hpart + int64_t(blockIdx.y) * rows * k * k;
    if (ea < k) {
#pragma unroll
        for (int r = 0; r < FR_RG; r++) {
            if (r0 + r >= rows) continue;
#pragma unroll
            for (int x = 0; x < 4; x++)
                if (eb + x < k) out[(r0 + r) * int64_t(k) * k + ea * k + eb + x] = acc[x][r];
        }
    }
}

// =============================================================================================
// eigenvalue-clamped solve
// =============================================================================================
struct SolveShared {
    int flag;
    int fail;
};

// Load the LOWER triangle of H (row-major, like scipy.linalg.eigh(lower=True)) into W (column-major
// == row-major for a symmetric matrix), adding `diag` on the diagonal.
template <typename T>
__device__ __forceinline__ void load_sym(double* W, const T* __restrict__ H, int k, double diag, double scale,
                                         const double* __restrict__ base = nullptr) {
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
        int r = e / k, c = e % k;
        int hi = r > c ? r : c, lo = r > c ? c : r;
        double v = scale * double(H[hi * k + lo]);
        if (base != nullptr) v += base[hi * k + lo];      // shared float64 part of the Hessian (never rounded to T)
        if (r == c) v += diag;
        W[e] = v;
    }
}

// In-place right-looking Cholesky of the lower triangle of W (element (r,c) at W[r*k+c]).
// Returns false (uniformly) when a pivot is not safely positive.
__device__ bool cholesky_inplace(double* W, int k, SolveShared* sh, double pivot_floor) {
    for (int j = 0; j < k; j++) {
        __syncthreads();
        double piv = W[j * k + j];
        if (!(piv > pivot_floor)) return false;  // uniform: every thread reads the same value
        double inv = 1.0 / sqrt(piv);
        __syncthreads();
        for (int r = j + threadIdx.x; r < k; r += blockDim.x) W[r * k + j] *= inv;
        __syncthreads();
        // W[j][j] now holds piv * inv = sqrt(piv)
        const int rem = k - j - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            int r = j + 1 + e / rem, c = j + 1 + e % rem;
            if (c <= r) W[r * k + c] -= W[r * k + j] * W[c * k + j];
        }
    }
    __syncthreads();
    return true;
}

// Solve L L^T x = b in place on b (length k, shared memory), L in the lower triangle of W.
__device__ void chol_solve(const double* W, int k, double* b) {
    for (int j = 0; j < k; j++) {
        __syncthreads();
        double yj = b[j] / W[j * k + j];
        __syncthreads();
        if (threadIdx.x == 0) b[j] = yj;
        for (int r = j + 1 + threadIdx.x; r < k; r += blockDim.x) b[r] -= W[r * k + j] * yj;
    }
    for (int j = k - 1; j >= 0; j--) {
        __syncthreads();
        double xj = b[j] / W[j * k + j];
        __syncthreads();
        if (threadIdx.x == 0) b[j] = xj;
        for (int r = threadIdx.x; r < j; r += blockDim.x) b[r] -= W[j * k + r] * xj;
    }
    __syncthreads();
}

// One-sided (Hestenes) Jacobi on the columns of W (column p = W[p*k .. p*k+k), W symmetric on entry so
// row-major == column-major).  On exit the columns are mutually orthogonal: W = Q diag(lambda) (up to
// column order / sign), so x = S(H) g = g/p + sum_{sigma_i >= p} (1/sigma_i - 1/p) (w_i.g)/sigma_i^2 w_i.
__device__ void jacobi_clamped_solve(double* W, int k, const double* g, double* x, double* xpart,
                                     SolveShared* sh, double pert, double tol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int kk = (k + 1) & ~1;
    const double skip2 = (1e-3 * pert) * (1e-3 * pert);
    for (int sweep = 0; sweep < 60; sweep++) {
        __syncthreads();
        if (threadIdx.x == 0) sh->flag = 0;
        __syncthreads();
        for (int step = 0; step < kk - 1; step++) {
            // TWO pairs per warp at a time, one per half-warp: the rotation scalars (three float64 rsqrt sequences) are
            // the expensive part of a pair -- they issue for the whole warp whatever the active lanes -- so packing two pairs
            // into one pass halves that cost, and the rsqrt-only formulation replaces two IEEE divisions and two square roots.
            const int half = lane >> 4, hl = lane & 15;
            for (int base = warp * 2; base < kk / 2; base += nwarps * 2) {
                const int mth = base + half;
                int p = 0, q = 0;
                bool valid = mth < kk / 2;
                if (valid) {
                    if (mth == 0) { p = step; q = kk - 1; }
                    else { p = (step + mth) % (kk - 1); q = (step - mth + (kk - 1)) % (kk - 1); }
                    valid = p < k && q < k;
                }
                double* wp = W + p * k;
                double* wq = W + q * k;
                double al = 0.0, be = 0.0, ga = 0.0;
                if (valid)
                    for (int r = hl; r < k; r += 16) {
                        double a = wp[r], b = wq[r];
                        al = fma(a, a, al); be = fma(b, b, be); ga = fma(a, b, ga);
                    }
                // the three butterfly reductions interleaved, inside each half-warp (offsets < 16 stay in the half)
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    const double a2 = __shfl_xor_sync(0xffffffffu, al, o), b2 = __shfl_xor_sync(0xffffffffu, be, o),
                                 g2 = __shfl_xor_sync(0xffffffffu, ga, o);
                    al += a2; be += b2; ga += g2;
                }
                const bool rot = valid && ga != 0.0 && fmax(al, be) >= skip2 && ga * ga > (tol * tol) * (al * be);
                if (rot) {
                    // t = tan(theta) = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (be - al) / (2 ga), written without
                    // divisions: t = 2 ga sgn / (|delta| + sqrt(delta^2 + 4 ga^2)); c = rsqrt(1 + t^2), s = c t
                    const double delta = be - al, g2 = 2.0 * ga;
                    const double h = fma(delta, delta, g2 * g2);
                    const double den = fabs(delta) + h * rsqrt(h);
                    const double rd = rsqrt(den);
                    const double t = ((delta >= 0.0) == (ga > 0.0) ? fabs(g2) : -fabs(g2)) * (rd * rd);
                    const double c = rsqrt(fma(t, t, 1.0)), s = c * t;
                    for (int r = hl; r < k; r += 16) {
                        double a = wp[r], b = wq[r];
                        wp[r] = c * a - s * b;
                        wq[r] = s * a + c * b;
                    }
                    if (hl == 0) sh->flag = 1;
                }
            }
            __syncthreads();
        }
        if (sh->flag == 0) break;
    }
    __syncthreads();
    // accumulate x
    const int per = (k + 31) / 32;  // <= 8
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; u++) acc[u] = 0.0;
    for (int i = warp; i < k; i += nwarps) {
        const double* wi = W + i * k;
        double al = 0.0, dg = 0.0;
        for (int r = lane; r < k; r += 32) { double a = wi[r]; al = fma(a, a, al); dg = fma(a, g[r], dg); }
        al = warp_sum(al); dg = warp_sum(dg);
        double sigma = sqrt(al);
        if (sigma >= pert) {
            double coef = (1.0 / sigma - 1.0 / pert) * dg / al;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int r = lane + 32 * u;
                if (u < per && r < k) acc[u] = fma(coef, wi[r], acc[u]);
            }
        }
    }
    // cross-warp reduction through xpart (nwarps x k)
#pragma unroll
    for (int u = 0; u < 8; u++) {
        int r = lane + 32 * u;
        if (u < per && r < k) xpart[warp * k + r] = acc[u];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
        double s = g[r] / pert;
        for (int wv = 0; wv < nwarps; wv++) s += xpart[wv * k + r];
        x[r] = s;
    }
    __syncthreads();
}

// The whole clamped solve for one matrix; H is (re)loaded from global memory as needed.
template <typename T, int KR = 0>
__device__ void safe_solve_one(double* W, const T* __restrict__ H, int k, double diag, const double* g,
                               double* x, double* xpart, SolveShared* sh, double pert, bool chol_fastpath,
                               double scale, const double* __restrict__ base = nullptr, double* tri_work = nullptr,
                               double* tri_scratch = nullptr) {
    bool done = false;
    if (chol_fastpath) {
        // lambda_min(H) > pert  <=>  H - pert I is positive definite  <=>  its Cholesky succeeds
        load_sym<T>(W, H, k, diag - pert, scale, base);
        __syncthreads();
        double tr = 0.0;
        for (int r = 0; r < k; r++) tr += fabs(W[r * k + r]);   // every thread, same value (k <= 256)
        bool ok = cholesky_inplace(W, k, sh, 1e-13 * (tr + pert));
        __syncthreads();
        if (ok) {
            load_sym<T>(W, H, k, diag, scale, base);
            __syncthreads();
            ok = cholesky_inplace(W, k, sh, 0.0);
            if (ok) {
                for (int r = threadIdx.x; r < k; r += blockDim.x) x[r] = g[r];
                __syncthreads();
                chol_solve(W, k, x);
                done = true;
            }
        }
        __syncthreads();
    }
    if (!done) {
        load_sym<T>(W, H, k, diag, scale, base);
        __syncthreads();
        // Every |lambda_i| <= ||H||_F.  Below the clamp level all eigenvalues are raised to `pert`, so S(H) = I / pert
        // EXACTLY and no decomposition is needed -- the regime of the first iterations of a sampled logit fit (C4), where the
        // weighted Grams of small factors have trace << pert and every row used to pay ten Jacobi sweeps for it.
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
            double f2 = 0.0;
            for (int e = threadIdx.x; e < k * k; e += blockDim.x) f2 = fma(W[e], W[e], f2);
            f2 = warp_sum(f2);
            if (lane == 0) xpart[warp] = f2;
            __syncthreads();
            double tot = 0.0;
            for (int wv = 0; wv < nwarps; wv++) tot += xpart[wv];
            __syncthreads();
            if (tot < pert * pert) {
                for (int r = threadIdx.x; r < k; r += blockDim.x) x[r] = g[r] / pert;
                __syncthreads();
                return;
            }
        }
        // clamp active on part of the spectrum: only the eigenpairs above the clamp level are needed (tridiag_solve.cuh);
        // one-sided Jacobi is the fallback
        if (tri_work != nullptr) {
            if (tri::clamped_solve<KR>(W, k, g, x, tri_work, tri_scratch, pert)) return;
            load_sym<T>(W, H, k, diag, scale, base);
            __syncthreads();
        }
        // rotations stop at |w_p . w_q| <= tol |w_p| |w_q|: float64 inputs to working precision; float32 inputs carry 6e-8
        // relative noise already, 1e-11 leaves the clamped solve exact to far below that and saves the last sweep
        jacobi_clamped_solve(W, k, g, x, xpart, sh, pert, sizeof(T) == 4 ? 1e-11 : 1e-15);
    }
}

// MODE 0: x_b = S(H_b) g_b (all float64).  MODE 1: Newton row update on F.
// KR > 0: k == KR, blockDim.x == 2 KR, the tridiagonalisation keeps the matrix in registers (tridiag_solve.cuh, tridiag_reg).
template <typename T, int MODE, int KR = 0>
__global__ void __launch_bounds__(KR > 0 ? 2 * KR : 1024)
safe_solve_kernel(int64_t batch, int k, const T* __restrict__ H, int64_t h_stride, const T* __restrict__ g,
                  T* __restrict__ out, double l1, double l2, double l2_diag, double pert, bool non_negative,
                  bool chol_fastpath, double* __restrict__ Wglobal, double h_scale,
                  const double* __restrict__ Hbase, double* __restrict__ tri_scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5;
    double* gv = reinterpret_cast<double*>(smem_raw);   // k
    double* xv = gv + k;                                // k
    double* xpart = xv + k;                             // nwarps * k
    SolveShared* sh = reinterpret_cast<SolveShared*>(xpart + nwarps * k);
    double* tri_work = tri_scratch != nullptr ? reinterpret_cast<double*>(sh + 2) : nullptr;
    double* W = Wglobal != nullptr ? Wglobal + int64_t(blockIdx.x) * k * k
                                   : reinterpret_cast<double*>(sh + 2) + (tri_scratch != nullptr ? tri::work_doubles(k) : 0);
    double* tri_zg = tri_scratch != nullptr ? tri_scratch + size_t(blockIdx.x) * tri::scratch_doubles(k) : nullptr;
    for (int64_t b = blockIdx.x; b < batch; b += gridDim.x) {
        __syncthreads();
        for (int r = threadIdx.x; r < k; r += blockDim.x) {
            double gr = double(g[b * k + r]);
            if (MODE == 1) {
                double f = double(out[b * k + r]);
                double sg = f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0);
                gr += l1 * sg + l2 * f;
            }
            gv[r] = gr;
        }
        __syncthreads();
        safe_solve_one<T, KR>(W, H + b * h_stride, k, l2_diag, gv, xv, xpart, sh, pert, chol_fastpath, h_scale, Hbase, tri_work,
                          tri_zg);
        __syncthreads();
        for (int r = threadIdx.x; r < k; r += blockDim.x) {
            if (MODE == 0) {
                out[b * k + r] = T(xv[r]);
            } else {
                double f = double(out[b * k + r]) - xv[r];
                if (non_negative && f < 0.0) f = 0.0;
                out[b * k + r] = T(f);
            }
        }
    }
}

// F (rows x k) <- F - (G + l1 sign F + l2 F) Hinv  (shared inverse, read through L1/L2), optional clamp
template <typename T>
__global__ void __launch_bounds__(256)
apply_shared_inverse_kernel(int64_t rows, int k, T* __restrict__ F, const T* __restrict__ G,
                            const double* __restrict__ Hinv, double l1, double l2, bool non_negative) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* gs = reinterpret_cast<double*>(smem_raw);   // RPB x k
    const int RPB = 8;
    const int64_t r0 = int64_t(blockIdx.x) * RPB;
    for (int e = threadIdx.x; e < RPB * k; e += blockDim.x) {
        int64_t r = r0 + e / k;
        int c = e % k;
        double v = 0.0;
        if (r < rows) {
            double f = double(F[r * k + c]);
            double sg = f > 0.0 ? 1.0 : (f < 0.0 ? -1.0 : 0.0);
            v = double(G[r * k + c]) + l1 * sg + l2 * f;
        }
        gs[e] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < RPB * k; e += blockDim.x) {
        int rr = e / k, c = e % k;
        int64_t r = r0 + rr;
        if (r >= rows) continue;
        double s = 0.0;
        for (int a = 0; a < k; a++) s = fma(gs[rr * k + a], __ldg(&Hinv[a * k + c]), s);
        double f = double(F[r * k + c]) - s;
        if (non_negative && f < 0.0) f = 0.0;
        F[r * k + c] = T(f);
    }
}

__global__ void identity_kernel(int k, double* I) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < k * k) I[e] = (e / k == e % k) ? 1.0 : 0.0;
}

template <typename T>
__global__ void cast_to_f64_kernel(int64_t n, const T* __restrict__ a, double* __restrict__ b) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < n) b[e] = double(a[e]);
}

// =============================================================================================
// sampler
// =============================================================================================
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void sample_indices_kernel(int64_t rows, int64_t row0, int64_t N, int64_t n_sample, uint64_t seed,
                                      uint64_t stream_id, int64_t lo, int64_t hi, int32_t* __restrict__ idx) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= rows * n_sample) return;
    int64_t row = row0 + e / n_sample;          // the key uses the GLOBAL row: every shard count draws the same sets
    uint32_t t = uint32_t(e % n_sample);
    int bits = 2;
    while ((int64_t(1) << bits) < N) bits += 2;   // even number of bits
    const int half = bits / 2;
    const uint32_t hmask = (1u << half) - 1u;
    uint32_t key0 = mix32(uint32_t(seed) ^ mix32(uint32_t(seed >> 32) + 0x9e3779b9U));
    key0 = mix32(key0 ^ uint32_t(stream_id) * 0x85ebca6bU);
    key0 = mix32(key0 ^ uint32_t(row) * 0xc2b2ae35U ^ uint32_t(uint64_t(row) >> 32));
    uint32_t v = t;
    do {
        uint32_t L = v >> half, R = v & hmask;
#pragma unroll
        for (int round = 0; round < 4; round++) {
            uint32_t f = mix32(R ^ key0 ^ (0x9e3779b9U * uint32_t(round + 1))) & hmask;
            uint32_t nl = R, nr = L ^ f;
            L = nl; R = nr;
        }
        v = (L << half) | R;
    } while (int64_t(v) >= N);
    // window [lo, hi) = the part of the population this rank holds (rows of U in the V update): others become -1
    idx[e] = hi > lo ? ((int64_t(v) >= lo && int64_t(v) < hi) ? int32_t(int64_t(v) - lo) : -1) : int32_t(v);
}

size_t solve_smem_bytes(int k, int nthreads, bool w_in_smem, bool tri_path = false) {
    int nwarps = nthreads / 32;
    size_t b = sizeof(double) * (size_t(2) * k + size_t(nwarps) * k) + 2 * sizeof(SolveShared);
    b = (b + 15) & ~size_t(15);
    if (tri_path) b += sizeof(double) * tri::work_doubles(k);
    if (w_in_smem) b += sizeof(double) * size_t(k) * k;
    return b + 16;
}

template <typename T, int MODE>
void launch_solve(pycmf_ctx* ctx, int64_t batch, int64_t k, const T* H, int64_t h_stride, const T* g, T* out,
                  double l1, double l2, double l2_diag, double pert, bool non_negative, double h_scale = 1.0,
                  bool known_pd = false, const double* Hbase = nullptr) {
    if (batch <= 0) return;
    PYCMF_CHECK(k >= 1 && k <= 256, "n_components must be in [1, 256] for the Newton solve");
    PYCMF_CHECK(pert > 0.0, "hessian_pertubation must be > 0");
    if (Hbase == nullptr &&
        safe_solve_small<T, MODE>(ctx, batch, k, H, h_stride, g, out, l1, l2, l2_diag, pert, non_negative, h_scale,
                                  known_pd))
        return;
    // one warp per Jacobi pair: k / 2 pairs per step, so wide matrices get a full CTA (k = 128: 64 pairs on 32 warps are two
    // rounds per step instead of eight on 8 warps)
    // tridiagonal path (default): one thread per wanted eigenvector, so at least k threads; Jacobi only: one warp per column pair
    const bool tri_path = ctx->solve_path != 0 && k >= 8;
    int nthreads = tri_path ? (ctx->solve_threads > 0 ? ctx->solve_threads : (k <= 64 ? 128 : (k <= 128 ? 256 : 512)))
                            : (k <= 32 ? 64 : (k <= 64 ? 128 : (k <= 96 ? 256 : 1024)));
    bool w_in_smem = solve_smem_bytes(int(k), nthreads, true, tri_path) <= size_t(ctx->max_smem_optin);
    size_t smem = solve_smem_bytes(int(k), nthreads, w_in_smem, tri_path);
    auto kern = safe_solve_kernel<T, MODE>;
    // solve_path 2: Householder steps on a register-resident matrix (k = 64 / 128 with the matrix in shared memory and 2 k threads)
    if (tri_path && ctx->solve_path >= 2 && w_in_smem && nthreads == 2 * k) {
        if (k == 64) kern = safe_solve_kernel<T, MODE, 64>;
        else if (k == 128) kern = safe_solve_kernel<T, MODE, 128>;
    }
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t grid = batch;
    double* Wg = nullptr;
    double* tri_scratch = nullptr;
    if (!w_in_smem || tri_path) grid = std::min<int64_t>(batch, 2 * ctx->num_sms);      // persistent: per-CTA scratch
    else grid = std::min<int64_t>(batch, int64_t(1) << 30);
    if (!w_in_smem) Wg = static_cast<double*>(scratch(ctx, 2, size_t(grid) * k * k * sizeof(double)));
    if (tri_path) tri_scratch = static_cast<double*>(scratch(ctx, 8, size_t(grid) * tri::scratch_doubles(int(k)) * sizeof(double)));
    Timed timer(ctx, "safe_solve");
    kern<<<(unsigned)grid, nthreads, smem, ctx->stream>>>(batch, int(k), H, h_stride, g, out, l1, l2, l2_diag,
                                                         pert, non_negative, ctx->chol_fastpath != 0, Wg, h_scale, Hbase,
                                                         tri_scratch);
    PYCMF_LAUNCH_CHECK(ctx);
}

}  // namespace

template <typename T>
void row_grad_hess(pycmf_ctx* ctx, int64_t rows, int64_t m, int64_t k, const T* A, const T* B,
                   const T* Tgt, int64_t ldt, bool trans_t,
                   const int32_t* rowptr, const int32_t* colidx, const T* vals,
                   int link, double w, const int32_t* idx, int64_t n_sample,
                   T* g, T* H, bool accumulate) {
    if (rows <= 0) return;
    PYCMF_CHECK(k >= 1 && k <= 256, "n_components must be in [1, 256] for the per-row Newton kernels");
    if (idx == nullptr && g == nullptr && H != nullptr && k <= FR_K && rows <= 4 * FR_RG && m >= 1024) {
        // few rows against every row of B (Z update): grouped-row kernel, sample-split, deterministic reduce
        const int64_t groups = ceil_div(rows, FR_RG);
        int64_t nsplit = std::max<int64_t>(1, std::min(ceil_div(m, FR_TS), ceil_div(int64_t(2) * ctx->num_sms, groups)));
        const int64_t per = ceil_div(ceil_div(m, nsplit), FR_TS) * FR_TS;
        nsplit = ceil_div(m, per);
        T* part = static_cast<T*>(scratch(ctx, 0, size_t(nsplit) * rows * k * k * sizeof(T)));
        {
            Timed timer(ctx, "row_grad_hess");
            few_rows_hess_kernel<T><<<dim3((unsigned)groups, (unsigned)nsplit), 256, 0, ctx->stream>>>(
                rows, m, int(k), A, B, link, T(w), part, per);
            PYCMF_LAUNCH_CHECK(ctx);
        }
        reduce_parts<T>(ctx, rows, k * k, int(nsplit), part, H, k * k, T(1), accumulate ? T(1) : T(0));
        return;
    }
    size_t smem = sizeof(T) * (size_t((k + 3) & ~3) + size_t(TJ) * (k + 1) + 3 * TJ);
    int hb = k <= 16 ? 1 : (k <= 32 ? 2 : (k <= 64 ? 4 : 8));
    int quads = k > 128 ? 2 : 1;
    // few rows, many samples (e.g. the Z update: l rows against d samples): split the samples over CTAs
    const int64_t total = idx != nullptr ? n_sample : m;
    int64_t nsplit = 1;
    if (quads == 1 && rows < 2 * ctx->num_sms && total >= 8 * TJ)
        nsplit = std::max<int64_t>(1, std::min(ceil_div(total, 4 * TJ), ceil_div(int64_t(4) * ctx->num_sms, rows)));
    int64_t per_split = ceil_div(ceil_div(std::max<int64_t>(total, 1), nsplit), TJ) * TJ;
    nsplit = std::max<int64_t>(1, ceil_div(std::max<int64_t>(total, 1), per_split));
    T *g_out = g, *H_out = H;
    int64_t g_stride = 0, h_stride = 0;
    bool acc_kernel = accumulate;
    if (nsplit > 1) {
        size_t gbytes = g ? size_t(nsplit) * rows * k * sizeof(T) : 0, hbytes = H ? size_t(nsplit) * rows * k * k * sizeof(T) : 0;
        gbytes = (gbytes + 255) & ~size_t(255);
        unsigned char* part = static_cast<unsigned char*>(scratch(ctx, 0, gbytes + hbytes + 256));
        if (g) { g_out = reinterpret_cast<T*>(part); g_stride = rows * k; }
        if (H) { H_out = reinterpret_cast<T*>(part + gbytes); h_stride = rows * k * k; }
        acc_kernel = false;
    }
    if constexpr (std::is_same<T, float>::value) {
        // tensor-core weighted Gram (mma.sync 3xTF32) for the wide factors of the sampled / logit Newton step
        if (H != nullptr && (k == 64 || k == 128) && w >= 0.0 && ctx->hess_mma != 0 &&
            (reinterpret_cast<uintptr_t>(B) & 15) == 0 && (reinterpret_cast<uintptr_t>(H_out) & 7) == 0) {
            const size_t sm = sizeof(float) * (size_t(k) + size_t(3) * TJM * (k + 8) + 4 * TJM + 2 * RCM);
            dim3 grid((unsigned)rows, (unsigned)nsplit);
            Timed timer(ctx, "row_grad_hess");
#define LAUNCHM(KK)                                                                                              \
    do {                                                                                                         \
        auto kern = row_grad_hess_mma_kernel<KK>;                                                                \
        PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));            \
        kern<<<grid, 256, sm, ctx->stream>>>(rows, m, A, B, Tgt, ldt, trans_t, rowptr, colidx, vals, link,       \
                                             float(w), idx, n_sample, g_out, H_out, acc_kernel, per_split,        \
                                             g_stride, h_stride);                                                \
    } while (0)
            if (k == 64) LAUNCHM(64); else LAUNCHM(128);
#undef LAUNCHM
            PYCMF_LAUNCH_CHECK(ctx);
            if (nsplit > 1) {
                if (g) reduce_parts<T>(ctx, rows, k, int(nsplit), g_out, g, k, T(1), accumulate ? T(1) : T(0));
                reduce_parts<T>(ctx, rows, k * k, int(nsplit), H_out, H, k * k, T(1), accumulate ? T(1) : T(0));
            }
            return;
        }
    }
    {
        Timed timer(ctx, "row_grad_hess");
        for (int qa = 0; qa < quads; qa++) {
            for (int qb = 0; qb < quads; qb++) {
                bool first = (qa == 0 && qb == 0);
                T* gq = first ? g_out : nullptr;
                if (!first && H == nullptr) continue;
                dim3 grid((unsigned)rows, (unsigned)nsplit);
#define LAUNCH(HB)                                                                                        \
    do {                                                                                                  \
        auto kern = row_grad_hess_kernel<T, HB>;                                                          \
        PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));   \
        kern<<<grid, 256, smem, ctx->stream>>>(rows, m, int(k), A, B, Tgt, ldt, trans_t, rowptr, colidx,  \
                                               vals, link, T(w), idx, n_sample, gq, H_out, acc_kernel,    \
                                               qa * 128, qb * 128, per_split, g_stride, h_stride);        \
    } while (0)
                if (hb == 1) LAUNCH(1);
                else if (hb == 2) LAUNCH(2);
                else if (hb == 4) LAUNCH(4);
                else LAUNCH(8);
#undef LAUNCH
                PYCMF_LAUNCH_CHECK(ctx);
            }
        }
    }
    if (nsplit > 1) {
        if (g) reduce_parts<T>(ctx, rows, k, int(nsplit), g_out, g, k, T(1), accumulate ? T(1) : T(0));
        if (H) reduce_parts<T>(ctx, rows, k * k, int(nsplit), H_out, H, k * k, T(1), accumulate ? T(1) : T(0));
    }
}

void safe_solve_f64(pycmf_ctx* ctx, int64_t batch, int64_t k, const double* H, int64_t h_stride,
                    const double* g, double* x, double pert) {
    launch_solve<double, 0>(ctx, batch, k, H, h_stride, g, x, 0.0, 0.0, 0.0, pert, false);
}

template <typename T>
void newton_solve_rows(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* g, const T* H, int64_t h_stride,
                       double l1, double l2, double l2_diag, double pert, bool non_negative, bool known_pd,
                       const double* Hbase) {
    if (rows <= 0) return;
    if (h_stride != 0) {
        launch_solve<T, 1>(ctx, rows, k, H, h_stride, g, F, l1, l2, l2_diag, pert, non_negative, 1.0, known_pd, Hbase);
        return;
    }
    PYCMF_CHECK(Hbase == nullptr, "newton_solve_rows: a float64 base goes with per-row Hessians only");
    // shared Hessian: invert once (k unit right-hand sides), then one small GEMM-like pass over the rows
    double* buf = static_cast<double*>(scratch(ctx, 3, sizeof(double) * size_t(3) * k * k));
    double *H64 = buf, *I64 = buf + k * k, *Hinv = buf + 2 * k * k;
    int64_t n = k * k;
    cast_to_f64_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(n, H, H64);
    PYCMF_LAUNCH_CHECK(ctx);
    identity_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(int(k), I64);
    PYCMF_LAUNCH_CHECK(ctx);
    // column c of Hinv = S(H + l2_diag I) e_c ; S symmetric, so rows of the result are its columns
    launch_solve<double, 0>(ctx, k, k, H64, 0, I64, Hinv, 0.0, 0.0, l2_diag, pert, false);
    size_t smem = sizeof(double) * size_t(8) * k;
    auto kern = apply_shared_inverse_kernel<T>;
    kern<<<(unsigned)ceil_div(rows, 8), 256, smem, ctx->stream>>>(rows, int(k), F, g, Hinv, l1, l2, non_negative);
    PYCMF_LAUNCH_CHECK(ctx);
}

// Hinv (k x k float64, in arena 7 of ctx) = S(h_scale * G64 + l2_diag I): k unit right-hand sides, one warp each.
// Split from the application so that a caller can run it on a side stream next to the gradient pass.
double* shared_inverse64(pycmf_ctx* ctx, int64_t k, const double* G64, double h_scale, double l2_diag, double pert) {
    double* buf = static_cast<double*>(scratch(ctx, 7, sizeof(double) * size_t(2) * k * k));
    double *I64 = buf, *Hinv = buf + k * k;
    int64_t n = k * k;
    identity_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(int(k), I64);
    PYCMF_LAUNCH_CHECK(ctx);
    // column c of Hinv = S(.) e_c ; S symmetric, so rows of the result are its columns
    launch_solve<double, 0>(ctx, k, k, G64, 0, I64, Hinv, 0.0, 0.0, l2_diag, pert, false, h_scale);
    return Hinv;
}

template <typename T>
void apply_shared_inverse(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* g, const double* Hinv, double l1,
                          double l2, bool non_negative) {
    if (rows <= 0) return;
    size_t smem = sizeof(double) * size_t(8) * k;
    Timed timer(ctx, "apply_shared_inverse");
    apply_shared_inverse_kernel<T><<<(unsigned)ceil_div(rows, 8), 256, smem, ctx->stream>>>(rows, int(k), F, g, Hinv, l1, l2,
                                                                                       non_negative);
    PYCMF_LAUNCH_CHECK(ctx);
}

void sample_indices(pycmf_ctx* ctx, int64_t rows, int64_t row0, int64_t N, int64_t n_sample, uint64_t seed,
                    uint64_t stream_id, int64_t lo, int64_t hi, int32_t* idx) {
    int64_t n = rows * n_sample;
    if (n <= 0) return;
    PYCMF_CHECK(n_sample <= N, "cannot sample more indices than the population");
    PYCMF_CHECK(N < (int64_t(1) << 31), "population too large for int32 indices");
    sample_indices_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(rows, row0, N, n_sample, seed, stream_id, lo, hi, idx);
    PYCMF_LAUNCH_CHECK(ctx);
}

#define INSTANTIATE(T)                                                                                          \
    template void row_grad_hess<T>(pycmf_ctx*, int64_t, int64_t, int64_t, const T*, const T*, const T*, int64_t, \
                                   bool, const int32_t*, const int32_t*, const T*, int, double, const int32_t*, \
                                   int64_t, T*, T*, bool);                                                      \
    template void newton_solve_rows<T>(pycmf_ctx*, int64_t, int64_t, T*, const T*, const T*, int64_t, double,   \
                                       double, double, double, bool, bool, const double*);                         \
    template void apply_shared_inverse<T>(pycmf_ctx*, int64_t, int64_t, T*, const T*, const double*, double,    \
                                          double, bool);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace pycmf

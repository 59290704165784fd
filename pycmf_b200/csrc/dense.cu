// Generic dense kernels (fp32 / fp64 FMA pipes): tiled GEMM with split-K, MU ratio epilogue,
// transpose, axpby, dot.  These are the exact-precision path; the tcgen05 TF32 kernels in
// tc_dense.cu take over the large fp32 contractions when ctx->dense_path != 0.
#include <type_traits>

#include "common.cuh"

namespace pycmf {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename T> __device__ __forceinline__ void load4(const T* p, T (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<double>(const double* p, double (&v)[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// C_part[z] (m x q) = op(A)[:, pz0:pz1] * B[pz0:pz1, :]
template <typename T, bool TRANS_A>
__global__ void __launch_bounds__(256)
gemm_kernel(int64_t m, int64_t q, int64_t p, int64_t p_per_split,
            const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb,
            T* __restrict__ C, int64_t ldc, int64_t c_split_stride, T alpha, T beta, bool direct) {
    __shared__ __align__(16) T As[BK][BM + PAD];
    __shared__ __align__(16) T Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = int64_t(blockIdx.y) * BM, col0 = int64_t(blockIdx.x) * BN;
    const int64_t pz0 = int64_t(blockIdx.z) * p_per_split;
    const int64_t pz1 = min(p, pz0 + p_per_split);

    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = T(0);

    for (int64_t k0 = pz0; k0 < pz1; k0 += BK) {
        // ---- stage A tile as As[kk][row]
        if (TRANS_A) {
            // A stored (p x m): element op(A)[row][kk] = A[kk*lda + row]; coalesced along row
#pragma unroll
            for (int it = 0; it < (BM * BK) / 256; it++) {
                int e = tid + it * 256;
                int kk = e / BM, r = e % BM;
                int64_t gk = k0 + kk, gr = row0 + r;
                As[kk][r] = (gk < pz1 && gr < m) ? A[gk * lda + gr] : T(0);
            }
        } else {
            // A stored (m x p): coalesced along kk
#pragma unroll
            for (int it = 0; it < (BM * BK) / 256; it++) {
                int e = tid + it * 256;
                int r = e / BK, kk = e % BK;
                int64_t gk = k0 + kk, gr = row0 + r;
                As[kk][r] = (gk < pz1 && gr < m) ? A[gr * lda + gk] : T(0);
            }
        }
#pragma unroll
        for (int it = 0; it < (BN * BK) / 256; it++) {
            int e = tid + it * 256;
            int kk = e / BN, c = e % BN;
            int64_t gk = k0 + kk, gc = col0 + c;
            Bs[kk][c] = (gk < pz1 && gc < q) ? B[gk * ldb + gc] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            T a[4], b[4];
            load4<T>(&As[kk][ty * 4], a);
            load4<T>(&Bs[kk][tx * 4], b);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    T* Cz = C + int64_t(blockIdx.z) * c_split_stride;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int64_t gr = row0 + ty * 4 + i;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t gc = col0 + tx * 4 + j;
            if (gc >= q) continue;
            if (direct) {
                T prev = beta != T(0) ? Cz[gr * ldc + gc] : T(0);
                Cz[gr * ldc + gc] = alpha * acc[i][j] + beta * prev;
            } else {
                Cz[gr * ldc + gc] = acc[i][j];
            }
        }
    }
}

template <typename T>
__global__ void splitk_reduce_kernel(int64_t m, int64_t q, int splits, const T* __restrict__ part,
                                     T* __restrict__ C, int64_t ldc, T alpha, T beta) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= m * q) return;
    T s = T(0);
    for (int z = 0; z < splits; z++) s += part[int64_t(z) * m * q + e];
    int64_t r = e / q, c = e % q;
    T prev = beta != T(0) ? C[r * ldc + c] : T(0);
    C[r * ldc + c] = alpha * s + beta * prev;
}

template <typename T>
__global__ void mu_apply_kernel(int64_t n, T* __restrict__ F, const T* __restrict__ N,
                                const T* __restrict__ D, T l1, T l2, T eps) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    T f = F[e];
    T den = D[e];
    if (l1 > T(0)) den += l1;
    if (l2 > T(0)) den = den + l2 * f;
    if (den == T(0)) den = eps;
    F[e] = f * (N[e] / den);
}

template <typename T>
__global__ void transpose_kernel(int64_t rows, int64_t cols, const T* __restrict__ A, int64_t lda,
                                 T* __restrict__ At, int64_t ldat) {
    __shared__ T tile[32][33];
    int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? A[r * lda + c] : T(0);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) At[c * ldat + r] = tile[threadIdx.x][i];
    }
}

template <typename T>
__global__ void axpby_kernel(int64_t n, T alpha, const T* __restrict__ a, T beta, const T* __restrict__ b,
                             T* __restrict__ out) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n) return;
    T v = alpha * a[e];
    if (b != nullptr) v += beta * b[e];
    out[e] = v;
}

template <typename T>
__global__ void dot_partial_kernel(int64_t n, const T* __restrict__ a, const T* __restrict__ b,
                                   double* __restrict__ part) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x)
        s += double(a[e]) * double(b[e]);
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void final_sum_kernel(int nparts, const double* __restrict__ part, double scale, double* out,
                                 bool accumulate) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = (accumulate ? *out : 0.0) + scale * s;
}


// H[i] (kk elements each) = (overwrite ? 0 : H[i]) + scale * Hs
template <typename T>
__global__ void broadcast_add_kernel(int64_t rows, int64_t kk, T* __restrict__ H, const T* __restrict__ Hs,
                                     T scale, bool overwrite) {
    int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= rows * kk) return;
    T v = scale * Hs[e % kk];
    H[e] = overwrite ? v : H[e] + v;
}

// Gram matrix with float64 accumulation: part[blockIdx] (k x k doubles) = sum over the CTA's rows a_r a_r^T
template <typename T, int HB>
__global__ void __launch_bounds__(256)
gram_f64_kernel(int64_t rows, int k, const T* __restrict__ A, double* __restrict__ part, int a_off, int b_off) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* A_s = reinterpret_cast<T*>(smem_raw);  // 32 x (k+1)
    const int kp = k + 1;
    const int tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;
    double acc[HB][HB];
#pragma unroll
    for (int x = 0; x < HB; x++)
#pragma unroll
        for (int y = 0; y < HB; y++) acc[x][y] = 0.0;
    for (int64_t r0 = int64_t(blockIdx.x) * 32; r0 < rows; r0 += int64_t(gridDim.x) * 32) {
        __syncthreads();
        for (int e = tid; e < 32 * k; e += 256) {
            int rr = e / k, c = e % k;
            A_s[rr * kp + c] = (r0 + rr < rows) ? A[(r0 + rr) * k + c] : T(0);
        }
        __syncthreads();
        for (int rr = 0; rr < 32; rr++) {
            double av[HB], bv[HB];
#pragma unroll
            for (int x = 0; x < HB; x++) { int a = a_off + ta + 16 * x; av[x] = a < k ? double(A_s[rr * kp + a]) : 0.0; }
#pragma unroll
            for (int y = 0; y < HB; y++) { int b = b_off + tb + 16 * y; bv[y] = b < k ? double(A_s[rr * kp + b]) : 0.0; }
#pragma unroll
            for (int x = 0; x < HB; x++)
#pragma unroll
                for (int y = 0; y < HB; y++) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
        }
    }
    double* P = part + int64_t(blockIdx.x) * k * k;
#pragma unroll
    for (int x = 0; x < HB; x++) {
        int a = a_off + ta + 16 * x;
        if (a >= k) continue;
#pragma unroll
        for (int y = 0; y < HB; y++) {
            int b = b_off + tb + 16 * y;
            if (b < k) P[a * k + b] = acc[x][y];
        }
    }
}

}  // namespace

void final_sum(pycmf_ctx* ctx, int nparts, const double* part, double scale, double* out, bool accumulate) {
    final_sum_kernel<<<1, 256, 0, ctx->stream>>>(nparts, part, scale, out, accumulate);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void gemm(pycmf_ctx* ctx, bool trans_a, int64_t m, int64_t q, int64_t p, const T* A, int64_t lda,
          const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta) {
    if (m <= 0 || q <= 0) return;
    if constexpr (std::is_same<T, double>::value) {
        if (p > 0 && dmma_gemm_eligible(ctx, m, q, p, A, lda, B, ldb)) {
            dmma_gemm(ctx, trans_a, m, q, p, A, lda, B, ldb, C, ldc, alpha, beta);
            return;
        }
    }
    Timed timer(ctx, "gemm");
    int64_t tiles = ceil_div(m, BM) * ceil_div(q, BN);
    int splits = 1;
    if (p > 0 && tiles < 2 * ctx->num_sms) {
        int64_t want = ceil_div(int64_t(4) * ctx->num_sms, tiles);
        int64_t maxs = ceil_div(p, 256);
        splits = int(std::max<int64_t>(1, std::min(want, maxs)));
    }
    int64_t p_per = ceil_div(std::max<int64_t>(p, 1), splits);
    p_per = ceil_div(p_per, BK) * BK;
    splits = int(std::max<int64_t>(1, ceil_div(std::max<int64_t>(p, 1), p_per)));
    dim3 grid((unsigned)ceil_div(q, BN), (unsigned)ceil_div(m, BM), (unsigned)splits);
    PYCMF_CHECK(grid.y <= 65535u * 32u, "gemm: too many row tiles");
    if (grid.y > 65535u) {
        // fold: process in row chunks
        int64_t chunk = int64_t(65535) * BM;
        for (int64_t r = 0; r < m; r += chunk) {
            int64_t mm = std::min(chunk, m - r);
            gemm<T>(ctx, trans_a, mm, q, p, trans_a ? A + r : A + r * lda, lda, B, ldb, C + r * ldc, ldc, alpha, beta);
        }
        return;
    }
    if (splits == 1) {
        if (trans_a)
            gemm_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(m, q, p, p_per, A, lda, B, ldb, C, ldc, 0, alpha, beta, true);
        else
            gemm_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(m, q, p, p_per, A, lda, B, ldb, C, ldc, 0, alpha, beta, true);
        PYCMF_LAUNCH_CHECK(ctx);
    } else {
        T* part = static_cast<T*>(scratch(ctx, 0, size_t(splits) * m * q * sizeof(T)));
        if (trans_a)
            gemm_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(m, q, p, p_per, A, lda, B, ldb, part, q, m * q, T(1), T(0), false);
        else
            gemm_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(m, q, p, p_per, A, lda, B, ldb, part, q, m * q, T(1), T(0), false);
        PYCMF_LAUNCH_CHECK(ctx);
        int64_t n = m * q;
        splitk_reduce_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(m, q, splits, part, C, ldc, alpha, beta);
        PYCMF_LAUNCH_CHECK(ctx);
    }
}

template <typename T>
void reduce_parts(pycmf_ctx* ctx, int64_t m, int64_t q, int splits, const T* part, T* C, int64_t ldc, T alpha, T beta) {
    int64_t n = m * q;
    if (n <= 0) return;
    splitk_reduce_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(m, q, splits, part, C, ldc, alpha, beta);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void mu_apply(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* N, const T* D, double l1, double l2) {
    int64_t n = rows * k;
    if (n <= 0) return;
    mu_apply_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(n, F, N, D, T(l1), T(l2), T(kEpsF32));
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void transpose(pycmf_ctx* ctx, int64_t rows, int64_t cols, const T* A, int64_t lda, T* At, int64_t ldat) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    PYCMF_CHECK(grid.y <= 65535u, "transpose: too many rows (tile over rows)");
    transpose_kernel<T><<<grid, dim3(32, 8), 0, ctx->stream>>>(rows, cols, A, lda, At, ldat);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void axpby(pycmf_ctx* ctx, int64_t n, T alpha, const T* a, T beta, const T* b, T* out) {
    if (n <= 0) return;
    axpby_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(n, alpha, a, beta, b, out);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void dot_f64(pycmf_ctx* ctx, int64_t n, const T* a, const T* b, double scale, double* out, bool accumulate) {
    int blocks = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(std::max<int64_t>(n, 1), 1024), 4 * ctx->num_sms)));
    double* part = static_cast<double*>(scratch(ctx, 1, size_t(blocks) * sizeof(double)));
    dot_partial_kernel<T><<<blocks, 256, 0, ctx->stream>>>(n, a, b, part);
    PYCMF_LAUNCH_CHECK(ctx);
    final_sum(ctx, blocks, part, scale, out, accumulate);
}


template <typename T>
void broadcast_add(pycmf_ctx* ctx, int64_t rows, int64_t kk, T* H, const T* Hs, T scale, bool overwrite) {
    int64_t n = rows * kk;
    if (n <= 0) return;
    broadcast_add_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(rows, kk, H, Hs, scale, overwrite);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void gram_f64(pycmf_ctx* ctx, int64_t rows, int64_t k, const T* A, double* G) {
    PYCMF_CHECK(k >= 1 && k <= 256, "n_components must be in [1, 256]");
    int blocks = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(std::max<int64_t>(rows, 1), 128), 2 * ctx->num_sms)));
    double* part = static_cast<double*>(scratch(ctx, 0, size_t(blocks) * k * k * sizeof(double)));
    size_t smem = sizeof(T) * size_t(32) * (k + 1);
    int hb = k <= 16 ? 1 : (k <= 32 ? 2 : (k <= 64 ? 4 : 8));
    int quads = k > 128 ? 2 : 1;
    for (int qa = 0; qa < quads; qa++)
        for (int qb = 0; qb < quads; qb++) {
#define LAUNCH(HB) gram_f64_kernel<T, HB><<<blocks, 256, smem, ctx->stream>>>(rows, int(k), A, part, qa * 128, qb * 128)
            if (hb == 1) LAUNCH(1);
            else if (hb == 2) LAUNCH(2);
            else if (hb == 4) LAUNCH(4);
            else LAUNCH(8);
#undef LAUNCH
            PYCMF_LAUNCH_CHECK(ctx);
        }
    reduce_parts<double>(ctx, k, k, blocks, part, G, k, 1.0, 0.0);
}

// ---- fused MU factor update ----------------------------------------------------------------------------------------------
//     F <- F * N / (F G + l1 + l2 F),  zero denominators -> float32 eps          (cmf_solvers.py:212-228, :233, :239, :245)
// with the denominator product F G (G = k x k Gram sum, k <= 128) formed inside the kernel: a CTA stages G and a block of
// MUF_ROWS rows of F in shared memory, every thread builds four entries of its rows of F G in registers and applies the ratio
// and the zero guard before anything is written.  Replaces a (rows x k x k) GEMM launch, the rows x k denominator round trip
// through HBM (write + read) and the separate elementwise launch: the update reads F and N once and writes F once.
constexpr int MUF_ROWS = 64;

template <typename T> struct Vec4 { T v[4]; };
template <typename T> __device__ __forceinline__ Vec4<T> ld4s(const T* p) {      // 4 consecutive elements, 16-byte aligned for float
    Vec4<T> o;
    load4<T>(p, o.v);
    return o;
}
template <typename T> __device__ __forceinline__ void st4s(T* p, const Vec4<T>& v);
template <> __device__ __forceinline__ void st4s<float>(float* p, const Vec4<float>& v) {
    *reinterpret_cast<float4*>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
}
template <> __device__ __forceinline__ void st4s<double>(double* p, const Vec4<double>& v) {
    *reinterpret_cast<double2*>(p) = make_double2(v.v[0], v.v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v.v[2], v.v[3]);
}

// VEC4: k % 4 == 0 -- a thread owns a 4 x 4 block of F G (4 rows, 4 columns) and walks the contraction four steps at a
// time with 128-bit shared-memory reads of both operands (16 FMAs per 2 loads); otherwise the scalar form (k = 10 on C1).
// KC > 0: n_components known at compile time (32 / 64 / 128): the index divisions of the staging loops become shifts and the
// contraction unrolls (they were 40 % of the instructions with a run-time k).
template <typename T, bool VEC4, int KC>
__global__ void __launch_bounds__(256)
mu_fused_kernel(int64_t rows, int k_rt, T* __restrict__ F, const T* __restrict__ N, const T* __restrict__ G, T l1, T l2,
                T eps) {
    const int k = KC > 0 ? KC : k_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ldg = VEC4 ? k + 4 : k + 1;                    // pitch of G: 16-byte aligned rows / odd (conflict-free columns)
    const int ldf = VEC4 ? k + 4 : k;
    T* Gs = reinterpret_cast<T*>(smem_raw);                  // k x ldg
    T* Fs = Gs + size_t(k) * ldg;                            // MUF_ROWS x ldf
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) Gs[(e / k) * ldg + e % k] = G[e];
    const int cgroups = (k + 3) / 4;                         // column groups of 4
    for (int64_t r0 = int64_t(blockIdx.x) * MUF_ROWS; r0 < rows; r0 += int64_t(gridDim.x) * MUF_ROWS) {
        __syncthreads();
        const int nr = int(min(int64_t(MUF_ROWS), rows - r0));
        if (VEC4) {                                          // 16-byte (float) / 2 x 16-byte (double) loads and stores
            for (int e = threadIdx.x; e < MUF_ROWS * cgroups; e += blockDim.x) {
                const int r = e / cgroups, c = (e % cgroups) * 4;
                Vec4<T> f;
#pragma unroll
                for (int w = 0; w < 4; w++) f.v[w] = T(0);
                if (r < nr) f = ld4s<T>(F + (r0 + r) * k + c);
                st4s<T>(Fs + r * ldf + c, f);
            }
        } else {
            for (int e = threadIdx.x; e < MUF_ROWS * k; e += blockDim.x) {
                const int r = e / k, c = e % k;
                Fs[r * ldf + c] = r < nr ? F[(r0 + r) * k + c] : T(0);
            }
        }
        __syncthreads();
        if (VEC4) {
            for (int item = threadIdx.x; item < (MUF_ROWS / 4) * cgroups; item += blockDim.x) {
                const int rg = item / cgroups, c0 = (item % cgroups) * 4;
                // the numerator entries of this 4 x 4 block are requested before the contraction: their HBM latency
                // hides behind the k FMAs per entry instead of following them
                Vec4<T> nv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
#pragma unroll
                    for (int w = 0; w < 4; w++) nv[u].v[w] = T(0);
                    if (rg * 4 + u < nr) nv[u] = ld4s<T>(N + (r0 + rg * 4 + u) * k + c0);
                }
                T d[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int w = 0; w < 4; w++) d[u][w] = T(0);
                const T* f0 = Fs + (rg * 4) * ldf;
                for (int j = 0; j < k; j += 4) {
                    Vec4<T> f[4], g[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) f[u] = ld4s<T>(f0 + u * ldf + j);
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) g[jj] = ld4s<T>(Gs + (j + jj) * ldg + c0);
#pragma unroll
                    for (int jj = 0; jj < 4; jj++)
#pragma unroll
                        for (int u = 0; u < 4; u++)
#pragma unroll
                            for (int w = 0; w < 4; w++) d[u][w] = fma(f[u].v[jj], g[jj].v[w], d[u][w]);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int r = rg * 4 + u;
                    if (r >= nr) break;
                    const Vec4<T> fv = ld4s<T>(f0 + u * ldf + c0);
                    Vec4<T> o;
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        T dd = d[u][w];
                        if (l1 > T(0)) dd += l1;
                        if (l2 > T(0)) dd = dd + l2 * fv.v[w];
                        if (dd == T(0)) dd = eps;
                        o.v[w] = fv.v[w] * (nv[u].v[w] / dd);
                    }
                    st4s<T>(F + (r0 + r) * k + c0, o);
                }
            }
        } else {
            for (int item = threadIdx.x; item < nr * cgroups; item += blockDim.x) {
                const int r = item / cgroups, c0 = (item % cgroups) * 4;
                T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);
                const T* fr = Fs + r * ldf;
                const int c1 = min(c0 + 1, k - 1), c2 = min(c0 + 2, k - 1), c3 = min(c0 + 3, k - 1);
                for (int j = 0; j < k; j++) {
                    const T f = fr[j];
                    const T* g = Gs + j * ldg;
                    d0 = fma(f, g[c0], d0); d1 = fma(f, g[c1], d1); d2 = fma(f, g[c2], d2); d3 = fma(f, g[c3], d3);
                }
                const T den[4] = {d0, d1, d2, d3};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = c0 + u;
                    if (c >= k) break;
                    const T f = fr[c];
                    T dd = den[u];
                    if (l1 > T(0)) dd += l1;
                    if (l2 > T(0)) dd = dd + l2 * f;
                    if (dd == T(0)) dd = eps;
                    F[(r0 + r) * k + c] = f * (N[(r0 + r) * k + c] / dd);
                }
            }
        }
    }
}

template <typename T>
bool mu_fused_apply(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* N, const T* G, double l1, double l2) {
    if (k > 128 || rows <= 0 || ctx->mu_fused == 0) return false;
    const bool vec = k % 4 == 0;
    const size_t smem = sizeof(T) * (size_t(k) * (k + 4) + size_t(MUF_ROWS) * (k + 4));
    if (smem > size_t(ctx->max_smem_optin)) return false;
    auto kern = k == 32 ? mu_fused_kernel<T, true, 32>
              : k == 64 ? mu_fused_kernel<T, true, 64>
              : k == 128 ? mu_fused_kernel<T, true, 128>
              : vec ? mu_fused_kernel<T, true, 0> : mu_fused_kernel<T, false, 0>;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int per_sm = 2;                                            // one resident wave: the row loop is grid-strided
    PYCMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    const int64_t grid = std::min<int64_t>(ceil_div(rows, MUF_ROWS), int64_t(std::max(per_sm, 1)) * ctx->num_sms);
    Timed timer(ctx, "mu_fused");
    kern<<<(unsigned)grid, 256, smem, ctx->stream>>>(rows, int(k), F, N, G, T(l1), T(l2), T(kEpsF32));
    PYCMF_LAUNCH_CHECK(ctx);
    return true;
}

// ---- top-k per column (topic-term extraction, reference analysis.py:1-16) -----------------------------------------------
// One CTA per column, topn selection rounds: round r finds the largest (value, index) pair below the pair picked in round
// r - 1 (pairs ordered by value, then index), i.e. a selection sort on the top end only -- topn x rows / 256 comparisons per
// thread, no scratch.  The pick of round r lands at position topn - 1 - r: ascending weight, ties by ascending index.
template <typename T>
__global__ void __launch_bounds__(256)
topk_columns_kernel(int64_t rows, int k, const T* __restrict__ F, int64_t ld, int topn, int32_t* __restrict__ out) {
    __shared__ T sv[256];
    __shared__ int64_t si[256];
    const int c = blockIdx.x;
    T pv = T(0);
    int64_t pi = -1;                      // previous pick; -1 = none yet
    bool exhausted = false;               // fewer than topn (non-NaN) rows: the remaining slots get -1
    for (int r = 0; r < topn; r++) {
        if (exhausted) {
            if (threadIdx.x == 0) out[int64_t(c) * topn + (topn - 1 - r)] = -1;
            continue;
        }
        T bv = T(0);
        int64_t bi = -1;
        for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
            const T v = F[i * ld + c];
            if (v != v) continue;                                             // NaN never ranks
            const bool below = pi < 0 || v < pv || (v == pv && i < pi);
            if (below && (bi < 0 || v > bv || (v == bv && i > bi))) { bv = v; bi = i; }
        }
        sv[threadIdx.x] = bv;
        si[threadIdx.x] = bi;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const T ov = sv[threadIdx.x + s];
                const int64_t oi = si[threadIdx.x + s];
                const int64_t mi = si[threadIdx.x];
                if (oi >= 0 && (mi < 0 || ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi > mi))) {
                    sv[threadIdx.x] = ov;
                    si[threadIdx.x] = oi;
                }
            }
            __syncthreads();
        }
        pv = sv[0];
        pi = si[0];
        if (threadIdx.x == 0) out[int64_t(c) * topn + (topn - 1 - r)] = int32_t(pi);    // -1 when fewer than topn rows
        __syncthreads();
        exhausted = pi < 0;
    }
}

template <typename T>
void topk_columns(pycmf_ctx* ctx, int64_t rows, int64_t k, const T* F, int64_t ld, int topn, int32_t* out) {
    if (k <= 0 || topn <= 0) return;
    PYCMF_CHECK(rows < (int64_t(1) << 31), "topk_columns: too many rows for int32 indices");
    PYCMF_CHECK(topn <= 1024, "topk_columns: topn too large");
    topk_columns_kernel<T><<<(unsigned)k, 256, 0, ctx->stream>>>(rows, int(k), F, ld, topn, out);
    PYCMF_LAUNCH_CHECK(ctx);
}

#define INSTANTIATE(T)                                                                                   \
    template void gemm<T>(pycmf_ctx*, bool, int64_t, int64_t, int64_t, const T*, int64_t, const T*,     \
                          int64_t, T*, int64_t, T, T);                                                   \
    template void reduce_parts<T>(pycmf_ctx*, int64_t, int64_t, int, const T*, T*, int64_t, T, T);        \
    template void broadcast_add<T>(pycmf_ctx*, int64_t, int64_t, T*, const T*, T, bool);                  \
    template void gram_f64<T>(pycmf_ctx*, int64_t, int64_t, const T*, double*);                           \
    template void mu_apply<T>(pycmf_ctx*, int64_t, int64_t, T*, const T*, const T*, double, double);     \
    template void transpose<T>(pycmf_ctx*, int64_t, int64_t, const T*, int64_t, T*, int64_t);           \
    template void axpby<T>(pycmf_ctx*, int64_t, T, const T*, T, const T*, T*);                          \
    template void dot_f64<T>(pycmf_ctx*, int64_t, const T*, const T*, double, double*, bool);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace pycmf

namespace pycmf {
template bool mu_fused_apply<float>(pycmf_ctx*, int64_t, int64_t, float*, const float*, const float*, double, double);
template bool mu_fused_apply<double>(pycmf_ctx*, int64_t, int64_t, double*, const double*, const double*, double, double);
template void topk_columns<float>(pycmf_ctx*, int64_t, int64_t, const float*, int64_t, int, int32_t*);
template void topk_columns<double>(pycmf_ctx*, int64_t, int64_t, const double*, int64_t, int, int32_t*);
}  // namespace pycmf

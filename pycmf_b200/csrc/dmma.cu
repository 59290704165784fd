// float64 tensor-core path (north_star: "an fp64 DMMA path"): C (m x q) = alpha * op(A) * B + beta * C on the
// mma.sync.m8n8k4.f64 pipe for the dense contractions of the float64 fit -- X V and X^T U of the MU step
// (cmf_solvers.py:232, :244), the factor-sized products (Y Z, Y^T V, F (B^T B), :233-245) and every np.dot the Newton
// phases route through gemm<double>.  q is a factor width (n_components or l), so the CTA tile is tall: 128 rows of C by
// 32 or 64 columns, K step 16, three cp.async stages.  Eight warps, each a 32 x (BN / 2) block of C = 4 x (BN / 16) DMMA
// tiles held in registers; operands come from shared memory with row pitches chosen so that a half-warp's 64-bit fragment
// loads hit 16 distinct double-wide banks (pitch = 4 mod 16 doubles).  Small C (X^T U on C2: 5000 x 32) is split along the
// contraction and reduced deterministically, like the FMA kernel it replaces.
#include "common.cuh"

namespace pycmf {
namespace {

constexpr int DBM = 128, DBK = 16, DSTAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

template <bool TRANS_A, int BN>
struct DmmaSmem {
    static constexpr int LDA = TRANS_A ? DBM + 4 : DBK + 4;             // doubles; = 4 mod 16
    static constexpr int A_ELEMS = TRANS_A ? DBK * LDA : DBM * LDA;
    static constexpr int LDB = BN + 4;
    static constexpr int B_ELEMS = DBK * LDB;
    static constexpr size_t BYTES = sizeof(double) * size_t(DSTAGES) * (A_ELEMS + B_ELEMS);
};

// C_part[z] (m x q) = op(A)[:, pz0:pz1] * B[pz0:pz1, :]; direct: C = alpha * (.) + beta * C
template <bool TRANS_A, int BN>
__global__ void __launch_bounds__(256)
dmma_gemm_kernel(int64_t m, int64_t q, int64_t p, int64_t p_per_split, const double* __restrict__ A, int64_t lda,
                 const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
                 int64_t c_split_stride, double alpha, double beta, bool direct) {
    using S = DmmaSmem<TRANS_A, BN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + DSTAGES * S::A_ELEMS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;                 // 4 x 2 warps
    constexpr int WNC = BN / 2;                              // columns per warp
    constexpr int MB = 4, NB = WNC / 8;
    const int64_t row0 = int64_t(blockIdx.y) * DBM, col0 = int64_t(blockIdx.x) * BN;
    const int64_t pz0 = int64_t(blockIdx.z) * p_per_split;
    const int64_t pz1 = min(p, pz0 + p_per_split);
    const int nk = int((pz1 - pz0 + DBK - 1) / DBK);

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        const int64_t k0 = pz0 + int64_t(kt) * DBK;
        double* as = As + stage * S::A_ELEMS;
        double* bs = Bs + stage * S::B_ELEMS;
        if (TRANS_A) {
            // op(A)[row][kk] = A[(k0 + kk) * lda + row0 + row]: rows of X, contiguous along the C-row index
#pragma unroll
            for (int it = 0; it < (DBK * DBM / 2) / 256; it++) {
                const int e = tid + it * 256;
                const int kk = e / (DBM / 2), c2 = e % (DBM / 2);
                const int64_t gk = k0 + kk, gr = row0 + 2 * c2;
                int bytes = 0;
                if (gk < pz1 && gr < m) bytes = gr + 1 < m ? 16 : 8;
                cp_async16(as + kk * S::LDA + 2 * c2, bytes ? A + gk * lda + gr : A, bytes);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (DBM * DBK / 2) / 256; it++) {
                const int e = tid + it * 256;
                const int r = e / (DBK / 2), c2 = e % (DBK / 2);
                const int64_t gr = row0 + r, gk = k0 + 2 * c2;
                int bytes = 0;
                if (gr < m && gk < pz1) bytes = gk + 1 < pz1 ? 16 : 8;
                cp_async16(as + r * S::LDA + 2 * c2, bytes ? A + gr * lda + gk : A, bytes);
            }
        }
#pragma unroll
        for (int it = 0; it < (DBK * BN / 2 + 255) / 256; it++) {
            const int e = tid + it * 256;
            if (e < DBK * BN / 2) {
                const int kk = e / (BN / 2), c2 = e % (BN / 2);
                const int64_t gk = k0 + kk, gc = col0 + 2 * c2;
                int bytes = 0;
                if (gk < pz1 && gc < q) bytes = gc + 1 < q ? 16 : 8;
                cp_async16(bs + kk * S::LDB + 2 * c2, bytes ? B + gk * ldb + gc : B, bytes);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < DSTAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    const int fr = lane >> 2, fk = lane & 3;                 // fragment row / column-in-k of this lane
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<DSTAGES - 2>();
        __syncthreads();
        if (kt + DSTAGES - 1 < nk) load_stage((kt + DSTAGES - 1) % DSTAGES, kt + DSTAGES - 1);
        cp_async_commit();
        const double* as = As + (kt % DSTAGES) * S::A_ELEMS;
        const double* bs = Bs + (kt % DSTAGES) * S::B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < DBK; kk += 4) {
            double a[MB], b[NB];
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const int r = wm * 32 + i * 8 + fr;
                a[i] = TRANS_A ? as[(kk + fk) * S::LDA + r] : as[r * S::LDA + kk + fk];
            }
#pragma unroll
            for (int j = 0; j < NB; j++) b[j] = bs[(kk + fk) * S::LDB + wn * WNC + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) dmma884(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    double* Cz = C + int64_t(blockIdx.z) * c_split_stride;
#pragma unroll
    for (int i = 0; i < MB; i++) {
        const int64_t gr = row0 + wm * 32 + i * 8 + fr;
        if (gr >= m) continue;
#pragma unroll
        for (int j = 0; j < NB; j++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t gc = col0 + wn * WNC + j * 8 + 2 * fk + h;
                if (gc >= q) continue;
                if (direct) {
                    const double prev = beta != 0.0 ? Cz[gr * ldc + gc] : 0.0;
                    Cz[gr * ldc + gc] = alpha * acc[i][j][h] + beta * prev;
                } else {
                    Cz[gr * ldc + gc] = acc[i][j][h];
                }
            }
        }
    }
}

template <bool TRANS_A, int BN>
void launch_dmma(pycmf_ctx* ctx, dim3 grid, int64_t m, int64_t q, int64_t p, int64_t p_per, const double* A, int64_t lda,
                 const double* B, int64_t ldb, double* C, int64_t ldc, int64_t stride, double alpha, double beta,
                 bool direct) {
    auto kern = dmma_gemm_kernel<TRANS_A, BN>;
    const size_t smem = DmmaSmem<TRANS_A, BN>::BYTES;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, 256, smem, ctx->stream>>>(m, q, p, p_per, A, lda, B, ldb, C, ldc, stride, alpha, beta, direct);
    PYCMF_LAUNCH_CHECK(ctx);
}


// ---- fused residual pass on the fp64 tensor cores ---------------------------------------------------------------------
//     R = f(Own Oth^T) - target tile  (never written to HBM),   out (own x k) += R Oth,   sq += sum R^2
// = the float64 form of reference cmf_solvers.py:399-400 (own = rows of U: out = R V) and :436-440 (own = rows of V,
// target read transposed: out = R^T U) and of the dense objective (:36-42).  A CTA owns 64 "own" rows (their factor rows
// stay in shared memory for the whole kernel) and walks the other factor in 32-row tiles (cp.async, double-buffered):
// GEMM 1 (64 x 32 x k DMMAs) -> link, minus target (loaded into the accumulator layout before the tile arrives) -> R tile
// in shared memory -> GEMM 2 (64 x k x 32 DMMAs) into register accumulators.  k <= 128.
constexpr int ROWN = 64, ROT = 32;

template <int NB2>          // n-blocks of GEMM 2 per warp: the padded factor width is 16 * NB2
__global__ void __launch_bounds__(256)
dmma_resid_kernel(int64_t own_n, int64_t oth_n, int k, const double* __restrict__ Own, const double* __restrict__ Oth,
                  const double* __restrict__ Tgt, int64_t ldt, bool t_own_major, int link,
                  double* __restrict__ out, int64_t out_split_stride, int64_t tiles_per_split,
                  double* __restrict__ sq_part, bool want_out) {
    constexpr int KPAD = 16 * NB2;
    constexpr int LD = KPAD + 4;                   // = 4 mod 16: conflict-free 64-bit fragment reads
    constexpr int LDR = ROT + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);              // 32
    double* own_s = red + 32;                                       // ROWN x LD
    double* oth_s = own_s + ROWN * LD;                              // 2 x ROT x LD
    double* R_s = oth_s + 2 * ROT * LD;                             // ROWN x LDR
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int fr = lane >> 2, fk = lane & 3;
    const int64_t own0 = int64_t(blockIdx.x) * ROWN;
    const int64_t t_begin = int64_t(blockIdx.y) * tiles_per_split;
    const int64_t t_end = min(t_begin + tiles_per_split, (oth_n + ROT - 1) / ROT);

    // factor rows -> shared memory, zero-padded to KPAD columns / missing rows (plain loads: k need not be even)
    for (int e = tid; e < ROWN * KPAD; e += 256) {
        const int r = e / KPAD, c = e % KPAD;
        own_s[r * LD + c] = (own0 + r < own_n && c < k) ? Own[(own0 + r) * k + c] : 0.0;
    }
    auto load_oth = [&](int buf, int64_t t) {
        double* dst = oth_s + buf * ROT * LD;
        const int64_t o0 = t * ROT;
        if ((k & 1) == 0) {
            for (int e = tid; e < ROT * (KPAD / 2); e += 256) {
                const int r = e / (KPAD / 2), c = 2 * (e % (KPAD / 2));
                const bool ok = o0 + r < oth_n && c < k;
                cp_async16(dst + r * LD + c, ok ? Oth + (o0 + r) * k + c : Oth, ok ? 16 : 0);
            }
        } else {
            for (int e = tid; e < ROT * KPAD; e += 256) {
                const int r = e / KPAD, c = e % KPAD;
                dst[r * LD + c] = (o0 + r < oth_n && c < k) ? Oth[(o0 + r) * k + c] : 0.0;
            }
        }
    };

    double acc2[2][NB2][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < NB2; j++) acc2[i][j][0] = acc2[i][j][1] = 0.0;
    double sq_local = 0.0;

    if (t_begin < t_end) load_oth(0, t_begin);
    cp_async_commit();
    for (int64_t t = t_begin; t < t_end; t++) {
        const int buf = int((t - t_begin) & 1);
        const int64_t oth0 = t * ROT;
        // target values of this thread's S elements, requested before the tile is waited for
        double tg[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t ro = own0 + wm * 16 + i * 8 + fr, co = oth0 + wn * 16 + j * 8 + 2 * fk + h;
                    double v = 0.0;
                    if (Tgt != nullptr && ro < own_n && co < oth_n) v = t_own_major ? __ldg(Tgt + ro * ldt + co) : __ldg(Tgt + co * ldt + ro);
                    tg[i][j][h] = v;
                }
        if (t + 1 < t_end) load_oth(buf ^ 1, t + 1);      // the other buffer was released by the barrier that ended tile t - 1
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* os = oth_s + buf * ROT * LD;
        // ---- GEMM 1: S (64 x 32) = Own (64 x KPAD) Oth^T
        double s1[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) s1[i][j][0] = s1[i][j][1] = 0.0;
#pragma unroll 4
        for (int kk = 0; kk < KPAD; kk += 4) {
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; i++) a[i] = own_s[(wm * 16 + i * 8 + fr) * LD + kk + fk];
#pragma unroll
            for (int j = 0; j < 2; j++) b[j] = os[(wn * 16 + j * 8 + fr) * LD + kk + fk];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma884(s1[i][j], a[i], b[j]);
        }
        // ---- link, minus target, R tile to shared memory
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int rl = wm * 16 + i * 8 + fr, cl = wn * 16 + j * 8 + 2 * fk + h;
                    double r = 0.0;
                    if (own0 + rl < own_n && oth0 + cl < oth_n) {
                        const double est = link == PYCMF_LOGIT ? sigmoid_<double>(s1[i][j][h]) : s1[i][j][h];
                        r = est - tg[i][j][h];
                        sq_local = fma(r, r, sq_local);
                    }
                    R_s[rl * LDR + cl] = r;
                }
        __syncthreads();
        // ---- GEMM 2: out (64 x KPAD) += R (64 x 32) Oth (32 x KPAD)
        if (want_out) {
#pragma unroll
            for (int kk = 0; kk < ROT; kk += 4) {
                double a[2], b[NB2];
#pragma unroll
                for (int i = 0; i < 2; i++) a[i] = R_s[(wm * 16 + i * 8 + fr) * LDR + kk + fk];
#pragma unroll
                for (int j = 0; j < NB2; j++) b[j] = os[(kk + fk) * LD + wn * (KPAD / 2) + j * 8 + fr];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < NB2; j++) dmma884(acc2[i][j], a[i], b[j]);
            }
        }
        __syncthreads();          // R_s and this tile's buffer are free again
    }
    cp_async_wait<0>();
    if (want_out) {
        double* o = out + int64_t(blockIdx.y) * out_split_stride;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int64_t gr = own0 + wm * 16 + i * 8 + fr;
            if (gr >= own_n) continue;
#pragma unroll
            for (int j = 0; j < NB2; j++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int c = wn * (KPAD / 2) + j * 8 + 2 * fk + h;
                    if (c < k) o[gr * k + c] = acc2[i][j][h];
                }
        }
    }
    if (sq_part != nullptr) {
        const double v = block_sum(sq_local, red);
        if (tid == 0) sq_part[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
}

template <int NB2>
void launch_dmma_resid(pycmf_ctx* ctx, const char* family, int64_t own_n, int64_t oth_n, int64_t k, const double* Own,
                       const double* Oth, const double* Tgt, int64_t ldt, bool t_own_major, int link, double* out,
                       double* sq) {
    constexpr int KPAD = 16 * NB2, LD = KPAD + 4;
    const int64_t own_blocks = ceil_div(own_n, ROWN), loop_tiles = ceil_div(oth_n, ROT);
    int64_t splits = 1;
    // two resident CTAs per SM (registers): aim at ~4 waves of CTAs so that the last wave's tail stays small
    if (own_blocks < 8 * ctx->num_sms)
        splits = std::max<int64_t>(1, std::min(ceil_div(loop_tiles, 8), ceil_div(int64_t(8) * ctx->num_sms, own_blocks)));
    int64_t tiles_per_split = ceil_div(loop_tiles, splits);
    splits = ceil_div(loop_tiles, tiles_per_split);
    PYCMF_CHECK(splits <= 65535, "resid_pass: too many splits");
    const size_t smem = sizeof(double) * (32 + size_t(ROWN) * LD + size_t(2) * ROT * LD + size_t(ROWN) * (ROT + 4));
    auto kern = dmma_resid_kernel<NB2>;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    double* target = out;
    int64_t stride = 0;
    if (out != nullptr && splits > 1) {
        target = static_cast<double*>(scratch(ctx, 0, size_t(splits) * own_n * k * sizeof(double)));
        stride = own_n * k;
    }
    double* sq_part = nullptr;
    const int64_t nparts = own_blocks * splits;
    if (sq != nullptr) sq_part = static_cast<double*>(scratch(ctx, 1, size_t(nparts) * sizeof(double)));
    dim3 grid((unsigned)own_blocks, (unsigned)splits);
    Timed timer(ctx, family);
    kern<<<grid, 256, smem, ctx->stream>>>(own_n, oth_n, int(k), Own, Oth, Tgt, ldt, t_own_major, link, target, stride,
                                           tiles_per_split, sq_part, out != nullptr);
    PYCMF_LAUNCH_CHECK(ctx);
    if (out != nullptr && splits > 1) reduce_parts<double>(ctx, own_n, k, int(splits), target, out, k, 1.0, 0.0);
    if (sq != nullptr) final_sum(ctx, int(nparts), sq_part, 1.0, sq, true);
}

}  // namespace

bool dmma_gemm_eligible(pycmf_ctx* ctx, int64_t m, int64_t q, int64_t p, const double* A, int64_t lda, const double* B,
                        int64_t ldb) {
    if (ctx->dense_path == 0) return false;                                   // option: FMA kernels only (tests)
    if (m < 64 || q < 8 || p < 64) return false;                              // launch-bound sizes: the FMA kernel is as good
    if ((lda & 1) || (ldb & 1)) return false;                                 // 16-byte cp.async chunks
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
    return true;
}

void dmma_gemm(pycmf_ctx* ctx, bool trans_a, int64_t m, int64_t q, int64_t p, const double* A, int64_t lda,
               const double* B, int64_t ldb, double* C, int64_t ldc, double alpha, double beta) {
    const int bn = q <= 32 ? 32 : 64;
    const int64_t tiles = ceil_div(m, DBM) * ceil_div(q, bn);
    int splits = 1;
    if (tiles < 2 * ctx->num_sms) {
        const int64_t want = ceil_div(int64_t(3) * ctx->num_sms, tiles);
        const int64_t maxs = ceil_div(p, 512);
        splits = int(std::max<int64_t>(1, std::min(want, maxs)));
    }
    int64_t p_per = ceil_div(p, splits);
    p_per = ceil_div(p_per, DBK) * DBK;
    splits = int(std::max<int64_t>(1, ceil_div(p, p_per)));
    PYCMF_CHECK(ceil_div(m, DBM) <= 65535, "dmma_gemm: too many row tiles");
    dim3 grid((unsigned)ceil_div(q, bn), (unsigned)ceil_div(m, DBM), (unsigned)splits);
    Timed timer(ctx, "dmma_gemm");
    double* out = C;
    int64_t ldo = ldc, stride = 0;
    bool direct = true;
    if (splits > 1) {
        out = static_cast<double*>(scratch(ctx, 0, size_t(splits) * m * q * sizeof(double)));
        ldo = q;
        stride = m * q;
        direct = false;
    }
#define GO(TR, BNV) launch_dmma<TR, BNV>(ctx, grid, m, q, p, p_per, A, lda, B, ldb, out, ldo, stride, alpha, beta, direct)
    if (trans_a) { if (bn == 32) GO(true, 32); else GO(true, 64); }
    else { if (bn == 32) GO(false, 32); else GO(false, 64); }
#undef GO
    if (splits > 1) reduce_parts<double>(ctx, m, q, splits, out, C, ldc, alpha, beta);
}

bool dmma_resid_eligible(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k) {
    return ctx->dense_path != 0 && k <= 128 && ra >= 64 && rb >= 64;
}

// mode 0: own = rows of A, out = R B (ra x k); mode 1: own = rows of B, out = R^T A (rb x k).  R = f(A B^T) - Tgt.
void dmma_resid(pycmf_ctx* ctx, int mode, int64_t ra, int64_t rb, int64_t k, const double* A, const double* B,
                const double* Tgt, int64_t ldt, bool trans_t, int link, double* out, double* sq) {
    const double* own = mode == 0 ? A : B;
    const double* oth = mode == 0 ? B : A;
    const int64_t own_n = mode == 0 ? ra : rb, oth_n = mode == 0 ? rb : ra;
    // target element of (own i, other j): Tgt[a, b] stored row-major (or transposed when trans_t)
    const bool t_own_major = (mode == 0) != trans_t;
    const char* fam = mode == 0 ? "dmma_resid_left" : "dmma_resid_right";
    if (k <= 16) launch_dmma_resid<1>(ctx, fam, own_n, oth_n, k, own, oth, Tgt, ldt, t_own_major, link, out, sq);
    else if (k <= 32) launch_dmma_resid<2>(ctx, fam, own_n, oth_n, k, own, oth, Tgt, ldt, t_own_major, link, out, sq);
    else if (k <= 64) launch_dmma_resid<4>(ctx, fam, own_n, oth_n, k, own, oth, Tgt, ldt, t_own_major, link, out, sq);
    else launch_dmma_resid<8>(ctx, fam, own_n, oth_n, k, own, oth, Tgt, ldt, t_own_major, link, out, sq);
}

}  // namespace pycmf

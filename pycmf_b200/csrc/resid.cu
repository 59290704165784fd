// Fused residual pass (generic FMA path, fp32 / fp64):
//     R = f(A B^T) - Tgt          (never written to HBM)
//     outL = R B,  outR = R^T A,  sq += sum R^2
// One CTA owns a TILE of "own" rows (rows of A in LEFT mode, rows of B in RIGHT mode) and walks
// the other operand tile by tile: S tile -> link -> minus target -> R tile in shared memory ->
// second product into register accumulators.  This is the reference's
//     res = inverse(np.dot(U, V.T), link) - X ;  np.dot(res, V) / np.dot(res.T, U)
// (cmf_solvers.py:399-400, :436-440) and the dense objective (:36-42) without the n x d temporary.
#include <type_traits>

#include "common.cuh"

namespace pycmf {
namespace {

template <typename T> struct Tile { static constexpr int value = 64; };
template <> struct Tile<double> { static constexpr int value = 32; };

// MODE 0 = LEFT (own = A rows, out = R B), MODE 1 = RIGHT (own = B rows, out = R^T A)
template <typename T, int KC, int MODE>
__global__ void __launch_bounds__(256)
resid_kernel(int64_t ra, int64_t rb, int k, const T* __restrict__ A, const T* __restrict__ B,
             const T* __restrict__ Tgt, int64_t ldt, bool trans_t, int link,
             T* __restrict__ out, int64_t out_split_stride, int64_t tiles_per_split,
             double* __restrict__ sq_part, bool want_out) {
    constexpr int TILE = Tile<T>::value;
    constexpr int TM = TILE / 16;
    constexpr int KCOLS = KC * 16;
    constexpr int KP = KCOLS + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);
    T* A_s = reinterpret_cast<T*>(smem_raw + 32 * sizeof(double));
    T* B_s = A_s + TILE * KP;
    T* R_s = B_s + TILE * KP;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t own0 = int64_t(blockIdx.x) * TILE;
    const int64_t other_n = MODE == 0 ? rb : ra;
    const int64_t t_begin = int64_t(blockIdx.y) * tiles_per_split;
    const int64_t t_end = min(t_begin + tiles_per_split, (other_n + TILE - 1) / TILE);

    const T* own_ptr = MODE == 0 ? A : B;
    const T* oth_ptr = MODE == 0 ? B : A;
    T* own_s = MODE == 0 ? A_s : B_s;
    T* oth_s = MODE == 0 ? B_s : A_s;
    const int64_t own_n = MODE == 0 ? ra : rb;

    for (int e = tid; e < TILE * KCOLS; e += 256) {
        int r = e / KCOLS, c = e % KCOLS;
        int64_t gr = own0 + r;
        own_s[r * KP + c] = (gr < own_n && c < k) ? own_ptr[gr * k + c] : T(0);
    }

    T acc[TM][KC];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int m = 0; m < KC; m++) acc[i][m] = T(0);
    double sq_local = 0.0;

    for (int64_t t = t_begin; t < t_end; t++) {
        const int64_t oth0 = t * TILE;
        __syncthreads();  // previous iteration finished reading oth_s / R_s
        for (int e = tid; e < TILE * KCOLS; e += 256) {
            int r = e / KCOLS, c = e % KCOLS;
            int64_t gr = oth0 + r;
            oth_s[r * KP + c] = (gr < other_n && c < k) ? oth_ptr[gr * k + c] : T(0);
        }
        __syncthreads();
        // ---- S micro tile: rows of A (ty + 16 i), rows of B (tx + 16 j)
        T s[TM][TM];
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < TM; j++) s[i][j] = T(0);
        for (int kk = 0; kk < k; kk++) {
            T a[TM], b[TM];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = A_s[(ty + 16 * i) * KP + kk];
#pragma unroll
            for (int j = 0; j < TM; j++) b[j] = B_s[(tx + 16 * j) * KP + kk];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TM; j++) s[i][j] = fma(a[i], b[j], s[i][j]);
        }
        const int64_t a0 = MODE == 0 ? own0 : oth0;
        const int64_t b0 = MODE == 0 ? oth0 : own0;
#pragma unroll
        for (int i = 0; i < TM; i++) {
#pragma unroll
            for (int j = 0; j < TM; j++) {
                int64_t row = a0 + ty + 16 * i, col = b0 + tx + 16 * j;
                T r = T(0);
                if (row < ra && col < rb) {
                    T est = link == PYCMF_LOGIT ? sigmoid_<T>(s[i][j]) : s[i][j];
                    T tg = T(0);
                    if (Tgt != nullptr) tg = trans_t ? Tgt[col * ldt + row] : Tgt[row * ldt + col];
                    r = est - tg;
                    sq_local += double(r) * double(r);
                }
                R_s[(ty + 16 * i) * (TILE + 1) + tx + 16 * j] = r;
            }
        }
        __syncthreads();
        if (want_out) {
            if (MODE == 0) {
                // acc[i][m] += sum_j R[ty+16i][j] * B_s[j][tx+16m]
                for (int j = 0; j < TILE; j++) {
                    T rv[TM];
#pragma unroll
                    for (int i = 0; i < TM; i++) rv[i] = R_s[(ty + 16 * i) * (TILE + 1) + j];
#pragma unroll
                    for (int m = 0; m < KC; m++) {
                        T b = B_s[j * KP + tx + 16 * m];
#pragma unroll
                        for (int i = 0; i < TM; i++) acc[i][m] = fma(rv[i], b, acc[i][m]);
                    }
                }
            } else {
                // acc[i][m] += sum_r R[r][ty+16i] * A_s[r][tx+16m]
                for (int r = 0; r < TILE; r++) {
                    T rv[TM];
#pragma unroll
                    for (int i = 0; i < TM; i++) rv[i] = R_s[r * (TILE + 1) + ty + 16 * i];
#pragma unroll
                    for (int m = 0; m < KC; m++) {
                        T a = A_s[r * KP + tx + 16 * m];
#pragma unroll
                        for (int i = 0; i < TM; i++) acc[i][m] = fma(rv[i], a, acc[i][m]);
                    }
                }
            }
        }
    }
    if (want_out) {
        T* o = out + int64_t(blockIdx.y) * out_split_stride;
#pragma unroll
        for (int i = 0; i < TM; i++) {
            int64_t gr = own0 + ty + 16 * i;
            if (gr >= own_n) continue;
#pragma unroll
            for (int m = 0; m < KC; m++) {
                int c = tx + 16 * m;
                if (c < k) o[gr * k + c] = acc[i][m];
            }
        }
    }
    if (sq_part != nullptr) {
        double v = block_sum(sq_local, red);
        if (tid == 0) sq_part[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
}

template <typename T, int KC, int MODE>
void launch_resid(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const T* A, const T* B, const T* Tgt,
                  int64_t ldt, bool trans_t, int link, T* out, double* sq) {
    constexpr int TILE = Tile<T>::value;
    constexpr int KP = KC * 16 + 1;
    const int64_t own_n = MODE == 0 ? ra : rb, other_n = MODE == 0 ? rb : ra;
    const int64_t own_blocks = ceil_div(own_n, TILE), loop_tiles = ceil_div(other_n, TILE);
    int64_t splits = 1;
    if (own_blocks < 2 * ctx->num_sms)
        splits = std::max<int64_t>(1, std::min(loop_tiles, ceil_div(int64_t(4) * ctx->num_sms, own_blocks)));
    int64_t tiles_per_split = ceil_div(loop_tiles, splits);
    splits = ceil_div(loop_tiles, tiles_per_split);
    PYCMF_CHECK(splits <= 65535, "resid_pass: too many splits");
    size_t smem = 32 * sizeof(double) + (size_t(2) * TILE * KP + size_t(TILE) * (TILE + 1)) * sizeof(T);
    auto kern = resid_kernel<T, KC, MODE>;
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    T* target = out;
    int64_t stride = 0;
    if (out != nullptr && splits > 1) {
        target = static_cast<T*>(scratch(ctx, 0, size_t(splits) * own_n * k * sizeof(T)));
        stride = own_n * k;
    }
    double* sq_part = nullptr;
    int64_t nparts = own_blocks * splits;
    if (sq != nullptr) sq_part = static_cast<double*>(scratch(ctx, 1, size_t(nparts) * sizeof(double)));
    dim3 grid((unsigned)own_blocks, (unsigned)splits);
    Timed timer(ctx, MODE == 0 ? "resid_left" : "resid_right");
    kern<<<grid, 256, smem, ctx->stream>>>(ra, rb, int(k), A, B, Tgt, ldt, trans_t, link, target, stride,
                                           tiles_per_split, sq_part, out != nullptr);
    PYCMF_LAUNCH_CHECK(ctx);
    if (out != nullptr && splits > 1) reduce_parts<T>(ctx, own_n, k, int(splits), target, out, k, T(1), T(0));
    if (sq != nullptr) final_sum(ctx, int(nparts), sq_part, 1.0, sq, true);
}

template <typename T, int MODE>
void dispatch_kc(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const T* A, const T* B, const T* Tgt,
                 int64_t ldt, bool trans_t, int link, T* out, double* sq) {
    if (k <= 16) launch_resid<T, 1, MODE>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, out, sq);
    else if (k <= 32) launch_resid<T, 2, MODE>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, out, sq);
    else if (k <= 64) launch_resid<T, 4, MODE>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, out, sq);
    else if (k <= 128) launch_resid<T, 8, MODE>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, out, sq);
    else launch_resid<T, 16, MODE>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, out, sq);
}

}  // namespace

template <typename T>
void resid_pass(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const T* A, const T* B,
                const T* Tgt, int64_t ldt, bool trans_t, int link, T* outL, T* outR, double* sq) {
    PYCMF_CHECK(k >= 1 && k <= 256, "n_components must be in [1, 256] for the fused residual pass");
    if (ra <= 0 || rb <= 0) {
        if (outL && ra > 0) PYCMF_CUDA(cudaMemsetAsync(outL, 0, size_t(ra) * k * sizeof(T), ctx->stream));
        if (outR && rb > 0) PYCMF_CUDA(cudaMemsetAsync(outR, 0, size_t(rb) * k * sizeof(T), ctx->stream));
        return;
    }
    if constexpr (std::is_same<T, float>::value) {
        if ((outL != nullptr || outR != nullptr) && tc_dense_eligible(ctx, ra, rb, k, Tgt, ldt, trans_t)) {
            tc_resid_pass(ctx, ra, rb, A, B, Tgt, ldt, link, outL, outR, sq);
            return;
        }
    }
    if constexpr (std::is_same<T, double>::value) {
        if (dmma_resid_eligible(ctx, ra, rb, k)) {               // float64 on the DMMA pipe (dmma.cu)
            if (outL != nullptr || (outR == nullptr && sq != nullptr))
                dmma_resid(ctx, 0, ra, rb, k, A, B, Tgt, ldt, trans_t, link, outL, sq);
            if (outR != nullptr) dmma_resid(ctx, 1, ra, rb, k, A, B, Tgt, ldt, trans_t, link, outR, outL == nullptr ? sq : nullptr);
            return;
        }
    }
    if (outL != nullptr || (outR == nullptr && sq != nullptr))
        dispatch_kc<T, 0>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, outL, sq);
    if (outR != nullptr)
        dispatch_kc<T, 1>(ctx, ra, rb, k, A, B, Tgt, ldt, trans_t, link, outR,
                          (outL == nullptr) ? sq : nullptr);
}

template void resid_pass<float>(pycmf_ctx*, int64_t, int64_t, int64_t, const float*, const float*, const float*,
                                int64_t, bool, int, float*, float*, double*);
template void resid_pass<double>(pycmf_ctx*, int64_t, int64_t, int64_t, const double*, const double*,
                                 const double*, int64_t, bool, int, double*, double*, double*);

}  // namespace pycmf

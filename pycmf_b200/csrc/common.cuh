// Shared declarations for libpycmf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <stdexcept>
#include <utility>
#include <vector>

#include "../../include/pycmf_b200.h"

namespace pycmf {

constexpr double kEpsF32 = 1.1920928955078125e-07;  // np.finfo(np.float32).eps, cmf_solvers.py:12

struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace pycmf

struct pycmf_ctx {
    int device = 0;
    int num_sms = 148;
    int max_smem_optin = 0;
    cudaStream_t stream = 0;
    int64_t launches = 0;
    // options
    int chol_fastpath = 1;
    int dense_path = 1;
    int tc_trace = 0;        // diagnostics: record a pipeline trace of CTA (0,0) of every tcgen05 pass into arena 2
    int tc_max_splits = 0;   // > 0: at most this many CTAs per own tile in the tcgen05 passes (tests: 1 = one long chain)
    int tc_ctas = 0;         // > 0: cap on the persistent CTA count of the tcgen05 passes (tests)
    int tc_prefetch = 0;     // L2 prefetch of X in the tcgen05 passes: 0 none (measured: no gain), 2 TMA prefetch
    int tc_x_promotion = 128; // L2 promotion (bytes) of the X tensor map of the tcgen05 passes: 0, 64, 128, 256
    int tc_chain = 0;        // > 0: accumulation chain cap in tiles (default 16)
    size_t max_scratch = size_t(2) << 30;
    // scratch arenas (grown on demand; growth synchronises the stream)
    pycmf::Scratch arena[10];
    std::vector<void*> retired;   // outgrown arenas: kept until destroy (a CUDA graph captured earlier may still point at them)
    // optional per-kernel-family timers (cudaEvent pairs recorded on the stream around each launch)
    int profile = 0;
    std::map<std::string, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>> timers;
    // Side context: a child with its own stream and arenas, used to run an independent branch of one phase next to the
    // main branch (pycmf::fork_side / join_side).  Its launches and timers are booked on the root.
    pycmf_ctx* root = nullptr;
    pycmf_ctx* side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int finish_minblocks = 2;   // option: resident CTAs per SM the V-finish kernel is compiled for (2: 255 registers, no
                                // spills: 86 us on C2; 3: 168 registers: 91 us; 4: 128 registers, spills: 187 us)
    int spmm_path = 1;       // option: 0 = generic SpMM kernel only (tests), 1 = vector kernels for k = 32 / 64 / 128 / 256, 2 = without the sub-warp grouping
    int spmm_blocks_per_sm = 0;  // option: resident 256-thread CTAs per SM of the nonzero-balanced SpMM (0 = default 4)
    int solve_path = 2;      // option: clamped solve with active clamp, k > 32: 1 = tridiagonalisation + bisection + inverse iteration
                             // (tridiag_solve.cuh), 2 = the same with the Householder steps on a register-resident matrix
                             // for k = 64 / 128, 0 = one-sided Jacobi only
    int solve_threads = 0;   // option (tuning): CTA size of the tridiagonal clamped solve (0 = by k; must be >= k, multiple of 32)
    int hess_mma = 1;        // option: 0 = per-row Hessian builds on the FMA pipes only (tests), 1 = mma.sync 3xTF32 for k = 64 / 128
    int mu_fused = 1;        // option: 0 = separate F G GEMM + elementwise ratio launches (tests)
    int spmm_lean = 1;       // option: 1 = shared-memory staged nonzeros + packed FMAs (spmm_nzb2_kernel), 0 = shuffle variant
    int spmm_unroll = 4;     // option: independent factor-row gathers per lane in flight (4 or 8)
    int side_streams = 1;    // option: 0 runs the side branches on the main stream (serial)
};

namespace pycmf {

void set_error(const std::string& msg);

struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define PYCMF_CHECK(cond, msg)                                                        \
    do {                                                                              \
        if (!(cond)) throw pycmf::Error(std::string(msg) + " [" #cond "]");           \
    } while (0)

#define PYCMF_CUDA(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess)                                                        \
            throw pycmf::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + \
                               " at " __FILE__ ":" + std::to_string(__LINE__));       \
    } while (0)

#define PYCMF_LAUNCH_CHECK(ctx)                                                       \
    do {                                                                              \
        ((ctx)->root ? (ctx)->root : (ctx))->launches++;                              \
        PYCMF_CUDA(cudaGetLastError());                                               \
    } while (0)

// RAII timer: when ctx->profile is on, brackets the enclosed launches with events on ctx->stream.
struct Timed {
    pycmf_ctx* ctx;
    cudaEvent_t stop = nullptr;
    Timed(pycmf_ctx* c, const char* name) : ctx(c) {
        pycmf_ctx* r = c->root ? c->root : c;
        if (!r->profile) return;
        cudaEvent_t start;
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
        cudaEventRecord(start, c->stream);
        r->timers[name].emplace_back(start, stop);
    }
    ~Timed() {
        if (stop) cudaEventRecord(stop, ctx->stream);
    }
};

// scratch arena `slot`, at least `bytes` large (256-B aligned by cudaMalloc)
void* scratch(pycmf_ctx* ctx, int slot, size_t bytes);

// Returns the side context after making its stream wait for everything enqueued so far on ctx->stream; work enqueued
// through the returned context runs concurrently with what the caller enqueues on ctx next.  join_side makes
// ctx->stream wait for the side branch.  Both are plain event record / wait pairs, so they also work while the
// stream is being captured into a CUDA graph (the side stream joins the capture and becomes a parallel branch).
pycmf_ctx* fork_side(pycmf_ctx* ctx);
void join_side(pycmf_ctx* ctx);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T> struct DType;
template <> struct DType<float> { static constexpr int code = PYCMF_F32; };
template <> struct DType<double> { static constexpr int code = PYCMF_F64; };

// ---- device helpers -----------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T sigmoid_(T x);
template <> __device__ __forceinline__ float sigmoid_<float>(float x) { return 1.0f / (1.0f + expf(-x)); }
template <> __device__ __forceinline__ double sigmoid_<double>(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of doubles; result valid in thread 0. `red` is >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ---- internal launchers (implemented across the .cu files; all enqueue on ctx->stream) -----
// dense.cu
template <typename T>
void gemm(pycmf_ctx* ctx, bool trans_a, int64_t m, int64_t q, int64_t p, const T* A, int64_t lda,
          const T* B, int64_t ldb, T* C, int64_t ldc, T alpha, T beta);
// C (m x q, ldc) = alpha * sum_z part[z] (each m x q contiguous) + beta * C
template <typename T>
void reduce_parts(pycmf_ctx* ctx, int64_t m, int64_t q, int splits, const T* part, T* C, int64_t ldc, T alpha, T beta);
// F *= N / (D + l1 + l2 F) with zero guard (cmf_solvers.py:212-228); all (rows x k) contiguous, ld given
template <typename T>
void mu_apply(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* N, const T* D, double l1, double l2);
// the same update with the denominator D = F G (G: k x k) formed inside the kernel (k <= 128); false = not eligible
template <typename T>
bool mu_fused_apply(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* N, const T* G, double l1, double l2);
template <typename T>
void transpose(pycmf_ctx* ctx, int64_t rows, int64_t cols, const T* A, int64_t lda, T* At, int64_t ldat);
// out[i] = alpha*a[i] + beta*b[i]
template <typename T>
void axpby(pycmf_ctx* ctx, int64_t n, T alpha, const T* a, T beta, const T* b, T* out);
// *out = sum a[i]*b[i] (double accumulation), out is a device double; accumulate adds to *out
template <typename T>
void dot_f64(pycmf_ctx* ctx, int64_t n, const T* a, const T* b, double scale, double* out, bool accumulate);

template <typename T>
void broadcast_add(pycmf_ctx* ctx, int64_t rows, int64_t kk, T* H, const T* Hs, T scale, bool overwrite);
// dmma.cu : the same product for float64 operands on the fp64 tensor-core pipe (mma.sync.m8n8k4.f64); gemm<double> routes
// to it when eligible (16-byte aligned operands, even pitches, not a launch-bound size, dense_path != 0)
bool dmma_gemm_eligible(pycmf_ctx* ctx, int64_t m, int64_t q, int64_t p, const double* A, int64_t lda, const double* B,
                        int64_t ldb);
void dmma_gemm(pycmf_ctx* ctx, bool trans_a, int64_t m, int64_t q, int64_t p, const double* A, int64_t lda,
               const double* B, int64_t ldb, double* C, int64_t ldc, double alpha, double beta);
// ... and the fused residual pass (R = f(A B^T) - Tgt in shared memory, out = R B or R^T A, sum R^2) for k <= 128
bool dmma_resid_eligible(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k);
void dmma_resid(pycmf_ctx* ctx, int mode, int64_t ra, int64_t rb, int64_t k, const double* A, const double* B,
                const double* Tgt, int64_t ldt, bool trans_t, int link, double* out, double* sq);
// out (k x topn int32): for every column c of F (rows x k, ld) the row indices of its topn largest entries in ASCENDING
// weight order (ties by ascending index) == np.argsort(F[:, c], kind="stable")[-topn:]   (reference analysis.py:6)
template <typename T>
void topk_columns(pycmf_ctx* ctx, int64_t rows, int64_t k, const T* F, int64_t ld, int topn, int32_t* out);
// G (k x k float64) = A^T A accumulated in float64
template <typename T>
void gram_f64(pycmf_ctx* ctx, int64_t rows, int64_t k, const T* A, double* G);

// *out = (accumulate ? *out : 0) + scale * sum(part[0..nparts))
void final_sum(pycmf_ctx* ctx, int nparts, const double* part, double scale, double* out, bool accumulate);

// sparse.cu
template <typename T>
void spmm(pycmf_ctx* ctx, int64_t rows, const int32_t* rowptr, const int32_t* colidx, const T* vals,
          const T* B, int64_t ldb, int64_t k, T* C, int64_t ldc, T alpha, T beta, int64_t b_rows = 0);
// mode 0: out += scale * sum_nz t_ij (a_i . b_j)
// mode 1: out += scale * sum_nz [ (t_ij - s_ij)^2 - s_ij^2 ],  s_ij = sigmoid(a_i . b_j)
// mode 2: out += scale * sum_nz t_ij^2
template <typename T>
void sddmm_reduce(pycmf_ctx* ctx, int mode, int64_t rows, const int32_t* rowptr, const int32_t* colidx,
                  const T* vals, const T* A, const T* B, int64_t k, double scale, double* out);

// resid.cu : R = f(A B^T) - Tgt (Tgt may be null);  outL = R B (ra x k), outR = R^T A (rb x k),
// *sq += sum R^2.  Any of outL / outR / sq may be null.  Tgt is (ra x rb, ldt) or, if trans_t,
// stored (rb x ra, ldt).
template <typename T>
void resid_pass(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const T* A, const T* B,
                const T* Tgt, int64_t ldt, bool trans_t, int link, T* outL, T* outR, double* sq);

// tc_resid.cu : tcgen05 / TMEM / TMA versions of the dense passes over X (fp32, n_components == 32)
int tc_trace_words();
bool tc_dense_eligible(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const float* X, int64_t ldx, bool trans_t);
void tc_resid_pass(pycmf_ctx* ctx, int64_t ra, int64_t rb, const float* A, const float* B, const float* X, int64_t ldx,
                   int link, float* outL, float* outR, double* sq);
void tc_xmul(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, const float* X, int64_t ldx, const float* Q,
             float* out);

// tc_mu.cu : tcgen05 MU numerators X Q / X^T Q for n_components in {64, 128, 192, 256} (fp32, 3xTF32)
bool tc_mu_eligible(pycmf_ctx* ctx, int64_t rows, int64_t cols, int64_t k, const float* X, int64_t ldx, bool trans_t);
// `family`: timer family the launch is booked under (default tc_xv / tc_xtu: the passes over X)
void tc_mu_xmul(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, int64_t k, const float* X, int64_t ldx,
                const float* Q, float* out, const char* family = nullptr);

// newton.cu
// Per-row gradient / Hessian accumulation for rows of A against (sampled) rows of B.
//   g_i (+)= w * sum_{j in s_i} (f(a_i.b_j) - t_ij) b_j ;  H_i (+)= w * sum_{j in s_i} f'(a_i.b_j) b_j b_j^T
// target: dense T (element (i,j) at T[i*ldt + j], or T[j*ldt + i] if trans_t), CSR row i lookup, or none (t = 0).
template <typename T>
void row_grad_hess(pycmf_ctx* ctx, int64_t rows, int64_t m, int64_t k, const T* A, const T* B,
                   const T* Tgt, int64_t ldt, bool trans_t,
                   const int32_t* rowptr, const int32_t* colidx, const T* vals,
                   int link, double w, const int32_t* idx, int64_t n_sample,
                   T* g, T* H, bool accumulate);
// F_i <- F_i - (g_i + l1 sign(F_i) + l2 F_i) S(H_i + l2_diag I); clamp. H: (rows x k x k) or shared (h_stride 0)
template <typename T>
void newton_solve_rows(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* g, const T* H, int64_t h_stride,
                       double l1, double l2, double l2_diag, double pert, bool non_negative, bool known_pd = false,
                       const double* Hbase = nullptr);   // Hbase: shared k x k float64 part added to every H_i in double
// shared Hessian given in float64 as h_scale * G (+ l2_diag I): the clamped inverse (k x k float64, arena 7 of ctx) ...
double* shared_inverse64(pycmf_ctx* ctx, int64_t k, const double* G64, double h_scale, double l2_diag, double pert);
// ... and its application to every row: F_i <- F_i - (g_i + l1 sign(F_i) + l2 F_i) Hinv ; clamp
template <typename T>
void apply_shared_inverse(pycmf_ctx* ctx, int64_t rows, int64_t k, T* F, const T* g, const double* Hinv, double l1,
                          double l2, bool non_negative);
// newton_small.cu : fused warp-per-row finish of the V update for k <= 32 and a small label factor; false = not eligible
template <typename T>
bool newton_finish_small(pycmf_ctx* ctx, int64_t rows, int64_t l, int64_t k, T* F, const T* Z, const T* Y, int64_t ldy,
                         int y_link, double wy, const T* gx, const void* Hx, bool hx_per_row, double l1, double l2,
                         double l2_diag, double pert, bool non_negative);   // Hx: per row (T) or shared (k x k FLOAT64)
// warp-per-matrix clamped solve for k <= 32 (false = not eligible). MODE 0: x = S(.) g ; MODE 1: Newton row update
template <typename T, int MODE>
bool safe_solve_small(pycmf_ctx* ctx, int64_t batch, int64_t k, const T* H, int64_t h_stride, const T* g, T* out,
                      double l1, double l2, double l2_diag, double pert, bool non_negative, double h_scale, bool known_pd);
void safe_solve_f64(pycmf_ctx* ctx, int64_t batch, int64_t k, const double* H, int64_t h_stride,
                    const double* g, double* x, double pert);
void sample_indices(pycmf_ctx* ctx, int64_t rows, int64_t row0, int64_t N, int64_t n_sample, uint64_t seed,
                    uint64_t stream_id, int64_t lo, int64_t hi, int32_t* idx);

}  // namespace pycmf

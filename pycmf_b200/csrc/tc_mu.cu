// tcgen05 / TMEM / TMA dense contractions of the multiplicative-update solver for n_components = 64 .. 256
// (fp32 data, 3xTF32 on the tensor cores):
//   LEFT  : out = X   Q     MU numerator X V      (cmf_solvers.py:232; also Y Z, :244)
//   RIGHT : out = X^T Q     MU numerator X^T U    (cmf_solvers.py:244)
// With k = 256 these two GEMMs are 4 n d k flop per iteration (1.0e13 at C5) and the only part of the MU
// step that is not factor-sized.
//
// A persistent CTA (one per SM) walks a range of 128 x 32 tiles of X (128 "own" rows: rows of X for LEFT, columns of X
// for RIGHT; 32 "other" rows per tile), linearised as in tc_resid.cu.  Per tile:
//   TMA (warp 16) : X tile (128 x 32 fp32, SWIZZLE_128B) into a 4-deep ring  -- the HBM stream
//   TMA (warp 17) : the K-major tile of Q^T (k x 32, tf32 hi and lo parts, 2 x k x 128 B) into its own ring of
//                   2 (k = 256) .. 4 slots -- L2 traffic, and the larger of the two: k = 256 moves 64 KB of Q^T per
//                   16 KB of X, which is what bounds this kernel ("MU wide kernel")
//   warps 0 .. 15 : X tile shared memory -> registers -> tf32 split (hi = truncation, lo = x - hi, both exact) ->
//                   tcgen05.st into one of four R buffers in TENSOR MEMORY (for RIGHT this is also the transposition:
//                   an MN-major tf32 A operand cannot be fed from shared memory)
//   warp 18       : 12 tcgen05.mma (TS form: A = R from tensor memory, B = Q^T tile), M = 128, N = k, K = 8:
//                   lo*hi, hi*lo first, hi*hi last, all into OUT[128 x k] in tensor memory
// OUT is read every CHAIN tiles by the 16 converter warps (tensor-memory accumulation truncates; the chains are added
// in fp32 registers, round to nearest) and written once per (CTA, own tile) as a partial; tc_mu_reduce_kernel adds the
// partials of an own tile in a fixed order (deterministic).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pycmf {
namespace {

constexpr int OWN = 128;           // own rows per CTA (UMMA M)
constexpr int KS = 32;             // other rows per tile (four K = 8 steps; one 128-byte swizzle span)
constexpr int NX = 4;              // X ring slots
constexpr int CONV_WARPS = 16;
constexpr int NTHREADS = 32 * CONV_WARPS + 96;
constexpr int W_TMAX = CONV_WARPS, W_TMAQ = CONV_WARPS + 1, W_MMA = CONV_WARPS + 2;
constexpr uint32_t X_BYTES = OWN * KS * 4;          // 16 KB
constexpr int TMEM_COLS = 512;

template <int KN> struct Cfg {
    static_assert(KN % 64 == 0 && KN >= 64 && KN <= 256, "n_components must be 64, 128, 192 or 256");
    static constexpr int NQ = KN > 128 ? 2 : 4;                    // Q^T ring slots
    static constexpr uint32_t QPART = uint32_t(KN) * 128u;           // one tf32 part of the Q^T tile
    static constexpr uint32_t QSLOT = 2u * QPART;
    static constexpr uint32_t x0 = 0, q0 = NX * X_BYTES, bars = q0 + NQ * QSLOT, total = bars + 256;
    static constexpr int CW = KN / 4;                               // OUT columns per converter warp
    static constexpr int NR = 4;                                    // R buffers in tensor memory
    static constexpr int TM_OUT = 0, TM_R = KN;                      // R[NR] : NR x (32 hi + 32 lo) columns
    static_assert(KN + NR * 2 * KS <= TMEM_COLS, "tensor memory budget");
    // barrier slots
    static constexpr int XFULL0 = 0, XEMPTY0 = XFULL0 + NX, QFULL0 = XEMPTY0 + NX, QEMPTY0 = QFULL0 + NQ,
                         RFULL0 = QEMPTY0 + NQ, RFREE0 = RFULL0 + NR, OUTFULL = RFREE0 + NR, OUTEMPTY = OUTFULL + 1,
                         NBARS = OUTEMPTY + 1;
    static_assert(NBARS * 8 + 16 <= 256, "barrier region too small");
};

struct MuParams {
    int64_t own_n, oth_n;
    int64_t own_tiles;          // tiles of OWN along the own dimension
    int n_oth_tiles;            // T : tiles of KS along the other dimension
    int chunk_tiles;            // Tc: other tiles per unit
    int n_chunks;               // ceil(T / Tc)
    int64_t n_units;            // own_tiles x n_chunks, unit u = chunk * own_tiles + own_tile (chunk-major)
    int chain;                  // accumulation chain cap in tiles
    float* part;                // [own tile][chunk][OWN x KN]
    long long* trace;           // optional pipeline trace of CTA 0: [event][tile] clock64 stamps (diagnostics)
};

constexpr int TRACE_TILES = 96;
enum MuTrace { TR_MMA_TOP = 0, TR_MMA_Q, TR_MMA_R, TR_MMA_ISSUED, TR_X_ISSUE, TR_Q_ISSUE, TR_CONV_X, TR_CONV_RFREE, TR_CONV_DONE };
#define MU_TRACE(ev, it)                                                                        \
    do {                                                                                        \
        if (prm.trace != nullptr && blockIdx.x == 0 && (it) < TRACE_TILES)                      \
            prm.trace[(ev) * TRACE_TILES + (it)] = clock64();                                   \
    } while (0)

// mbarrier wait with a watchdog: a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t spins = 0;; spins++) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if ((spins & 1023u) == 1023u) {          // ~2 s at 2 GHz
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    }
}

__device__ __forceinline__ void umma_tf32_ts_rt(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(0u), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

// MODE 0 = LEFT (own = rows of X), 1 = RIGHT (own = columns of X).
// Schedule: a unit is (own tile o, chunk ch of Tc other tiles); units are numbered chunk-major and dealt round-robin
// (CTA c takes u = c, c + grid, ...), so that at any time all CTAs work inside one or two neighbouring chunks of the
// other dimension: the Q^T tiles of a chunk (Tc x k x 256 B) are fetched from HBM once and then hit in L2 for every
// own tile.  (With contiguous tile ranges per CTA every CTA sweeps a different part of Q^T at any moment: at k = 256
// the Q^T parts of a 200k-row factor are 410 MB and every tile's 64 KB would come from HBM.)
template <int MODE, int KN>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_mu_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_qt_hi,
             const __grid_constant__ CUtensorMap tm_qt_lo, const MuParams prm) {
    using C = Cfg<KN>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + C::bars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + C::bars + C::NBARS * 8);
    auto bar = [&](int i) { return bars + 8u * uint32_t(i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = prm.n_oth_tiles, Tc = prm.chunk_tiles;
    const int chain = prm.chain;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NX; s++) { mbar_init(bar(C::XFULL0 + s), 1); mbar_init(bar(C::XEMPTY0 + s), CONV_WARPS); }
        for (int s = 0; s < C::NQ; s++) { mbar_init(bar(C::QFULL0 + s), 1); mbar_init(bar(C::QEMPTY0 + s), 1); }
        for (int s = 0; s < C::NR; s++) { mbar_init(bar(C::RFULL0 + s), CONV_WARPS); mbar_init(bar(C::RFREE0 + s), 1); }
        mbar_init(bar(C::OUTFULL), 1);
        mbar_init(bar(C::OUTEMPTY), CONV_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == W_TMAX) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(uint32_t(TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == W_TMAX) {
        // =============================== X producer (HBM stream) ===============================
        int it = 0;
        for (int64_t u = blockIdx.x; u < prm.n_units; u += gridDim.x) {
            const int ch = int(u / prm.own_tiles);
            const int own0 = int(u - int64_t(ch) * prm.own_tiles) * OWN;
            const int t0 = ch * Tc, t1 = min(T, t0 + Tc);
            for (int t = t0; t < t1; t++, it++) {
                const int s = it % NX;
                mbar_wait_wd(bar(C::XEMPTY0 + s), (uint32_t(it / NX) & 1u) ^ 1u);
                if (elect_one()) {
                    MU_TRACE(TR_X_ISSUE, it);
                    const uint32_t dst = base + C::x0 + uint32_t(s) * X_BYTES;
                    const int oth0 = t * KS;
                    mbar_expect_tx(bar(C::XFULL0 + s), X_BYTES);
                    if (MODE == 0) {
                        tma_load_2d(dst, &tm_x, bar(C::XFULL0 + s), oth0, own0);       // 128 own rows x 32 other columns
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; b++)                                    // 32 other rows x 4 x 32 own columns
                            tma_load_2d(dst + uint32_t(b) * (KS * 128), &tm_x, bar(C::XFULL0 + s), own0 + 32 * b, oth0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == W_TMAQ) {
        // =============================== Q^T producer (L2 traffic) ===============================
        int it = 0;
        for (int64_t u = blockIdx.x; u < prm.n_units; u += gridDim.x) {
            const int ch = int(u / prm.own_tiles);
            const int t0 = ch * Tc, t1 = min(T, t0 + Tc);
            for (int t = t0; t < t1; t++, it++) {
                const int s = it % C::NQ;
                mbar_wait_wd(bar(C::QEMPTY0 + s), (uint32_t(it / C::NQ) & 1u) ^ 1u);
                if (elect_one()) {
                    MU_TRACE(TR_Q_ISSUE, it);
                    const uint32_t dst = base + C::q0 + uint32_t(s) * C::QSLOT;
                    mbar_expect_tx(bar(C::QFULL0 + s), C::QSLOT);
                    tma_load_2d(dst, &tm_qt_hi, bar(C::QFULL0 + s), t * KS, 0);
                    tma_load_2d(dst + C::QPART, &tm_qt_lo, bar(C::QFULL0 + s), t * KS, 0);
                }
                __syncwarp();
            }
        }
    } else if (warp == W_MMA) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc(OWN, KN, 0, 0);
        int it = 0, chains_done = 0;
        for (int64_t u = blockIdx.x; u < prm.n_units; u += gridDim.x) {
            const int ch = int(u / prm.own_tiles);
            const int t0 = ch * Tc, t1 = min(T, t0 + Tc);
            int cpos = 0;
            for (int t = t0; t < t1; t++, it++) {
                const int qs = it % C::NQ, rb = it % C::NR;
                const bool first = cpos == 0, last = cpos == chain - 1 || t == t1 - 1;
                if (lane == 0) MU_TRACE(TR_MMA_TOP, it);
                mbar_wait_wd(bar(C::QFULL0 + qs), uint32_t(it / C::NQ) & 1u);
                if (lane == 0) MU_TRACE(TR_MMA_Q, it);
                mbar_wait_wd(bar(C::RFULL0 + rb), uint32_t(it / C::NR) & 1u);
                if (lane == 0) MU_TRACE(TR_MMA_R, it);
                if (first && chains_done > 0) mbar_wait_wd(bar(C::OUTEMPTY), uint32_t(chains_done - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t q = base + C::q0 + uint32_t(qs) * C::QSLOT;
                    const uint64_t qh = make_desc(q, 16, 1024), ql = make_desc(q + C::QPART, 16, 1024);
                    const uint32_t r_hi = tmem + uint32_t(C::TM_R + rb * 2 * KS), r_lo = r_hi + KS;
                    const uint32_t d = tmem + uint32_t(C::TM_OUT);
#pragma unroll
                    for (int term = 0; term < 3; term++) {           // lo*hi, hi*lo, hi*hi
                        const uint32_t ra = term == 0 ? r_lo : r_hi;
                        const uint64_t qa = term == 1 ? ql : qh;
#pragma unroll
                        for (int kk = 0; kk < KS / 8; kk++)
                            umma_tf32_ts_rt(d, ra + uint32_t(kk * 8), qa + uint64_t(kk * 2), idesc,
                                            (first && term == 0 && kk == 0) ? 0u : 1u);
                    }
                    umma_commit(bar(C::QEMPTY0 + qs));      // Q^T slot read
                    umma_commit(bar(C::RFREE0 + rb));       // R buffer read
                    if (last) umma_commit(bar(C::OUTFULL));
                    MU_TRACE(TR_MMA_ISSUED, it);
                }
                __syncwarp();
                if (last) { chains_done++; cpos = 0; }
                else cpos++;
            }
        }
    } else if (warp < CONV_WARPS) {
        // ================================ converters (warps 0 .. 15) ================================
        const int q = warp & 3;                   // TMEM lane quadrant (hardware: warp id % 4)
        const int c = warp >> 2;                  // 8-column chunk of the tile / CW-column chunk of OUT
        const int i = q * 32 + lane;              // own row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + (uint32_t(q * 32) << 16);
        // loop-invariant byte offsets of this thread's 8 X elements inside a ring slot (other indices j = 8c .. 8c + 7)
        uint32_t xoff[MODE == 0 ? 2 : 8];
        if (MODE == 0) {
#pragma unroll
            for (int h = 0; h < 2; h++) xoff[h] = uint32_t(i * 128 + (((2 * c + h) ^ (i & 7)) << 4));
        } else {
            const int blk = i >> 5, ch = (i & 31) >> 2, w = i & 3;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = 8 * c + e;
                xoff[e] = uint32_t(blk * (KS * 128) + j * 128 + ((ch ^ (j & 7)) << 4) + w * 4);
            }
        }
        float acc[C::CW];
#pragma unroll
        for (int e = 0; e < C::CW; e++) acc[e] = 0.0f;
        int chains_seen = 0;
        bool pend = false, pend_unit_ends = false;
        int64_t pend_slot = 0;                    // (own tile * n_chunks + chunk) of the pending chain
        // OUT -> register accumulators (fp32, round to nearest); at the end of a unit: accumulators -> the unit's partial
        auto flush_chain = [&](int64_t slot, bool unit_ends) {
            mbar_wait_wd(bar(C::OUTFULL), uint32_t(chains_seen) & 1u);
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < C::CW / 16; b++) {
                float o[16];
                tmem_ld16(lane_addr + uint32_t(C::TM_OUT + c * C::CW + b * 16), o);
#pragma unroll
                for (int e = 0; e < 16; e++) acc[b * 16 + e] += o[e];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(C::OUTEMPTY));
            chains_seen++;
            if (unit_ends) {
                const int64_t own_idx = (slot / prm.n_chunks) * OWN + i;
                if (own_idx < prm.own_n) {
                    float4* dst = reinterpret_cast<float4*>(prm.part + (slot * OWN + i) * KN + c * C::CW);
#pragma unroll
                    for (int e = 0; e < C::CW / 4; e++)
                        dst[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
                }
#pragma unroll
                for (int e = 0; e < C::CW; e++) acc[e] = 0.0f;
            }
        };
        int it = 0;
        for (int64_t u = blockIdx.x; u < prm.n_units; u += gridDim.x) {
            const int ch = int(u / prm.own_tiles);
            const int64_t slot = (u - int64_t(ch) * prm.own_tiles) * prm.n_chunks + ch;
            const int t0 = ch * Tc, t1 = min(T, t0 + Tc);
            const bool last_unit = u + gridDim.x >= prm.n_units;
            int cpos = 0;
            for (int t = t0; t < t1; t++, it++) {
                const int s = it % NX, rb = it % C::NR;
                mbar_wait_wd(bar(C::XFULL0 + s), uint32_t(it / NX) & 1u);
                if (threadIdx.x == 0) MU_TRACE(TR_CONV_X, it);
                const unsigned char* xs = gen + C::x0 + s * X_BYTES;
                float xv[8];
                if (MODE == 0) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 t4 = *reinterpret_cast<const float4*>(xs + xoff[h]);
                        xv[4 * h] = t4.x; xv[4 * h + 1] = t4.y; xv[4 * h + 2] = t4.z; xv[4 * h + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) xv[e] = *reinterpret_cast<const float*>(xs + xoff[e]);
                }
                // the X slot is free as soon as every converter warp holds its elements in registers
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(C::XEMPTY0 + s));
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    hi[e] = __uint_as_float(__float_as_uint(xv[e]) & 0xffffe000u);
                    lo[e] = xv[e] - hi[e];
                }
                // R buffer it % NR: the MMAs of tile it - NR must have read it
                mbar_wait_wd(bar(C::RFREE0 + rb), (uint32_t(it / C::NR) & 1u) ^ 1u);
                if (threadIdx.x == 0) MU_TRACE(TR_CONV_RFREE, it);
                tc_fence_after();
                const uint32_t r_hi = lane_addr + uint32_t(C::TM_R + rb * 2 * KS + c * 8);
                tmem_st8(r_hi, hi);
                tmem_st8(r_hi + KS, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(C::RFULL0 + rb));
                if (threadIdx.x == 0) MU_TRACE(TR_CONV_DONE, it);
                const bool unit_ends = t == t1 - 1;
                const bool last = cpos == chain - 1 || unit_ends;
                // end of a chain: OUT -> registers, deferred by one tile (the chain's last MMAs have completed by then)
                if (pend) flush_chain(pend_slot, pend_unit_ends);
                pend = false;
                if (last) {
                    if (last_unit && unit_ends) flush_chain(slot, true);
                    else { pend = true; pend_slot = slot; pend_unit_ends = unit_ends; }
                    cpos = 0;
                } else {
                    cpos++;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == W_TMAX) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(TMEM_COLS)) : "memory");
    }
}

// out[r][:] = sum over the chunks of own tile r / OWN of part[tile][chunk][r % OWN][:]   (fixed order: deterministic)
__global__ void tc_mu_reduce_kernel(int64_t own_n, int kn, int n_chunks, const float* __restrict__ part,
                                    float* __restrict__ out) {
    const int64_t e4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;     // one float4 of the output per thread
    const int k4 = kn / 4;
    if (e4 >= own_n * k4) return;
    const int64_t r = e4 / k4, c4 = e4 % k4;
    const int64_t tile = r / OWN;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < n_chunks; e++) {
        const float4 v = *reinterpret_cast<const float4*>(part + ((tile * n_chunks + e) * OWN + (r % OWN)) * kn + c4 * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * kn + c4 * 4) = s;
}

// transposed tf32 parts of a factor F (rows x k): hi_t, lo_t (k x ldt), hi = rna_tf32(F), lo = rna_tf32(F - hi)
__global__ void split_t_kernel(int64_t rows, int k, int64_t ldt, const float* __restrict__ x, float* __restrict__ hi_t,
                               float* __restrict__ lo_t) {
    __shared__ float th[32][33], tl[32][33];
    const int64_t r0 = int64_t(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int64_t r = r0 + rr;
        float h = 0.f, l = 0.f;
        if (r < rows && c0 + threadIdx.x < k) {
            const float v = x[r * k + c0 + threadIdx.x];
            h = tf32_rna(v);
            l = tf32_rna(v - h);
        }
        th[rr][threadIdx.x] = h;
        tl[rr][threadIdx.x] = l;
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int64_t r = r0 + threadIdx.x;
        if (r < ldt && c0 + cc < k) {
            hi_t[int64_t(c0 + cc) * ldt + r] = th[threadIdx.x][cc];
            lo_t[int64_t(c0 + cc) * ldt + r] = tl[threadIdx.x][cc];
        }
    }
}

// Chunks of the other dimension per own tile: enough units for a balanced round-robin deal over the CTAs, a Q^T chunk
// that stays in L2, and as little partial-output traffic as possible.  Cost model per CTA in clocks: tensor pipe
// 6.5 k per tile (measured), HBM 22 B per clock per SM (6.5 TB/s / 148 / 1.97 GHz).
int pick_chunks(pycmf_ctx* ctx, int64_t own_tiles, int64_t T, int kn, int64_t n_cta) {
    if (ctx->tc_max_splits > 0) return int(std::min<int64_t>(T, ctx->tc_max_splits));
    const double l2_budget = 32e6;                                   // bytes of Q^T (hi + lo) per chunk
    const int64_t c_min = std::max<int64_t>(1, ceil_div(int64_t(double(T) * kn * 256.0), int64_t(l2_budget)));
    int best = int(std::min<int64_t>(c_min, T));
    double best_cost = 1e300;
    for (int64_t c = c_min; c <= std::min<int64_t>(T, c_min + 96); c++) {
        const int64_t tc = ceil_div(T, c), chunks = ceil_div(T, tc);
        const int64_t waves = ceil_div(own_tiles * chunks, n_cta);
        const double tensor = double(tc) * 6.5 * kn;
        const double hbm = (double(tc) * X_BYTES + 2.0 * OWN * kn * 4) / 22.0;
        const double cost = double(waves) * std::max(tensor, hbm);
        if (cost < best_cost * 0.995) { best_cost = cost; best = int(chunks); }
    }
    return best;
}

template <int MODE, int KN>
void launch_mu(pycmf_ctx* ctx, int64_t own_n, int64_t oth_n, const float* X, int64_t x_rows, int64_t x_cols, int64_t ldx,
               const float* qt_hi, const float* qt_lo, int64_t ldt, float* out, const char* family) {
    using C = Cfg<KN>;
    const int64_t own_tiles = ceil_div(own_n, OWN), T = ceil_div(oth_n, KS);
    int64_t n_cta = ctx->num_sms;
    if (ctx->tc_ctas > 0) n_cta = std::min<int64_t>(n_cta, ctx->tc_ctas);
    const int want_chunks = pick_chunks(ctx, own_tiles, T, KN, n_cta);
    const int64_t tc = ceil_div(T, want_chunks);
    const int n_chunks = int(ceil_div(T, tc));
    const int64_t n_units = own_tiles * n_chunks;
    n_cta = std::min<int64_t>(n_cta, n_units);
    // tensor-memory accumulation truncates: at most 16 tiles (K = 512, 192 MMAs) per chain by default
    const int chain = ctx->tc_chain > 0 ? ctx->tc_chain : 16;
    const int promo = ctx->tc_x_promotion;
    CUtensorMap tm_x = make_map(X, x_rows, x_cols, ldx, MODE == 0 ? OWN : KS,
                                promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                             : (promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                            : (promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_128B)));
    CUtensorMap tm_hi = make_map(qt_hi, KN, oth_n, ldt, KN);
    CUtensorMap tm_lo = make_map(qt_lo, KN, oth_n, ldt, KN);
    MuParams prm;
    prm.own_n = own_n;
    prm.oth_n = oth_n;
    prm.own_tiles = own_tiles;
    prm.n_oth_tiles = int(T);
    prm.chunk_tiles = int(tc);
    prm.n_chunks = n_chunks;
    prm.n_units = n_units;
    prm.chain = chain;
    prm.part = static_cast<float*>(scratch(ctx, 0, size_t(n_units) * OWN * KN * sizeof(float)));
    prm.trace = nullptr;
    if (ctx->tc_trace && family == nullptr) {      // the passes over X only
        prm.trace = static_cast<long long*>(scratch(ctx, 2, sizeof(long long) * tc_trace_words()));
        PYCMF_CUDA(cudaMemsetAsync(prm.trace, 0, sizeof(long long) * tc_trace_words(), ctx->stream));
    }
    auto kern = tc_mu_kernel<MODE, KN>;
    const size_t smem = C::total + 1024;
    PYCMF_CHECK(smem <= size_t(ctx->max_smem_optin), "tc mu pass: shared memory budget exceeded");
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    {
        Timed timer(ctx, family != nullptr ? family : (MODE == 0 ? "tc_xv" : "tc_xtu"));
        kern<<<(unsigned)n_cta, NTHREADS, smem, ctx->stream>>>(tm_x, tm_hi, tm_lo, prm);
        PYCMF_LAUNCH_CHECK(ctx);
    }
    tc_mu_reduce_kernel<<<(unsigned)ceil_div(own_n * (KN / 4), 256), 256, 0, ctx->stream>>>(own_n, KN, n_chunks, prm.part,
                                                                                          out);
    PYCMF_LAUNCH_CHECK(ctx);
}

template <int KN>
void xmul_k(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, const float* X, int64_t ldx, const float* qt_hi,
            const float* qt_lo, int64_t ldt, float* out, const char* family) {
    if (!trans) launch_mu<0, KN>(ctx, rows, cols, X, rows, cols, ldx, qt_hi, qt_lo, ldt, out, family);
    else launch_mu<1, KN>(ctx, cols, rows, X, rows, cols, ldx, qt_hi, qt_lo, ldt, out, family);
}

}  // namespace

bool tc_mu_eligible(pycmf_ctx* ctx, int64_t rows, int64_t cols, int64_t k, const float* X, int64_t ldx, bool trans_t) {
    if (ctx->dense_path == 0 || trans_t || X == nullptr) return false;
    if (k != 64 && k != 128 && k != 192 && k != 256) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0 || (ldx % 4) != 0) return false;
    if (rows * cols < (int64_t(1) << 16)) return false;
    return rows >= 1 && cols >= 1 && rows < (int64_t(1) << 31) && cols < (int64_t(1) << 31);
}

// out = X Q   (trans == false: X rows x cols, Q cols x k, out rows x k)
//       X^T Q (trans == true : Q rows x k, out cols x k)
void tc_mu_xmul(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, int64_t k, const float* X, int64_t ldx,
                const float* Q, float* out, const char* family) {
    const int64_t qn = trans ? rows : cols;
    const int64_t ldt = (qn + 3) & ~int64_t(3);
    float* buf = static_cast<float*>(scratch(ctx, 3, sizeof(float) * size_t(2) * k * ldt));
    float *hi_t = buf, *lo_t = buf + size_t(k) * ldt;
    split_t_kernel<<<dim3((unsigned)ceil_div(ldt, 32), (unsigned)(k / 32)), dim3(32, 8), 0, ctx->stream>>>(qn, int(k), ldt, Q,
                                                                                                        hi_t, lo_t);
    PYCMF_LAUNCH_CHECK(ctx);
    switch (k) {
        case 64: xmul_k<64>(ctx, trans, rows, cols, X, ldx, hi_t, lo_t, ldt, out, family); break;
        case 128: xmul_k<128>(ctx, trans, rows, cols, X, ldx, hi_t, lo_t, ldt, out, family); break;
        case 192: xmul_k<192>(ctx, trans, rows, cols, X, ldx, hi_t, lo_t, ldt, out, family); break;
        case 256: xmul_k<256>(ctx, trans, rows, cols, X, ldx, hi_t, lo_t, ldt, out, family); break;
        default: PYCMF_CHECK(false, "tc_mu_xmul: unsupported n_components");
    }
}

}  // namespace pycmf

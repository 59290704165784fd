// tcgen05 / TMEM / TMA version of the dense streaming passes over X (fp32 data, TF32 tensor cores).
//
// One kernel template covers the four dense hot loops of the fit (n_components == 32):
//   RESID, LEFT  : outL = (f(A B^T) - X)   B     Newton U gradient   (cmf_solvers.py:399-400)
//   RESID, RIGHT : outR = (f(A B^T) - X)^T A     Newton V gradient   (cmf_solvers.py:436-440)
//   COPY,  LEFT  : out  = X   B                  MU numerator X V    (cmf_solvers.py:232)
//   COPY,  RIGHT : out  = X^T A                  MU numerator X^T U  (cmf_solvers.py:244)
//
// A persistent CTA (one per SM) owns a range of 128 x 64 tiles of X (128 "own" rows: rows of X for LEFT, columns of X
// for RIGHT; 64 "other" rows per tile).  Per tile:
//   TMA        : X tile (128 x 64 fp32, SWIZZLE_128B) + the 64 x 32 factor tile Q (tf32 hi / lo parts) into a
//                4-stage shared-memory ring (48 KB per stage)
//   transpose  : the epilogue warps turn the Q tile into the K-major Q^T tile GEMM2 needs (two slots), on chip, each
//                warp 1/16 of it next to its R stores: loading Q^T through TMA as well doubled the factor traffic and
//                cost the ring its fourth stage; a single transposer warp needed 1700 clk per tile (measured)
//   tcgen05.mma: S[128 x 64]  = P Q^T     GEMM1, A operand P resident in TENSOR MEMORY for the whole own tile
//   epilogue   : tcgen05.ld S -> f() -> minus X -> R (tf32 hi / lo) -> tcgen05.st back into TENSOR MEMORY
//   tcgen05.mma: OUT[128 x 32] += R Q     GEMM2, A operand R read from tensor memory, B = K-major Q^T tile
// Neither U V^T nor the residual ever touches HBM -- or shared memory: with N = 32..64 the MMAs have too little
// operand reuse for shared-memory A operands (measured, profiles/r01_tcgen05_mma_rate.txt: SS form 32 + N/4 clk per
// MMA, TS form N/2), so both A operands live in TMEM and shared memory only carries the rings.
// 3xTF32: hi*hi + hi*lo + lo*hi.  Two S buffers and two R buffers in TMEM; GEMM1 and GEMM2 are issued by two different
// warps (measured: one warp doing the waits, 36 MMAs and the commits of a tile needs ~2600 clk per tile and was the
// bottleneck of the whole pass; the tensor pipe itself needs 768).
//
// Warp roles (640 threads): warps 0-15 = epilogue (warp w works on TMEM lane quadrant w % 4 and on the 16-column chunk
// w / 4 of the 64-wide tile, so every SM sub-partition has four epilogue warps); warp 16 = TMA producer, 17 = GEMM2
// issuer (hi*hi chain into OUT), 18 = GEMM2 issuer of the correction terms (into OUT2), 19 = GEMM1 issuer.  Several issuing warps because one thread cannot feed the tensor pipe with N = 32 MMAs
// (profiles/r01_mma_issue_mix.txt: one warp 1035 clk per tile's worth of MMAs, two warps 614).  The control warps have
// the highest warp ids on purpose: the issue arbiter prefers high warp ids, and a late MMA issue stalls everything.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace pycmf {
namespace {

constexpr int OWN = 128;      // own rows per CTA  (UMMA M)
constexpr int OTH = 64;       // other rows per tile (GEMM1 N, GEMM2 K)
constexpr int KC = 32;        // n_components handled by this kernel (one 128-byte swizzle span)
constexpr int NSTAGE = 4;
constexpr int NQT = 2;             // Q^T slots (produced on chip from the Q tile, consumed by GEMM2)
constexpr int PREFETCH_DIST = 8;   // X tiles requested into L2 ahead of the shared-memory ring
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 32 * EPI_WARPS + 128;
constexpr int W_TMA = EPI_WARPS, W_G2 = EPI_WARPS + 1, W_G2B = EPI_WARPS + 2, W_G1 = EPI_WARPS + 3;
static_assert(NQT == 2, "G2DONE doubles as the R-buffer and the Q^T-slot release: both rings must have two slots");

// tensor-memory columns (512 x 128 lanes x 32 bit, all allocated: one CTA per SM)
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0;            // S[2]   : 2 x 64 columns
constexpr int TM_OUT = 128;        // OUT    : 32 columns
constexpr int TM_P_HI = 160;       // P hi   : 32 columns (A operand of GEMM1)
constexpr int TM_P_LO = 192;       // P lo   : 32 columns
constexpr int TM_OUT2 = 224;       // OUT2   : 32 columns, accumulates only the small hi*lo + lo*hi terms of GEMM2
constexpr int TM_R = 256;          // R[2]   : 2 x (64 hi + 64 lo) columns (A operand of GEMM2)

constexpr uint32_t Q_BYTES = OTH * 128;          //  8 KB (64 rows x 32 fp32; also 2 x 32 x 32 for Q^T)
constexpr uint32_t X_BYTES = OWN * OTH * 4;      // 32 KB

struct SmemLayout {
    // all tile buffers are 1024-byte aligned (SWIZZLE_128B atoms)
    // stage: Q hi / lo (64 x 32, K-major for GEMM1; source of the transposer), X tile
    static constexpr uint32_t q_hi = 0, q_lo = Q_BYTES, x = 2 * Q_BYTES;
    static constexpr uint32_t stage_bytes = 2 * Q_BYTES + X_BYTES;
    static constexpr uint32_t stage0 = 0;
    // Q^T slot: hi / lo, each 32 x 64 as two 32 x 32 K-blocks (GEMM2's B operand)
    static constexpr uint32_t qt0 = stage0 + NSTAGE * stage_bytes;
    static constexpr uint32_t qt_hi = 0, qt_lo = Q_BYTES, qt_bytes = 2 * Q_BYTES;
    static constexpr uint32_t bars = qt0 + NQT * qt_bytes;
    static constexpr uint32_t total = bars + 256;
};

// barrier slots (8 bytes each) inside the `bars` region
enum Bar { FULL0 = 0, EMPTY0 = FULL0 + NSTAGE, SFULL0 = EMPTY0 + NSTAGE, RFULL0 = SFULL0 + 2,
           // RFULL(t & 1): R(t) and Q^T(t) published by the 16 epilogue warps
           G2DONE0 = RFULL0 + 2,      // GEMM2(t) complete: R buffer t & 1 and Q^T slot t & 1 are free again
           PFULL = G2DONE0 + 2, OUTFULL, OUTEMPTY, NBARS };
static_assert(NBARS * 8 + 32 <= 256, "barrier region too small");


struct Params {
    int64_t own_n, oth_n;       // extents of the own / other dimension
    int64_t n_oth_tiles;        // T: tiles of OTH along the other dimension
    int64_t n_tiles;            // G: own tiles x T, linearised g = own_tile * T + t
    int chain;                  // accumulation chain cap in tiles (TMEM accumulation is not round-to-nearest)
    int link;
    const float *p_hi, *p_lo;   // tf32 parts of the own-side factor (own_n x 32), RESID only
    const float* x;             // X itself (rows x cols, row stride ldx) for the L2 prefetch warp
    int64_t ldx, x_rows, x_cols;
    int pf_mode;                // 0 none, 2 TMA L2 prefetch of X from the producer
    float* part;                // partial outputs: [own tile][entry][OWN x 32]
    int max_entries;            // entries per own tile in `part`
    double* sq_part;            // per-CTA partial of sum R^2 (may be null)
    long long* trace;           // optional pipeline trace of CTA 0: [event][tile] clock64 stamps (diagnostics)
};

// Persistent tiling: CTA c of N walks the linearised tiles [c G / N, (c + 1) G / N).  first_cta(g) is the CTA that
// owns tile g (inverse of the floor partition); the partial of own tile o written by CTA c goes to entry
// c - first_cta(o T), and the reduction kernel sums entries 0 .. first_cta(o T + T - 1) - first_cta(o T).
__host__ __device__ __forceinline__ int64_t tile_begin(int64_t c, int64_t G, int64_t N) { return (c * G) / N; }
__host__ __device__ __forceinline__ int64_t first_cta(int64_t g, int64_t G, int64_t N) { return ((g + 1) * N - 1) / G; }

constexpr int TRACE_TILES = 96;
enum TraceEvent { TR_TMA_ISSUE = 0, TR_G1_ISSUE, TR_S_SEEN, TR_R_DONE, TR_G2_ISSUE, TR_EMPTY_SEEN, TR_FULL_SEEN_EPI,
                  TR_CTA = 7,          // [0] kernel entry, [1] after the setup barrier, [2] CTA done
                  TR_NEVENTS = 10 };
#define TC_TRACE(ev, it)                                                                        \
    do {                                                                                        \
        if (prm.trace != nullptr && blockIdx.x == 0 && (it) < TRACE_TILES)                      \
            prm.trace[(ev) * TRACE_TILES + (it)] = clock64();                                   \
    } while (0)

// MODE 0 = LEFT (own = rows of X), 1 = RIGHT (own = columns of X).  RESID: R = f(S) - X, else R = X.
// One persistent CTA per SM.  Tile i of the CTA is g = g0 + i -> (own tile o = g / T, other tile t = g % T).
// A chain (one accumulation in OUT) starts at i == 0 or t % chain == 0 and ends at the CTA's last tile, at
// (t + 1) % chain == 0 or at t == T - 1; the epilogue warps add the chains of one own tile in registers (fp32,
// round to nearest) and write one partial per (CTA, own tile).
template <int MODE, bool RESID, int NSPLIT, bool LOGIT>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_pass_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
               const __grid_constant__ CUtensorMap tm_qt_hi, const __grid_constant__ CUtensorMap tm_qt_lo,
               const __grid_constant__ CUtensorMap tm_x, const Params prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // dynamic shared memory is only guaranteed 16-byte aligned: round up to 1024 for the swizzle atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + SmemLayout::bars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SmemLayout::bars + NBARS * 8);
    double* sq_slot = reinterpret_cast<double*>(gen + SmemLayout::bars + NBARS * 8 + 16);
    auto bar = [&](int i) { return bars + 8u * uint32_t(i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TC_TRACE(TR_CTA, 0);
    if (prm.trace != nullptr && threadIdx.x == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.trace[TR_NEVENTS * TRACE_TILES + 2 * blockIdx.x] = gt;
    }
    const int64_t T = prm.n_oth_tiles;
    const int64_t g0 = tile_begin(blockIdx.x, prm.n_tiles, gridDim.x);
    const int n_it = int(tile_begin(int64_t(blockIdx.x) + 1, prm.n_tiles, gridDim.x) - g0);
    const int chain = prm.chain;
    // schedule predicates, identical in every role; cpos = t % chain is carried incrementally (no divisions in the loops)
    auto first_of_chain = [&](int it, int cpos) { return it == 0 || cpos == 0; };
    auto last_of_chain = [&](int it, int t, int cpos) { return it == n_it - 1 || cpos == chain - 1 || t == T - 1; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(bar(FULL0 + s), 1); mbar_init(bar(EMPTY0 + s), 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar(SFULL0 + s), 1);
            mbar_init(bar(RFULL0 + s), EPI_WARPS);
            mbar_init(bar(G2DONE0 + s), NSPLIT == 3 ? 2 : 1);     // both GEMM2 issuers commit
        }
        mbar_init(bar(PFULL), EPI_WARPS);
        mbar_init(bar(OUTFULL), NSPLIT == 3 ? 2 : 1);
        mbar_init(bar(OUTEMPTY), 2 * 4);          // the eight warps that read OUT
        *sq_slot = 0.0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == W_TMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(uint32_t(TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) TC_TRACE(TR_CTA, 1);
    const int64_t o_first = g0 / T;
    const int t_first = int(g0 % T);               // the only 64-bit divisions: every role walks (own tile, t) incrementally
    const int c_first = t_first % chain;

    if (warp == W_TMA) {
        // =============================== TMA producer ===============================
        // The whole warp walks the loop (converged); one elected lane issues.
        int64_t o = o_first;
        int t = t_first;
        for (int it = 0; it < n_it; it++) {
            const int s = it % NSTAGE;
            const uint32_t ph = uint32_t(it / NSTAGE) & 1u;
            mbar_wait(bar(EMPTY0 + s), ph ^ 1u);
            if (elect_one()) {
                TC_TRACE(TR_EMPTY_SEEN, it);
                const uint32_t st = base + SmemLayout::stage0 + uint32_t(s) * SmemLayout::stage_bytes;
                const int own0 = int(o * OWN), oth0 = t * OTH;
                mbar_expect_tx(bar(FULL0 + s), (NSPLIT == 3 ? 2 * Q_BYTES : Q_BYTES) + X_BYTES);
                if (MODE == 0) {
                    // X tile: own rows x 64 other columns, two 32-column boxes of 128 rows
                    tma_load_2d(st + SmemLayout::x, &tm_x, bar(FULL0 + s), oth0, own0);
                    tma_load_2d(st + SmemLayout::x + OWN * 128, &tm_x, bar(FULL0 + s), oth0 + 32, own0);
                } else {
                    // X tile: 64 other rows x 128 own columns, four 32-column boxes of 64 rows
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        tma_load_2d(st + SmemLayout::x + uint32_t(b) * OTH * 128, &tm_x, bar(FULL0 + s), own0 + 32 * b, oth0);
                }
                tma_load_2d(st + SmemLayout::q_hi, &tm_q_hi, bar(FULL0 + s), 0, oth0);
                if (NSPLIT == 3) tma_load_2d(st + SmemLayout::q_lo, &tm_q_lo, bar(FULL0 + s), 0, oth0);
                TC_TRACE(TR_TMA_ISSUE, it);
                if (prm.pf_mode == 2) {
                    // TMA L2 prefetch PREFETCH_DIST tiles ahead (same own tile only: the box is clipped at the edge)
                    const int tp = t + PREFETCH_DIST;
                    if (tp < T) {
                        if (MODE == 0) {
                            tma_prefetch_2d(&tm_x, tp * OTH, own0);
                            tma_prefetch_2d(&tm_x, tp * OTH + 32, own0);
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; b++) tma_prefetch_2d(&tm_x, own0 + 32 * b, tp * OTH);
                        }
                    }
                }
            }
            __syncwarp();
            if (++t == T) { t = 0; o++; }
        }
    } else if (warp == W_G1 || warp == W_G2 || warp == W_G2B) {
        // =============================== MMA issuers ================================
        // Each of the two warps walks its schedule converged, with blocking waits; one elected lane issues the MMAs and
        // the commits (tcgen05.commit tracks the MMAs of the issuing thread: elect.sync always picks the same lane).
        // Under `lane == 0` ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop (~80 clk per MMA issued).
        constexpr uint32_t idesc1 = make_idesc(OWN, OTH, 0, 0);   // S   = P (TMEM) x Q (K-major)
        constexpr uint32_t idesc2 = make_idesc(OWN, KC, 0, 0);    // OUT = R (TMEM) x Q^T tile (K-major)
        // B descriptors: inside a tile only the 14-bit start-address field changes, by small multiples of 16 bytes that
        // cannot carry out of the field (all tiles live below 256 KB).
        if (warp == W_G1) {
            if (RESID) {
                // GEMM1(it) needs: Q(it) landed (FULL); S buffer it & 1 read by the epilogue of tile it - 2 (RFULL(it - 2):
                // this warp only observes that barrier); at the first tile of an own tile, P in tensor memory (PFULL).
                // Order inside a tile: the small correction terms hi*lo, lo*hi first, the four hi*hi MMAs last
                // (tensor-core accumulation into TMEM truncates: ~2^-24 |acc| lost per MMA).
                int t1 = t_first, p_phase = 0;
                for (int it = 0; it < n_it; it++) {
                    const int s = it % NSTAGE, sb = it & 1;
                    mbar_wait(bar(FULL0 + s), uint32_t(it / NSTAGE) & 1u);
                    if (it >= 2) mbar_wait(bar(RFULL0 + sb), uint32_t((it - 2) >> 1) & 1u);
                    if (it == 0 || t1 == 0) {
                        mbar_wait(bar(PFULL), uint32_t(p_phase) & 1u);
                        p_phase++;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        TC_TRACE(TR_G1_ISSUE, it);
                        const uint32_t st = base + SmemLayout::stage0 + uint32_t(s) * SmemLayout::stage_bytes;
                        const uint64_t qh = make_desc(st + SmemLayout::q_hi, 16, 1024), ql = make_desc(st + SmemLayout::q_lo, 16, 1024);
                        const uint32_t d = tmem + uint32_t(TM_S + sb * OTH);
#pragma unroll
                        for (int tt = 0; tt < NSPLIT; tt++) {
                            const int term = NSPLIT == 3 ? (tt == 0 ? 1 : (tt == 1 ? 2 : 0)) : 0;   // hi*lo, lo*hi, hi*hi
                            const uint32_t pa = tmem + uint32_t((term == 2) ? TM_P_LO : TM_P_HI);
                            const uint64_t qa = (term == 1) ? ql : qh;
#pragma unroll
                            for (int kk = 0; kk < KC / 8; kk++) {
                                if (tt == 0 && kk == 0) umma_tf32_ts<false>(d, pa, qa, idesc1);
                                else umma_tf32_ts<true>(d, pa + uint32_t(kk * 8), qa + uint64_t(kk * 2), idesc1);
                            }
                        }
                        umma_commit(bar(SFULL0 + sb));
                    }
                    __syncwarp();
                    if (++t1 == T) t1 = 0;
                }
            }
        } else {
            // GEMM2(it) needs R(it) and Q^T(it) published (RFULL: X, S and Q of the tile consumed)
            // and, at the start of a chain, the previous chain's OUT read (OUTEMPTY).  Corrections go to their own
            // accumulator OUT2 (their rounding is relative to 2^-11 |OUT|), only the eight hi*hi MMAs per tile extend the
            // long chain in OUT; the epilogue adds OUT + OUT2.
            uint64_t dqt_hi[2], dqt_lo[2];
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const uint32_t qt = base + SmemLayout::qt0 + uint32_t(b) * SmemLayout::qt_bytes;
                dqt_hi[b] = make_desc(qt + SmemLayout::qt_hi, 16, 1024);
                dqt_lo[b] = make_desc(qt + SmemLayout::qt_lo, 16, 1024);
            }
            const bool main_issuer = warp == W_G2;       // hi*hi into OUT; the other warp: hi*lo + lo*hi into OUT2
            const int n_g2 = (NSPLIT == 1 && !main_issuer) ? 0 : n_it;   // 1xTF32: no correction terms, no second issuer
            auto issue_g2 = [&](int it, bool first, bool last) {
                const int rb = it & 1;
                const uint32_t r_hi = tmem + uint32_t(TM_R + rb * 2 * OTH), r_lo = r_hi + OTH;
                const uint64_t qh = rb ? dqt_hi[1] : dqt_hi[0], ql = rb ? dqt_lo[1] : dqt_lo[0];
#pragma unroll
                for (int term = 0; term < NSPLIT; term++) {
                    if ((term == 0) != main_issuer) continue;
                    const uint32_t d = tmem + uint32_t(term == 0 ? TM_OUT : TM_OUT2);
                    const uint32_t ra = (term == 2) ? r_lo : r_hi;
                    const uint64_t qa = (term == 1) ? ql : qh;
#pragma unroll
                    for (int kk = 0; kk < OTH / 8; kk++) {
                        // K-step kk: 8 TMEM columns of R; Q^T K-block kk / 4 (32 rows x 128 B = 256 x 16 B)
                        const uint64_t bd = qa + uint64_t((kk / 4) * (KC * 128 / 16) + (kk % 4) * 2);
                        if (term <= 1 && kk == 0 && first) umma_tf32_ts<false>(d, ra, bd, idesc2);
                        else umma_tf32_ts<true>(d, ra + uint32_t(kk * 8), bd, idesc2);
                    }
                }
                umma_commit(bar(G2DONE0 + rb));
                if (last) umma_commit(bar(OUTFULL));
            };
            int t2 = t_first, c2 = c_first, chains_done = 0;
            for (int it = 0; it < n_g2; it++) {
                const bool first = first_of_chain(it, c2), last = last_of_chain(it, t2, c2);
                mbar_wait(bar(RFULL0 + (it & 1)), uint32_t(it >> 1) & 1u);
                if (first && chains_done > 0) mbar_wait(bar(OUTEMPTY), uint32_t(chains_done - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    // X (epilogue), Q (GEMM1: complete before S was read; transposition) of this stage are no longer needed.
                    // (Releasing the stage earlier -- from the epilogue warps when they publish R, or from the GEMM1 warp --
                    //  was measured and gave nothing: the load phase grows, the GEMM1 latency grows by as much.)
                    if (main_issuer) {
                        mbar_arrive(bar(EMPTY0 + it % NSTAGE));
                        TC_TRACE(TR_G2_ISSUE, it);
                    }
                    if (first) issue_g2(it, true, last); else issue_g2(it, false, last);
                }
                __syncwarp();
                if (last) chains_done++;
                if (++t2 == T) { t2 = 0; c2 = 0; }
                else if (++c2 == chain) c2 = 0;
            }
        }
    } else {
        // ================================ epilogue (warps 0 .. 15) ================================
        const int q = warp & 3;                   // TMEM lane quadrant (hardware: warp id % 4)
        const int cchunk = warp >> 2;             // which 16-column chunk of the 64-wide tile this warp converts
        const int i = q * 32 + lane;              // own row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + (uint32_t(q * 32) << 16);
        const bool out_warp = cchunk < KC / 16;   // two warps per quadrant own the two 16-column halves of OUT
        // this thread's 8 + 8 columns of P (tf32 hi / lo of the own-side factor row) -> tensor memory
        auto park_p = [&](int64_t own_tile) {
            const int64_t own_idx = own_tile * OWN + i;
            float ph[16];
#pragma unroll
            for (int e = 0; e < 16; e++) ph[e] = 0.0f;
            if (own_idx < prm.own_n) {
                const float4* src_h = reinterpret_cast<const float4*>(prm.p_hi + own_idx * KC + cchunk * 8);
                const float4* src_l = reinterpret_cast<const float4*>(prm.p_lo + own_idx * KC + cchunk * 8);
                const float4 a0 = src_h[0], a1 = src_h[1];
                ph[0] = a0.x; ph[1] = a0.y; ph[2] = a0.z; ph[3] = a0.w; ph[4] = a1.x; ph[5] = a1.y; ph[6] = a1.z; ph[7] = a1.w;
                if (NSPLIT == 3) {
                    const float4 b0 = src_l[0], b1 = src_l[1];
                    ph[8] = b0.x; ph[9] = b0.y; ph[10] = b0.z; ph[11] = b0.w; ph[12] = b1.x; ph[13] = b1.y; ph[14] = b1.z; ph[15] = b1.w;
                }
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(lane_addr + uint32_t(TM_P_HI + cchunk * 8)), "r"(__float_as_uint(ph[0])),
                           "r"(__float_as_uint(ph[1])), "r"(__float_as_uint(ph[2])), "r"(__float_as_uint(ph[3])),
                           "r"(__float_as_uint(ph[4])), "r"(__float_as_uint(ph[5])), "r"(__float_as_uint(ph[6])),
                           "r"(__float_as_uint(ph[7])) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(lane_addr + uint32_t(TM_P_LO + cchunk * 8)), "r"(__float_as_uint(ph[8])),
                           "r"(__float_as_uint(ph[9])), "r"(__float_as_uint(ph[10])), "r"(__float_as_uint(ph[11])),
                           "r"(__float_as_uint(ph[12])), "r"(__float_as_uint(ph[13])), "r"(__float_as_uint(ph[14])),
                           "r"(__float_as_uint(ph[15])) : "memory");
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(PFULL));
        };
        if (RESID && n_it > 0) park_p(o_first);
        double sq = 0.0;
        float acc[16];                            // sum of the chains of the current own tile (OUT + OUT2), out_warp only
#pragma unroll
        for (int e = 0; e < 16; e++) acc[e] = 0.0f;
        int chains_seen = 0;                      // chains whose OUT this (out) warp has read
        bool pend = false, pend_own_ends = false; // a finished chain whose OUT has not been read yet
        int64_t pend_own_tile = 0;
        int64_t own_tile = o_first;
        int t = t_first, cpos = c_first;
        const int c = cchunk;
        // loop-invariant byte offsets of this thread's 16 X elements inside a stage (swizzle XOR folded in)
        uint32_t xoff[MODE == 0 ? 4 : 8];
        if (MODE == 0) {
#pragma unroll
            for (int gq = 0; gq < 4; gq++) {
                const int j = c * 16 + gq * 4;
                const int blk = j >> 5, ch = (j & 31) >> 2;
                xoff[gq] = uint32_t(SmemLayout::x + blk * (OWN * 128) + i * 128 + ((ch ^ (i & 7)) << 4));
            }
        } else {
            const int blk = i >> 5, ch = (i & 31) >> 2, w = i & 3;
#pragma unroll
            for (int m = 0; m < 8; m++)       // m = j & 7 for the other index j = 16c + 4gq + e
                xoff[m] = uint32_t(SmemLayout::x + blk * (OTH * 128) + c * 16 * 128 + ((ch ^ m) << 4) + w * 4);
        }
        // this warp's share of the transposition Q tile (64 other rows x 32 components, 128-byte rows, 16-byte chunk cc of
        // row j at position cc ^ (j & 7)) -> Q^T as two K-blocks of 32 component rows x 32 other columns in the same
        // swizzle: K-block qb, chunk qc of every row of the block (lane = other row inside the block): one conflict-free
        // 128-bit read and four conflict-free 32-bit writes per part
        const int qb = warp >> 3, qc = warp & 7;
        const uint32_t qsrc = uint32_t((qb * 32 + lane) * 128 + ((qc ^ (lane & 7)) << 4));
        uint32_t qdst[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int kidx = 4 * qc + e;
            qdst[e] = uint32_t(SmemLayout::qt0 + qb * (KC * 128) + kidx * 128 + (((lane >> 2) ^ (kidx & 7)) << 4) + (lane & 3) * 4);
        }
        for (int it = 0; it < n_it; it++) {
            const int s = it % NSTAGE, sb = it & 1;
            const int64_t own0 = own_tile * OWN, oth0 = int64_t(t) * OTH;
            const bool own_ok = own0 + i < prm.own_n;
            mbar_wait(bar(FULL0 + s), uint32_t(it / NSTAGE) & 1u);
            if (warp == 0 && lane == 0) TC_TRACE(TR_FULL_SEEN_EPI, it);
            const unsigned char* xs = gen + SmemLayout::stage0 + s * SmemLayout::stage_bytes;
            // this thread's 4 + 4 elements of the Q tile and 16 elements of the X tile (other indices j = 16c .. 16c + 15),
            // read while GEMM1 still runs
            const float4 qv_hi = *reinterpret_cast<const float4*>(xs + SmemLayout::q_hi + qsrc);
            float4 qv_lo = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NSPLIT == 3) qv_lo = *reinterpret_cast<const float4*>(xs + SmemLayout::q_lo + qsrc);
            float xv[16];
            if (MODE == 0) {
#pragma unroll
                for (int gq = 0; gq < 4; gq++) {
                    const float4 t4 = *reinterpret_cast<const float4*>(xs + xoff[gq]);
                    xv[gq * 4] = t4.x; xv[gq * 4 + 1] = t4.y; xv[gq * 4 + 2] = t4.z; xv[gq * 4 + 3] = t4.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) xv[e] = *reinterpret_cast<const float*>(xs + xoff[e & 7] + e * 128);
            }
            float rr[16];
            if (RESID) {
                mbar_wait(bar(SFULL0 + sb), uint32_t(it >> 1) & 1u);
                tc_fence_after();
                if (warp == 0 && lane == 0) TC_TRACE(TR_S_SEEN, it);
                float sv[16];
                tmem_ld16(lane_addr + uint32_t(TM_S + sb * OTH + c * 16), sv);
                if (!LOGIT) {
                    // Linear link: no bounds masks.  Out-of-range other rows of Q and out-of-range own rows of P are zero
                    // (TMA zero fill / park_p), so S is exactly 0 there, and so is X (TMA zero fill): R = 0 - 0.
#pragma unroll
                    for (int e = 0; e < 16; e++) rr[e] = sv[e] - xv[e];
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e++) rr[e] = 1.0f / (1.0f + __expf(-sv[e])) - xv[e];
                    // sigmoid(0) = 0.5 outside the matrix: the last row / column tiles are masked
                    const bool interior = (own0 + OWN <= prm.own_n) && (oth0 + OTH <= prm.oth_n);
                    if (!interior) {
                        const int64_t oth_rem = prm.oth_n - oth0;
                        const int oth_left = oth_rem < OTH ? int(oth_rem) : OTH;
#pragma unroll
                        for (int e = 0; e < 16; e++)
                            if (!own_ok || c * 16 + e >= oth_left) rr[e] = 0.0f;
                    }
                }
                if (prm.sq_part != nullptr) {
                    float sq_tile = 0.0f;
#pragma unroll
                    for (int e = 0; e < 16; e++) sq_tile = fmaf(rr[e], rr[e], sq_tile);
                    sq += double(sq_tile);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; e++) rr[e] = xv[e];           // TMA zero-fills out-of-range elements
            }
            // tf32 split by truncation: hi = r with the 13 low mantissa bits cleared (exact tf32), lo = r - hi (exact
            // in fp32; the tensor core reads its top 19 bits) -> hi*hi + hi*lo + lo*hi carries ~2^-20 relative error
            float hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; e++) {
                hi[e] = NSPLIT == 3 ? __uint_as_float(__float_as_uint(rr[e]) & 0xffffe000u) : rr[e];
                lo[e] = rr[e] - hi[e];
            }
            // R buffer it & 1 in tensor memory: GEMM2 of tile it - 2 must have drained it
            mbar_wait(bar(G2DONE0 + sb), (uint32_t(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t r_hi = lane_addr + uint32_t(TM_R + sb * 2 * OTH + c * 16);
            tmem_st16(r_hi, hi);
            if (NSPLIT == 3) tmem_st16(r_hi + OTH, lo);
            {
                // Q^T slot it & 1 (free for the same reason as the R buffer)
                unsigned char* qd = gen + sb * SmemLayout::qt_bytes;
                const float h4[4] = {qv_hi.x, qv_hi.y, qv_hi.z, qv_hi.w}, l4[4] = {qv_lo.x, qv_lo.y, qv_lo.z, qv_lo.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    *reinterpret_cast<float*>(qd + SmemLayout::qt_hi + qdst[e]) = h4[e];
                    if (NSPLIT == 3) *reinterpret_cast<float*>(qd + SmemLayout::qt_lo + qdst[e]) = l4[e];
                }
                fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(RFULL0 + sb));
                if (warp == 0) TC_TRACE(TR_R_DONE, it);
            }
            const bool last = last_of_chain(it, t, cpos);
            const bool own_ends = (it == n_it - 1) || (t == T - 1);
            // the next tile belongs to another own tile: its P may go to tensor memory now (every GEMM1 that read the
            // old P has completed: this warp has just consumed the S of the last one)
            if (RESID && own_ends && it + 1 < n_it) park_p(own_tile + 1);
            // ---- end of a chain: OUT (+ OUT2) from tensor memory into the register accumulators.  Deferred by one tile:
            // reading it right away would park half of the epilogue warps until the chain's last GEMM2 has drained
            // (~2500 clk per chain, measured); one tile later OUTFULL has long completed and only GEMM2 of the new chain's
            // first tile waits, for the few hundred cycles of the read itself.
            auto flush_chain = [&](int64_t f_own_tile, bool f_own_ends) {
                mbar_wait(bar(OUTFULL), uint32_t(chains_seen) & 1u);
                tc_fence_after();
                float o[16];
                tmem_ld16(lane_addr + uint32_t(TM_OUT + cchunk * 16), o);
#pragma unroll
                for (int e = 0; e < 16; e++) acc[e] += o[e];
                if (NSPLIT == 3) {
                    tmem_ld16(lane_addr + uint32_t(TM_OUT2 + cchunk * 16), o);
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[e] += o[e];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(OUTEMPTY));
                chains_seen++;
                if (f_own_ends) {
                    // ---- one partial per (CTA, own tile)
                    const int64_t entry = int64_t(blockIdx.x) - first_cta(f_own_tile * T, prm.n_tiles, gridDim.x);
                    if (f_own_tile * OWN + i < prm.own_n) {
                        float4* dst = reinterpret_cast<float4*>(
                            prm.part + ((f_own_tile * prm.max_entries + entry) * OWN + i) * KC + cchunk * 16);
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            dst[e] = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
                    }
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[e] = 0.0f;
                }
            };
            if (out_warp) {
                if (pend) flush_chain(pend_own_tile, pend_own_ends);
                pend = false;
                if (last) {
                    if (it == n_it - 1) flush_chain(own_tile, true);
                    else { pend = true; pend_own_tile = own_tile; pend_own_ends = own_ends; }
                }
            }
            if (++t == T) { t = 0; cpos = 0; own_tile++; }
            else if (++cpos == chain) cpos = 0;
        }
        if (RESID && prm.sq_part != nullptr) {
            sq = warp_sum(sq);
            if (lane == 0) atomicAdd(sq_slot, sq);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(TR_CTA, 2);
    if (prm.trace != nullptr && threadIdx.x == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.trace[TR_NEVENTS * TRACE_TILES + 2 * blockIdx.x + 1] = gt;
    }
    if (RESID && prm.sq_part != nullptr && threadIdx.x == 0) prm.sq_part[blockIdx.x] = *sq_slot;
    if (warp == W_TMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(TMEM_COLS)) : "memory");
    }
}

// out[r][:] = sum over the entries of own tile r / OWN of part[tile][entry][r % OWN][:]   (fixed order: deterministic)
__global__ void tc_reduce_kernel(int64_t own_n, int64_t T, int64_t G, int64_t n_cta, int max_entries,
                                 const float* __restrict__ part, float* __restrict__ out) {
    const int64_t e4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;     // one float4 of the output per thread
    if (e4 >= own_n * (KC / 4)) return;
    const int64_t r = e4 / (KC / 4), c4 = e4 % (KC / 4);
    const int64_t tile = r / OWN;
    const int n = int(first_cta(tile * T + T - 1, G, n_cta) - first_cta(tile * T, G, n_cta)) + 1;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < n; e++) {
        const float4 v = *reinterpret_cast<const float4*>(part + ((tile * max_entries + e) * OWN + (r % OWN)) * KC + c4 * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * KC + c4 * 4) = s;
}

// tf32 hi / lo parts of a factor F (rows x 32), in the original layout and transposed (32 x ldt)
__global__ void split_tf32_kernel(int64_t rows, int64_t ldt, const float* __restrict__ x, float* __restrict__ hi,
                                  float* __restrict__ lo, float* __restrict__ hi_t, float* __restrict__ lo_t) {
    __shared__ float th[32][33], tl[32][33];
    const int64_t r0 = int64_t(blockIdx.x) * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        int64_t r = r0 + rr;
        float h = 0.f, l = 0.f;
        if (r < rows) {
            float v = x[r * KC + threadIdx.x];
            h = tf32_rna(v);
            l = tf32_rna(v - h);
            hi[r * KC + threadIdx.x] = h;
            lo[r * KC + threadIdx.x] = l;
        }
        th[rr][threadIdx.x] = h;
        tl[rr][threadIdx.x] = l;
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += blockDim.y) {
        int64_t r = r0 + threadIdx.x;
        if (r < ldt) {
            hi_t[int64_t(c) * ldt + r] = th[threadIdx.x][c];
            lo_t[int64_t(c) * ldt + r] = tl[threadIdx.x][c];
        }
    }
}


// tf32 parts of one factor in HBM: hi / lo in the factor's own layout (rows x 32) and transposed (32 x ldt)
struct FactorParts {
    const float *hi, *lo, *hi_t, *lo_t;
    int64_t rows, ldt;
};

// generic 2-D map: `cols` contiguous, box = (32, box_rows)
CUtensorMap factor_map(const float* p, int64_t rows, int box_rows) { return make_map(p, rows, KC, KC, box_rows); }
CUtensorMap factor_t_map(const float* p, int64_t rows, int64_t ldt) { return make_map(p, KC, rows, ldt, KC); }

template <int MODE, bool RESID, int NSPLIT, bool LOGIT>
void launch_tc_impl(pycmf_ctx* ctx, int64_t own_n, int64_t oth_n, const FactorParts& P, const FactorParts& Q,
               const float* X, int64_t x_rows, int64_t x_cols, int64_t ldx, int link, float* out, double* sq) {
    const int64_t own_tiles = ceil_div(own_n, OWN), T = ceil_div(oth_n, OTH), G = own_tiles * T;
    // one persistent CTA per SM; options for tests: tc_ctas caps the CTA count, tc_max_splits = s keeps the old meaning
    // "at most s CTAs per own tile" (s == 1 also lifts the chain cap: one long accumulation chain per own tile)
    int64_t n_cta = std::min<int64_t>(ctx->num_sms, G);
    if (ctx->tc_ctas > 0) n_cta = std::min<int64_t>(n_cta, ctx->tc_ctas);
    if (ctx->tc_max_splits > 0) n_cta = std::min<int64_t>(n_cta, own_tiles * ctx->tc_max_splits);
    // TMEM accumulation is fp32 without round-to-nearest: bound the chain length per accumulator (measured: 65 tiles
    // of random-sign data -> 3e-5 relative, 8 tiles -> 7e-7): at most 16 tiles, the chains are added in fp32 registers
    int chain = ctx->tc_chain > 0 ? ctx->tc_chain : 16;
    if (ctx->tc_max_splits == 1) chain = 1 << 30;
    int max_entries = 1;
    for (int64_t o = 0; o < own_tiles; o++)
        max_entries = std::max<int>(max_entries, int(first_cta(o * T + T - 1, G, n_cta) - first_cta(o * T, G, n_cta)) + 1);
    CUtensorMap tm_q_hi = factor_map(Q.hi, Q.rows, OTH);
    CUtensorMap tm_q_lo = factor_map(Q.lo, Q.rows, OTH);
    CUtensorMap tm_qt_hi = factor_t_map(Q.hi_t, Q.rows, Q.ldt);
    CUtensorMap tm_qt_lo = factor_t_map(Q.lo_t, Q.rows, Q.ldt);
    // X rows are in general not 256-byte aligned (row pitch 20000 B on C2): with 256-byte L2 promotion the RIGHT pass, whose
    // neighbouring column blocks are read by other CTAs at other times, pulled 34 % more than X from HBM (ncu, r01)
    const int promo = ctx->tc_x_promotion;
    CUtensorMap tm_x = make_map(X, x_rows, x_cols, ldx, MODE == 0 ? OWN : OTH,
                                promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                             : (promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                            : (promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                                          : CU_TENSOR_MAP_L2_PROMOTION_L2_128B)));
    Params prm;
    prm.own_n = own_n;
    prm.oth_n = oth_n;
    prm.n_oth_tiles = T;
    prm.n_tiles = G;
    prm.chain = chain;
    prm.link = link;
    prm.p_hi = P.hi;
    prm.p_lo = P.lo;
    prm.x = X;
    prm.ldx = ldx;
    prm.x_rows = x_rows;
    prm.x_cols = x_cols;
    prm.pf_mode = ctx->tc_prefetch;
    prm.part = static_cast<float*>(scratch(ctx, 0, size_t(own_tiles) * max_entries * OWN * KC * sizeof(float)));
    prm.max_entries = max_entries;
    prm.sq_part = nullptr;
    prm.trace = ctx->tc_trace ? static_cast<long long*>(scratch(ctx, 2, sizeof(long long) * (TR_NEVENTS * TRACE_TILES + 2 * 160))) : nullptr;
    if (RESID && sq != nullptr) prm.sq_part = static_cast<double*>(scratch(ctx, 1, size_t(n_cta) * sizeof(double)));
    auto kern = tc_pass_kernel<MODE, RESID, NSPLIT, LOGIT>;
    const size_t smem = SmemLayout::total + 1024;
    PYCMF_CHECK(smem <= size_t(ctx->max_smem_optin), "tc pass: shared memory budget exceeded");
    PYCMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    {
        Timed timer(ctx, RESID ? (MODE == 0 ? "tc_resid_left" : "tc_resid_right") : (MODE == 0 ? "tc_xv" : "tc_xtu"));
        kern<<<(unsigned)n_cta, NTHREADS, smem, ctx->stream>>>(tm_q_hi, tm_q_lo, tm_qt_hi, tm_qt_lo, tm_x, prm);
        PYCMF_LAUNCH_CHECK(ctx);
    }
    tc_reduce_kernel<<<(unsigned)ceil_div(own_n * (KC / 4), 256), 256, 0, ctx->stream>>>(own_n, T, G, n_cta, max_entries,
                                                                                       prm.part, out);
    PYCMF_LAUNCH_CHECK(ctx);
    if (prm.sq_part != nullptr) final_sum(ctx, int(n_cta), prm.sq_part, 1.0, sq, true);
}

template <int MODE, bool RESID, int NSPLIT>
void launch_tc(pycmf_ctx* ctx, int64_t own_n, int64_t oth_n, const FactorParts& P, const FactorParts& Q,
               const float* X, int64_t x_rows, int64_t x_cols, int64_t ldx, int link, float* out, double* sq) {
    if (RESID && link == PYCMF_LOGIT)
        launch_tc_impl<MODE, RESID, NSPLIT, RESID>(ctx, own_n, oth_n, P, Q, X, x_rows, x_cols, ldx, link, out, sq);
    else
        launch_tc_impl<MODE, RESID, NSPLIT, false>(ctx, own_n, oth_n, P, Q, X, x_rows, x_cols, ldx, link, out, sq);
}

size_t parts_floats(int64_t rows) { return size_t(2) * rows * KC + size_t(2) * KC * ((rows + 3) & ~int64_t(3)); }

// writes the four tf32 part arrays of F into `buf` (parts_floats(rows) floats)
FactorParts split_factor(pycmf_ctx* ctx, int64_t rows, const float* F, float* buf) {
    const int64_t ldt = (rows + 3) & ~int64_t(3);
    float *hi = buf, *lo = hi + rows * KC, *hi_t = lo + rows * KC, *lo_t = hi_t + KC * ldt;
    split_tf32_kernel<<<(unsigned)ceil_div(ldt, 32), dim3(32, 8), 0, ctx->stream>>>(rows, ldt, F, hi, lo, hi_t, lo_t);
    PYCMF_LAUNCH_CHECK(ctx);
    return FactorParts{hi, lo, hi_t, lo_t, rows, ldt};
}

}  // namespace

int tc_trace_words() { return TR_NEVENTS * TRACE_TILES + 2 * 160; }

bool tc_dense_eligible(pycmf_ctx* ctx, int64_t ra, int64_t rb, int64_t k, const float* X, int64_t ldx, bool trans_t) {
    if (ctx->dense_path == 0 || trans_t || X == nullptr) return false;
    if (k != KC) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0 || (ldx % 4) != 0) return false;
    if (ra * rb < (int64_t(1) << 16)) return false;      // tiny problems: the generic kernel has less setup
    return ra >= 1 && rb >= 1 && ra < (int64_t(1) << 31) && rb < (int64_t(1) << 31);
}

// R = f(A B^T) - X (RESID) : outL = R B and / or outR = R^T A, *sq += sum R^2.   A: ra x 32, B: rb x 32, X: ra x rb.
void tc_resid_pass(pycmf_ctx* ctx, int64_t ra, int64_t rb, const float* A, const float* B, const float* X, int64_t ldx,
                   int link, float* outL, float* outR, double* sq) {
    PYCMF_CHECK(outL != nullptr || outR != nullptr, "tc_resid_pass: objective-only passes use the generic kernel");
    const bool three = ctx->dense_path != 2;
    float* buf = static_cast<float*>(scratch(ctx, 3, sizeof(float) * (parts_floats(ra) + parts_floats(rb) + 64)));
    FactorParts Ap = split_factor(ctx, ra, A, buf);
    FactorParts Bp = split_factor(ctx, rb, B, buf + ((parts_floats(ra) + 31) & ~size_t(31)));
    if (outL != nullptr) {
        if (three) launch_tc<0, true, 3>(ctx, ra, rb, Ap, Bp, X, ra, rb, ldx, link, outL, sq);
        else launch_tc<0, true, 1>(ctx, ra, rb, Ap, Bp, X, ra, rb, ldx, link, outL, sq);
    }
    if (outR != nullptr) {
        double* s2 = outL == nullptr ? sq : nullptr;
        if (three) launch_tc<1, true, 3>(ctx, rb, ra, Bp, Ap, X, ra, rb, ldx, link, outR, s2);
        else launch_tc<1, true, 1>(ctx, rb, ra, Bp, Ap, X, ra, rb, ldx, link, outR, s2);
    }
}

// out = X Q (trans == false: X is rows x cols, Q is cols x 32, out rows x 32)
//       X^T Q (trans == true : Q is rows x 32, out cols x 32)
void tc_xmul(pycmf_ctx* ctx, bool trans, int64_t rows, int64_t cols, const float* X, int64_t ldx, const float* Q,
             float* out) {
    const bool three = ctx->dense_path != 2;
    const int64_t qn = trans ? rows : cols;
    float* buf = static_cast<float*>(scratch(ctx, 3, sizeof(float) * (parts_floats(qn) + 64)));
    FactorParts Qp = split_factor(ctx, qn, Q, buf);
    if (!trans) {
        if (three) launch_tc<0, false, 3>(ctx, rows, cols, Qp, Qp, X, rows, cols, ldx, 0, out, nullptr);
        else launch_tc<0, false, 1>(ctx, rows, cols, Qp, Qp, X, rows, cols, ldx, 0, out, nullptr);
    } else {
        if (three) launch_tc<1, false, 3>(ctx, cols, rows, Qp, Qp, X, rows, cols, ldx, 0, out, nullptr);
        else launch_tc<1, false, 1>(ctx, cols, rows, Qp, Qp, X, rows, cols, ldx, 0, out, nullptr);
    }
}

}  // namespace pycmf

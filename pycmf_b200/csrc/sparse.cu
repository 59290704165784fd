// CSR kernels: SpMM (C = alpha * S * B + beta * C) and SDDMM-style reductions over the nonzeros.
// Work unit = a CHUNK of at most 512 nonzeros of one row, one warp per chunk (rows are split so that a hot row /
// column of a tf-idf-like matrix -- 10^5 nonzeros in one CSC row is normal -- does not serialise on one warp);
// chunk offsets come from a per-call count + exclusive scan, the kernel is grid-stride over the chunks so no host
// synchronisation is needed.  The nonzeros of a chunk are fetched coalesced (lane-strided) and broadcast by shuffle;
// sub-warp groups of G lanes walk different nonzeros when k < 32 so no lane idles.
#include <cub/device/device_scan.cuh>

#include <type_traits>

#include "common.cuh"

namespace pycmf {
namespace {

constexpr int MAXT = 8;  // columns per lane: k <= 32 * MAXT = 256

constexpr int CHUNK = 512;

// chunks[r] = max(1, ceil(len_r / CHUNK)); rows that will be combined by atomics are pre-scaled by beta here
template <typename T>
__global__ void spmm_count_kernel(int64_t rows, const int32_t* __restrict__ rowptr, int* __restrict__ chunks,
                                  T* __restrict__ C, int64_t ldc, int k, T beta) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r > rows) return;
    if (r == rows) { chunks[r] = 0; return; }
    const int len = rowptr[r + 1] - rowptr[r];
    const int nc = len <= CHUNK ? 1 : (len + CHUNK - 1) / CHUNK;
    chunks[r] = nc;
    if (nc > 1)
        for (int c = 0; c < k; c++) C[r * ldc + c] = beta != T(0) ? beta * C[r * ldc + c] : T(0);
}

template <typename T, int G>
__global__ void __launch_bounds__(256)
spmm_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
            const T* __restrict__ vals, const T* __restrict__ B, int64_t ldb, int k,
            T* __restrict__ C, int64_t ldc, T alpha, T beta, const int* __restrict__ chunk_off) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = chunk_off[rows];
    constexpr int NG = 32 / G;            // nonzeros in flight per warp
    const int gl = lane % G, gid = lane / G;
    for (int64_t ch = warp0; ch < total; ch += nwarps) {
        // row of this chunk: last r with chunk_off[r] <= ch
        int64_t lo = 0, hi = rows;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (int64_t(chunk_off[mid]) <= ch) lo = mid; else hi = mid;
        }
        const int64_t row = lo;
        const int ci = int(ch - chunk_off[row]);
        const bool split = chunk_off[row + 1] - chunk_off[row] > 1;
        T acc[MAXT];
#pragma unroll
        for (int t = 0; t < MAXT; t++) acc[t] = T(0);
        const int start = rowptr[row] + ci * CHUNK;
        const int end = min(rowptr[row + 1], start + CHUNK);
        for (int base = start; base < end; base += 32) {
            int cnt = min(32, end - base);
            int c = 0; T v = T(0);
            if (lane < cnt) { c = colidx[base + lane]; v = vals[base + lane]; }
            for (int j = 0; j < cnt; j += NG) {
                int src = j + gid;
                int cj = __shfl_sync(0xffffffffu, c, src & 31);
                T vj = __shfl_sync(0xffffffffu, v, src & 31);
                if (src < cnt) {
                    const T* brow = B + int64_t(cj) * ldb;
#pragma unroll
                    for (int t = 0; t < MAXT; t++) {
                        int col = gl + t * G;
                        if (col < k) acc[t] = fma(vj, brow[col], acc[t]);
                    }
                }
            }
        }
        // reduce across the NG groups
#pragma unroll
        for (int t = 0; t < MAXT; t++) {
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
        }
        if (gid == 0) {
#pragma unroll
            for (int t = 0; t < MAXT; t++) {
                int col = gl + t * G;
                if (col < k) {
                    if (split) {
                        atomicAdd(&C[row * ldc + col], alpha * acc[t]);     // row pre-scaled by beta in the count pass
                    } else {
                        T prev = beta != T(0) ? C[row * ldc + col] : T(0);
                        C[row * ldc + col] = alpha * acc[t] + beta * prev;
                    }
                }
            }
        }
    }
}


// k = 32 * VEC: lane owns the VEC contiguous columns [lane * VEC, lane * VEC + VEC) of every factor row, so one
// nonzero is one vector load per lane (a full 128 .. 1024-byte row of B per warp).  A warp walks a CONTIGUOUS range
// of chunks (one binary search per warp instead of one per chunk: 18 dependent L2 round trips per ~100-nonzero row
// were most of the old kernel's time) and keeps UNROLL independent row gathers in flight.
template <typename T, int VEC> struct VecLoad;
template <> struct VecLoad<float, 1> { static __device__ __forceinline__ void ld(const float* p, float (&o)[1]) { o[0] = __ldg(p); } };
template <> struct VecLoad<float, 2> { static __device__ __forceinline__ void ld(const float* p, float (&o)[2]) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p)); o[0] = v.x; o[1] = v.y; } };
template <> struct VecLoad<float, 4> { static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; } };
template <> struct VecLoad<float, 8> { static __device__ __forceinline__ void ld(const float* p, float (&o)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; } };
template <> struct VecLoad<double, 1> { static __device__ __forceinline__ void ld(const double* p, double (&o)[1]) { o[0] = __ldg(p); } };
template <> struct VecLoad<double, 2> { static __device__ __forceinline__ void ld(const double* p, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p)); o[0] = v.x; o[1] = v.y; } };
template <> struct VecLoad<double, 4> { static __device__ __forceinline__ void ld(const double* p, double (&o)[4]) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; } };
template <> struct VecLoad<double, 8> { static __device__ __forceinline__ void ld(const double* p, double (&o)[8]) {
#pragma unroll
    for (int h = 0; h < 4; h++) { const double2 a = __ldg(reinterpret_cast<const double2*>(p) + h); o[2 * h] = a.x; o[2 * h + 1] = a.y; } } };

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
spmm_vec_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const T* __restrict__ vals, const T* __restrict__ B, int64_t ldb,
                T* __restrict__ C, int64_t ldc, T alpha, T beta, const int* __restrict__ chunk_off) {
    constexpr int UNROLL = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = chunk_off[rows];
    const int64_t ch0 = (warp * total) / nwarps, ch1 = ((warp + 1) * total) / nwarps;
    if (ch0 >= ch1) return;
    // row of the first chunk: last r with chunk_off[r] <= ch0; chunk_off[r] >= r (every row has a chunk) and
    // chunk_off[r] <= r + (total - rows), so r lies in [ch0 - (total - rows), ch0]
    int64_t lo = max(int64_t(0), ch0 - (total - rows)), hi = min(rows, ch0 + 1);
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (int64_t(chunk_off[mid]) <= ch0) lo = mid; else hi = mid;
    }
    int64_t row = lo;
    int row_first = chunk_off[row], row_next = chunk_off[row + 1];
    for (int64_t ch = ch0; ch < ch1; ch++) {
        while (ch >= row_next) { row++; row_first = row_next; row_next = chunk_off[row + 1]; }
        const int ci = int(ch - row_first);
        const bool split = row_next - row_first > 1;
        T acc[VEC];
#pragma unroll
        for (int t = 0; t < VEC; t++) acc[t] = T(0);
        const int start = rowptr[row] + ci * CHUNK;
        const int end = min(rowptr[row + 1], start + CHUNK);
        for (int base = start; base < end; base += 32) {
            const int cnt = min(32, end - base);
            // lanes past the end carry a valid column (the chunk's first) and a zero value: no branches in the gather loop
            int c = colidx[base + (lane < cnt ? lane : 0)];
            T v = lane < cnt ? vals[base + lane] : T(0);
            for (int j = 0; j < cnt; j += UNROLL) {
                T bv[UNROLL][VEC], vj[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int cj = __shfl_sync(0xffffffffu, c, (j + u) & 31);
                    vj[u] = __shfl_sync(0xffffffffu, v, (j + u) & 31);
                    VecLoad<T, VEC>::ld(B + int64_t(cj) * ldb + lane * VEC, bv[u]);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
#pragma unroll
                    for (int t = 0; t < VEC; t++) acc[t] = fma(vj[u], bv[u][t], acc[t]);
            }
        }
        T* crow = C + row * ldc + lane * VEC;
#pragma unroll
        for (int t = 0; t < VEC; t++) {
            if (split) {
                atomicAdd(&crow[t], alpha * acc[t]);                 // row pre-scaled by beta in the count pass
            } else {
                const T prev = beta != T(0) ? crow[t] : T(0);
                crow[t] = alpha * acc[t] + beta * prev;
            }
        }
    }
}

// fp32, k = 32 / 64: a nonzero needs only LPN = k / 4 lanes for a 16-byte load each, so one load instruction gathers
// 32 / LPN factor rows (ncu on the C3 slice: the one-row-per-instruction kernel was issue-bound at ~15 warp
// instructions per nonzero, smsp issue active 52 - 64 %).  The sub-warp groups are summed by shuffles at the end of a chunk.
template <int LPN>
__global__ void __launch_bounds__(256)
spmm_grp_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const float* __restrict__ vals, const float* __restrict__ B, int64_t ldb,
                float* __restrict__ C, int64_t ldc, float alpha, float beta, const int* __restrict__ chunk_off) {
    constexpr int NG = 32 / LPN;          // nonzeros per load instruction
    constexpr int UNROLL = 4;
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPN, gid = lane / LPN;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = chunk_off[rows];
    const int64_t ch0 = (warp * total) / nwarps, ch1 = ((warp + 1) * total) / nwarps;
    if (ch0 >= ch1) return;
    int64_t lo = max(int64_t(0), ch0 - (total - rows)), hi = min(rows, ch0 + 1);
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (int64_t(chunk_off[mid]) <= ch0) lo = mid; else hi = mid;
    }
    int64_t row = lo;
    int row_first = chunk_off[row], row_next = chunk_off[row + 1];
    const float* Bl = B + gl * 4;
    for (int64_t ch = ch0; ch < ch1; ch++) {
        while (ch >= row_next) { row++; row_first = row_next; row_next = chunk_off[row + 1]; }
        const int ci = int(ch - row_first);
        const bool split = row_next - row_first > 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int start = rowptr[row] + ci * CHUNK;
        const int end = min(rowptr[row + 1], start + CHUNK);
        for (int base = start; base < end; base += 32) {
            const int cnt = min(32, end - base);
            // lanes past the end carry a valid column (the chunk's first) and a zero value: no branches in the gather loop
            const int c = colidx[base + (lane < cnt ? lane : 0)];
            const float v = lane < cnt ? vals[base + lane] : 0.0f;
            for (int j = 0; j < cnt; j += NG * UNROLL) {
                float4 bv[UNROLL];
                float vj[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int src = (j + u * NG + gid) & 31;        // wraps only onto lanes already consumed ...
                    const int cj = __shfl_sync(0xffffffffu, c, src);
                    const float vv = __shfl_sync(0xffffffffu, v, src);
                    vj[u] = (j + u * NG + gid) < 32 ? vv : 0.0f;     // ... whose value is masked here
                    bv[u] = __ldg(reinterpret_cast<const float4*>(Bl + int64_t(cj) * ldb));
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    acc.x = fmaf(vj[u], bv[u].x, acc.x); acc.y = fmaf(vj[u], bv[u].y, acc.y);
                    acc.z = fmaf(vj[u], bv[u].z, acc.z); acc.w = fmaf(vj[u], bv[u].w, acc.w);
                }
            }
        }
#pragma unroll
        for (int o = 16; o >= LPN; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (gid == 0) {
            float* crow = C + row * ldc + gl * 4;
            if (split) {
                atomicAdd(&crow[0], alpha * acc.x); atomicAdd(&crow[1], alpha * acc.y);
                atomicAdd(&crow[2], alpha * acc.z); atomicAdd(&crow[3], alpha * acc.w);
            } else {
                float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
                if (beta != 0.0f) prev = *reinterpret_cast<const float4*>(crow);
                *reinterpret_cast<float4*>(crow) = make_float4(alpha * acc.x + beta * prev.x, alpha * acc.y + beta * prev.y,
                                                               alpha * acc.z + beta * prev.z, alpha * acc.w + beta * prev.w);
            }
        }
    }
}


// ---- nonzero-balanced SpMM (fp32, k = 32 / 64 / 128) ------------------------------------------------------------------
// Every warp owns an EQUAL share of the nonzeros, [w * nnz / W, (w + 1) * nnz / W), found by one binary search in rowptr
// (merge-path style) -- no per-call chunk count / scan, and no tail: with chunk-count shares a warp that drew 512-nonzero
// chunks of a hot tf-idf column ran 50x longer than one that drew tail columns (ncu: 29 % achieved occupancy).
// The warp streams its share in flat 32-nonzero blocks (coalesced colidx / vals, the next block prefetched while the
// current one is gathered) and walks the row boundaries inside; a row that lies completely inside the share is stored
// (C = alpha * acc + beta * C), a row cut by a share boundary is combined by atomics -- such rows, and empty rows, are
// pre-scaled by beta in the prologue kernel, which finds them in O(1) per row from the same partition formula.
// What bounds the kernel is the L2 -> SM gather of one k * 4-byte factor row per nonzero (UNROLL independent 16-byte loads
// per lane in flight): see DESIGN section 4.
__device__ __forceinline__ int64_t nzb_share_begin(int64_t w, int64_t total, int64_t nwarps) {
    return (w * total) / nwarps;
}

__global__ void spmm_nzb_prologue_kernel(int64_t rows, const int32_t* __restrict__ rowptr, int64_t nwarps,
                                         float* __restrict__ C, int64_t ldc, int k, float beta) {
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int64_t base = rowptr[0];                        // a slice of a larger CSR keeps absolute offsets
    const int64_t total = rowptr[rows] - base;
    const int64_t a = rowptr[r] - base, b = rowptr[r + 1] - base;
    bool touch = a == b;                                   // empty row: nobody else writes it
    if (!touch && total > 0) {
        // is there a share boundary strictly inside (a, b)?  smallest w with begin(w) > a
        int64_t w = (a * nwarps) / total;
        while (w <= nwarps && nzb_share_begin(w, total, nwarps) <= a) w++;
        touch = w < nwarps && nzb_share_begin(w, total, nwarps) < b;
    }
    if (touch)
        for (int c = 0; c < k; c++) C[r * ldc + c] = beta != 0.0f ? beta * C[r * ldc + c] : 0.0f;
}

template <int LPN, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB)
spmm_nzb_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const float* __restrict__ vals, const float* __restrict__ B, int64_t ldb,
                float* __restrict__ C, int64_t ldc, float alpha, float beta) {
    constexpr int NG = 32 / LPN;          // nonzeros per load instruction
    constexpr int BATCH = NG * UNROLL;    // nonzeros per round trip
    const int lane = threadIdx.x & 31;
    const int gl = lane % LPN, gid = lane / LPN;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int base = rowptr[0];                           // a slice of a larger CSR keeps absolute offsets
    const int total = rowptr[rows] - base;                // nnz < 2^31 (int32 CSR): 32-bit positions keep registers down
    const int e0 = base + int(nzb_share_begin(warp, total, nwarps)), e1 = base + int(nzb_share_begin(warp + 1, total, nwarps));
    if (e0 >= e1) return;
    // row holding nonzero e0: last r with rowptr[r] <= e0
    int lo = 0, hi = int(rows);
    while (hi - lo > 1) {
        const int mid = int((unsigned(lo) + unsigned(hi)) >> 1);
        if (rowptr[mid] <= e0) lo = mid; else hi = mid;
    }
    int row = lo;
    int row_begin = rowptr[row], row_end = rowptr[row + 1];
    const float* Bl = B + gl * 4;
    const int last = base + total - 1;
    // current block [pos, pos + 32) of the flat nonzero stream; lanes past the share end read a valid address, value 0
    int pos = e0;
    int c = colidx[min(pos + lane, last)];
    float v = pos + lane < e1 ? vals[pos + lane] : 0.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    while (pos < e1) {
        const int npos = pos + 32;
        int cn = 0; float vn = 0.0f;
        if (npos < e1) {                                     // prefetch the next block
            cn = colidx[min(npos + lane, last)];
            vn = npos + lane < e1 ? vals[npos + lane] : 0.0f;
        }
        const int bend = min(32, e1 - pos);    // nonzeros of this block inside the share
        int j0 = 0;
        while (j0 < bend) {
            const int j1 = min(bend, row_end - pos);      // this row's part of the block: [j0, j1)
            for (int j = j0; j < j1; j += BATCH) {
                float4 bv[UNROLL];
                float vj[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int src = j + u * NG + gid;
                    const int s2 = src < j1 ? src : j0;          // past the row's part: a valid column, value masked
                    const int cj = __shfl_sync(0xffffffffu, c, s2);
                    const float vv = __shfl_sync(0xffffffffu, v, s2);
                    vj[u] = src < j1 ? vv : 0.0f;
                    bv[u] = __ldg(reinterpret_cast<const float4*>(Bl + int64_t(cj) * ldb));
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    acc.x = fmaf(vj[u], bv[u].x, acc.x); acc.y = fmaf(vj[u], bv[u].y, acc.y);
                    acc.z = fmaf(vj[u], bv[u].z, acc.z); acc.w = fmaf(vj[u], bv[u].w, acc.w);
                }
            }
            j0 = j1;
            if (pos + j1 == row_end || pos + j1 == e1) {
                // the row (or the share) ends here: combine the sub-warp groups and write
#pragma unroll
                for (int o = 16; o >= LPN; o >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                }
                const bool whole = row_begin >= e0 && row_end <= e1;
                if (gid == 0) {
                    float* crow = C + int64_t(row) * ldc + gl * 4;
                    if (!whole) {
                        atomicAdd(&crow[0], alpha * acc.x); atomicAdd(&crow[1], alpha * acc.y);
                        atomicAdd(&crow[2], alpha * acc.z); atomicAdd(&crow[3], alpha * acc.w);
                    } else {
                        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (beta != 0.0f) prev = *reinterpret_cast<const float4*>(crow);
                        *reinterpret_cast<float4*>(crow) = make_float4(alpha * acc.x + beta * prev.x, alpha * acc.y + beta * prev.y,
                                                                       alpha * acc.z + beta * prev.z, alpha * acc.w + beta * prev.w);
                    }
                }
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pos + j1 == row_end && pos + j1 < e1) {
                    // next non-empty row (empty rows were written by the prologue)
                    do { row++; row_begin = row_end; row_end = rowptr[row + 1]; } while (row_end == row_begin);
                }
            }
        }
        pos = npos; c = cn; v = vn;
    }
}


// ---- the same kernel on an instruction diet -----------------------------------------------------------------------------
// ncu on the C3 shard put spmm_nzb_kernel at 12.9 warp instructions per nonzero with the issue slots 63 - 66 % busy and only
// 57 % of the L2 throughput used (profiles/r02_ncu_c3_spmm_nzb_summary.txt): two shuffles, two selects and 64-bit address
// arithmetic per gathered row, four scalar FMAs per 16 bytes.  Here a warp stages each 32-nonzero block in shared memory as
// 16-byte records {row offset of the factor row (32-bit, elements), unused, value, value}: a sub-warp group fetches its
// nonzero with ONE broadcast LDS.128, forms the address with one IMAD.WIDE, and feeds the value pair straight into two packed
// FMAs (fma.rn.f32x2) per 16-byte factor segment -- 5 instructions per gathered row.  Masking (a row ending inside a batch)
// lives in a separate tail batch.  Needs rows(B) * ldb < 2^31.
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int LPN, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB)
spmm_nzb2_kernel(int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const float* __restrict__ vals, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int64_t ldc, float alpha, float beta) {
    constexpr int NG = 32 / LPN;          // nonzeros per load instruction
    constexpr int BATCH = NG * UNROLL;    // nonzeros per round trip
    __shared__ uint4 stage[8][32];        // per warp: one record per nonzero of the current block
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane % LPN, gid = lane / LPN;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int base = rowptr[0];
    const int total = rowptr[rows] - base;
    const int e0 = base + int(nzb_share_begin(warp, total, nwarps)), e1 = base + int(nzb_share_begin(warp + 1, total, nwarps));
    if (e0 >= e1) return;
    int lo = 0, hi = int(rows);
    while (hi - lo > 1) {
        const int mid = int((unsigned(lo) + unsigned(hi)) >> 1);
        if (rowptr[mid] <= e0) lo = mid; else hi = mid;
    }
    int row = lo;
    int row_begin = rowptr[row], row_end = rowptr[row + 1];
    const float* Bl = B + gl * 4;
    const int last = base + total - 1;
    uint4* st = stage[wib];
    int pos = e0;
    int c = colidx[min(pos + lane, last)];
    float v = pos + lane < e1 ? vals[pos + lane] : 0.0f;
    unsigned long long acc0 = 0ull, acc1 = 0ull;        // (x, y) and (z, w) of this lane's 16-byte output segment
    while (pos < e1) {
        const int npos = pos + 32;
        int cn = 0; float vn = 0.0f;
        if (npos < e1) {                                     // prefetch the next block
            cn = colidx[min(npos + lane, last)];
            vn = npos + lane < e1 ? vals[npos + lane] : 0.0f;
        }
        __syncwarp();                                        // everybody finished reading the previous block's records
        st[lane] = make_uint4(unsigned(c) * unsigned(ldb), 0u, __float_as_uint(v), __float_as_uint(v));
        __syncwarp();
        const int bend = min(32, e1 - pos);
        int j0 = 0;
        while (j0 < bend) {
            const int j1 = min(bend, row_end - pos);         // this row's part of the block: [j0, j1)
            int j = j0;
            for (; j + BATCH <= j1; j += BATCH) {            // full batches: no masking
                uint4 rec[UNROLL];
                ulonglong2 bv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) rec[u] = st[j + u * NG + gid];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) bv[u] = __ldg(reinterpret_cast<const ulonglong2*>(Bl + rec[u].x));
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const unsigned long long vv = (static_cast<unsigned long long>(rec[u].w) << 32) | rec[u].z;
                    ffma2(acc0, vv, bv[u].x);
                    ffma2(acc1, vv, bv[u].y);
                }
            }
            if (j < j1) {                                    // tail batch of this row's part: masked
                uint4 rec[UNROLL];
                ulonglong2 bv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int src = j + u * NG + gid;
                    rec[u] = st[src < j1 ? src : j0];
                    if (src >= j1) { rec[u].z = 0u; rec[u].w = 0u; }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) bv[u] = __ldg(reinterpret_cast<const ulonglong2*>(Bl + rec[u].x));
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const unsigned long long vv = (static_cast<unsigned long long>(rec[u].w) << 32) | rec[u].z;
                    ffma2(acc0, vv, bv[u].x);
                    ffma2(acc1, vv, bv[u].y);
                }
            }
            j0 = j1;
            if (pos + j1 == row_end || pos + j1 == e1) {
                float4 acc = make_float4(__uint_as_float(unsigned(acc0)), __uint_as_float(unsigned(acc0 >> 32)),
                                         __uint_as_float(unsigned(acc1)), __uint_as_float(unsigned(acc1 >> 32)));
#pragma unroll
                for (int o = 16; o >= LPN; o >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                }
                const bool whole = row_begin >= e0 && row_end <= e1;
                if (gid == 0) {
                    float* crow = C + int64_t(row) * ldc + gl * 4;
                    if (!whole) {
                        atomicAdd(&crow[0], alpha * acc.x); atomicAdd(&crow[1], alpha * acc.y);
                        atomicAdd(&crow[2], alpha * acc.z); atomicAdd(&crow[3], alpha * acc.w);
                    } else {
                        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (beta != 0.0f) prev = *reinterpret_cast<const float4*>(crow);
                        *reinterpret_cast<float4*>(crow) = make_float4(alpha * acc.x + beta * prev.x, alpha * acc.y + beta * prev.y,
                                                                       alpha * acc.z + beta * prev.z, alpha * acc.w + beta * prev.w);
                    }
                }
                acc0 = 0ull; acc1 = 0ull;
                if (pos + j1 == row_end && pos + j1 < e1) {
                    do { row++; row_begin = row_end; row_end = rowptr[row + 1]; } while (row_end == row_begin);
                }
            }
        }
        pos = npos; c = cn; v = vn;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
sddmm_reduce_kernel(int mode, int64_t rows, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                    const T* __restrict__ A, const T* __restrict__ B, int k, double* __restrict__ part) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    double s = 0.0;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const int start = rowptr[row], end = rowptr[row + 1];
        if (mode == 2) {
            for (int e = start + lane; e < end; e += 32) { double v = double(vals[e]); s += v * v; }
            continue;
        }
        T a[MAXT];
#pragma unroll
        for (int t = 0; t < MAXT; t++) { int col = lane + 32 * t; a[t] = col < k ? A[row * k + col] : T(0); }
        for (int e = start; e < end; e++) {
            const T* brow = B + int64_t(colidx[e]) * k;
            T d = T(0);
#pragma unroll
            for (int t = 0; t < MAXT; t++) { int col = lane + 32 * t; if (col < k) d = fma(a[t], brow[col], d); }
            d = warp_sum(d);
            if (lane == 0) {
                double x = double(vals[e]);
                if (mode == 0) {
                    s += x * double(d);
                } else {
                    double sg = 1.0 / (1.0 + exp(-double(d)));
                    s += (x - sg) * (x - sg) - sg * sg;
                }
            }
        }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

}  // namespace

template <typename T>
void spmm(pycmf_ctx* ctx, int64_t rows, const int32_t* rowptr, const int32_t* colidx, const T* vals,
          const T* B, int64_t ldb, int64_t k, T* C, int64_t ldc, T alpha, T beta, int64_t b_rows) {
    if (rows <= 0 || k <= 0) return;
    PYCMF_CHECK(k <= 32 * MAXT, "spmm: n_components > 256 is not supported");
    PYCMF_CHECK(rows < (int64_t(1) << 31) - 1, "spmm: too many rows");
    Timed timer(ctx, "spmm");
    if constexpr (std::is_same<T, float>::value) {
        // fp32, k = 32 / 64 / 128, 16-byte aligned rows: the nonzero-balanced kernel (no chunk count / scan per call)
        if ((k == 32 || k == 64 || k == 128) && ctx->spmm_path == 1 && (ldb % 4) == 0 && (ldc % 4) == 0 &&
            (reinterpret_cast<uintptr_t>(B) & 15) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
            const bool deep = ctx->spmm_unroll != 4;
            // one resident wave: 3 CTAs per SM at 80 registers (8 gathers in flight per lane), 4 at 56 (4 in flight)
            const int per_sm = ctx->spmm_blocks_per_sm > 0 ? ctx->spmm_blocks_per_sm : (deep ? 3 : 4);
            const unsigned nb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows * 4, 8), int64_t(per_sm) * ctx->num_sms));
            const int64_t nwarps = int64_t(nb) * 8;
            spmm_nzb_prologue_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, ctx->stream>>>(rows, rowptr, nwarps, C, ldc, int(k), beta);
            PYCMF_LAUNCH_CHECK(ctx);
            // lean variant (shared-memory staged nonzeros, packed FMAs): needs 32-bit element offsets into B
            const bool lean = ctx->spmm_lean != 0 && b_rows > 0 && b_rows * ldb < (int64_t(1) << 31);
#define LAUNCHN(L, U, M) spmm_nzb_kernel<L, U, M><<<nb, 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, ldb, C, ldc, alpha, beta)
#define LAUNCHL(L, U, M) spmm_nzb2_kernel<L, U, M><<<nb, 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, int(ldb), C, ldc, alpha, beta)
            if (lean) {
                if (k == 32) { if (deep) LAUNCHL(8, 8, 3); else LAUNCHL(8, 4, 4); }
                // (five / six resident CTAs per SM at 48 / 40 registers were measured slower: 0.45 / 0.53 ms against 0.41)
                else if (k == 64) { if (deep) LAUNCHL(16, 8, 3); else LAUNCHL(16, 4, 4); }
                else { if (deep) LAUNCHL(32, 8, 3); else LAUNCHL(32, 4, 4); }
            } else {
                if (k == 32) { if (deep) LAUNCHN(8, 8, 3); else LAUNCHN(8, 4, 4); }
                else if (k == 64) { if (deep) LAUNCHN(16, 8, 3); else LAUNCHN(16, 4, 4); }
                else { if (deep) LAUNCHN(32, 8, 3); else LAUNCHN(32, 4, 4); }
            }
#undef LAUNCHN
#undef LAUNCHL
            PYCMF_LAUNCH_CHECK(ctx);
            return;
        }
    }
    // chunk counts -> exclusive scan -> chunk offsets (rows + 1 ints), all on the stream
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (int*)nullptr, (int*)nullptr, int(rows + 1), ctx->stream);
    const size_t off_bytes = ((size_t(rows) + 1) * sizeof(int) + 255) & ~size_t(255);
    unsigned char* buf = static_cast<unsigned char*>(scratch(ctx, 2, 2 * off_bytes + cub_bytes + 256));
    int* counts = reinterpret_cast<int*>(buf);
    int* offsets = reinterpret_cast<int*>(buf + off_bytes);
    void* cub_tmp = buf + 2 * off_bytes;
    spmm_count_kernel<T><<<(unsigned)ceil_div(rows + 1, 256), 256, 0, ctx->stream>>>(rows, rowptr, counts, C, ldc, int(k), beta);
    PYCMF_LAUNCH_CHECK(ctx);
    PYCMF_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, offsets, int(rows + 1), ctx->stream));
    ctx->launches++;
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(rows * 32, 256), int64_t(32) * ctx->num_sms);
    // wide factors (k = 32, 64, 128, 256; 16-byte aligned rows): the vector kernel
    const bool vec_ok = (k == 32 || k == 64 || k == 128 || k == 256) && (ldb * sizeof(T)) % 16 == 0 &&
                        (reinterpret_cast<uintptr_t>(B) & 15) == 0 && ctx->spmm_path != 0;
    if (vec_ok) {
        const unsigned vb = (unsigned)std::min<int64_t>(ceil_div(rows * 32, 256 * 4), int64_t(16) * ctx->num_sms);
        if constexpr (std::is_same<T, float>::value) {
            if ((k == 32 || k == 64) && ctx->spmm_path == 3 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
                if (k == 32) spmm_grp_kernel<8><<<std::max(1u, vb), 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, ldb, C, ldc, alpha, beta, offsets);
                else spmm_grp_kernel<16><<<std::max(1u, vb), 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, ldb, C, ldc, alpha, beta, offsets);
                PYCMF_LAUNCH_CHECK(ctx);
                return;
            }
        }
#define LAUNCHV(V) spmm_vec_kernel<T, V><<<std::max(1u, vb), 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, ldb, C, ldc, alpha, beta, offsets)
        if (k == 32) LAUNCHV(1);
        else if (k == 64) LAUNCHV(2);
        else if (k == 128) LAUNCHV(4);
        else LAUNCHV(8);
#undef LAUNCHV
        PYCMF_LAUNCH_CHECK(ctx);
        return;
    }
#define LAUNCH(G) spmm_kernel<T, G><<<blocks, 256, 0, ctx->stream>>>(rows, rowptr, colidx, vals, B, ldb, int(k), C, ldc, alpha, beta, offsets)
    if (k <= 1) LAUNCH(1);
    else if (k <= 2) LAUNCH(2);
    else if (k <= 4) LAUNCH(4);
    else if (k <= 8) LAUNCH(8);
    else if (k <= 16) LAUNCH(16);
    else LAUNCH(32);
#undef LAUNCH
    PYCMF_LAUNCH_CHECK(ctx);
}

template <typename T>
void sddmm_reduce(pycmf_ctx* ctx, int mode, int64_t rows, const int32_t* rowptr, const int32_t* colidx,
                  const T* vals, const T* A, const T* B, int64_t k, double scale, double* out) {
    PYCMF_CHECK(k <= 32 * MAXT, "sddmm: n_components > 256 is not supported");
    int blocks = int(std::max<int64_t>(1, std::min<int64_t>(ceil_div(std::max<int64_t>(rows, 1), 8), 8 * ctx->num_sms)));
    double* part = static_cast<double*>(scratch(ctx, 1, size_t(blocks) * sizeof(double)));
    Timed timer(ctx, "sddmm");
    sddmm_reduce_kernel<T><<<blocks, 256, 0, ctx->stream>>>(mode, rows, rowptr, colidx, vals, A, B, int(k), part);
    PYCMF_LAUNCH_CHECK(ctx);
    final_sum(ctx, blocks, part, scale, out, true);
}

#define INSTANTIATE(T)                                                                                        \
    template void spmm<T>(pycmf_ctx*, int64_t, const int32_t*, const int32_t*, const T*, const T*, int64_t,   \
                          int64_t, T*, int64_t, T, T, int64_t);                                                      \
    template void sddmm_reduce<T>(pycmf_ctx*, int, int64_t, const int32_t*, const int32_t*, const T*,         \
                                  const T*, const T*, int64_t, double, double*);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace pycmf

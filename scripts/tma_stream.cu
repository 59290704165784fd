// Micro-benchmark (diagnostics, not product): what HBM bandwidth does a pure TMA ring reach on the access pattern of the
// tcgen05 passes?  Persistent CTAs walk 128 x 64 fp32 tiles of a row-major matrix (pitch 20000 B like the C2 workload)
// exactly like tc_pass_kernel does (LEFT: two boxes of 32 columns x 128 rows; RIGHT: four boxes of 32 columns x 64 rows),
// a consumer warp frees every stage as soon as it has landed.  Optionally an L2-hot "factor" box rides along.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tma_stream.bin scripts/tma_stream.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// mode 0 = LEFT (own = rows), 1 = RIGHT (own = columns); wide = 1: one box of 64 (LEFT) / 128 (RIGHT) columns without
// swizzle instead of 32-column swizzled boxes
__global__ void __launch_bounds__(64, 1)
tma_stream(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q, int mode, int wide,
           int nstage, int qbytes, int64_t T, int64_t G, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = 32768 + 16384;
    const uint32_t bars = base + nstage * stage_bytes;
    const int warp = threadIdx.x >> 5;
    const int64_t g0 = (int64_t(blockIdx.x) * G) / gridDim.x;
    const int n_it = int((int64_t(blockIdx.x + 1) * G) / gridDim.x - g0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (nstage + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
        for (int it = 0; it < n_it; it++) {
            const int s = it % nstage;
            mbar_wait(bars + 8 * (nstage + s), ((it / nstage) & 1) ^ 1);
            if (elect_one()) {
                const int64_t g = g0 + it;
                const int own0 = int((g / T) * 128), oth0 = int((g % T) * 64);
                const uint32_t st = base + s * stage_bytes, fb = bars + 8 * s;
                mbar_expect_tx(fb, 32768 + qbytes);
                if (mode == 0) {
                    if (wide) tma_load_2d(st, &tm_x, fb, oth0, own0);
                    else { tma_load_2d(st, &tm_x, fb, oth0, own0); tma_load_2d(st + 16384, &tm_x, fb, oth0 + 32, own0); }
                } else {
                    if (wide) tma_load_2d(st, &tm_x, fb, own0, oth0);
                    else for (int b = 0; b < 4; b++) tma_load_2d(st + b * 8192, &tm_x, fb, own0 + 32 * b, oth0);
                }
                for (int q = 0; q < qbytes / 8192; q++) tma_load_2d(st + 32768 + q * 8192, &tm_q, fb, 0, oth0);
            }
            __syncwarp();
        }
    } else {
        for (int it = 0; it < n_it; it++) {
            const int s = it % nstage;
            mbar_wait(bars + 8 * s, (it / nstage) & 1);
            if (elect_one()) mbar_arrive(bars + 8 * (nstage + s));
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn fn, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                            bool swizzle) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(float)};
    cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); exit(1); }
    return m;
}

int main() {
    const int64_t n = 20000, d = 5000;
    float *X, *Q;
    long long* out;
    cudaMalloc(&X, n * d * 4);
    cudaMalloc(&Q, 20000 * 32 * 4);
    cudaMalloc(&out, 1024 * 8);
    cudaMemset(X, 0, n * d * 4);
    cudaMemset(Q, 0, 20000 * 32 * 4);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%-6s %-5s %-7s %-7s %-5s %10s\n", "mode", "wide", "stages", "qbytes", "ctas", "GB/s");
    for (int mode = 0; mode < 2; mode++)
        for (int wide = 0; wide < 2; wide++) {
            const int bc = wide ? (mode == 0 ? 64 : 128) : 32, br = mode == 0 ? 128 : 64;
            CUtensorMap tm_x = make_map(fn, X, n, d, d, bc, br, !wide);
            CUtensorMap tm_q = make_map(fn, Q, 20000, 32, 32, 32, 64, true);
            const int64_t own_n = mode == 0 ? n : d, oth_n = mode == 0 ? d : n;
            const int64_t T = (oth_n + 63) / 64, G = ((own_n + 127) / 128) * T;
            for (int nstage : {2, 3, 4})
                for (int qbytes : {0, 16384})
                    for (int ctas : {148, 296}) {
                        const size_t smem = size_t(nstage) * 49152 + 1024 + 256;
                        if (ctas == 296 && smem > 110 * 1024) continue;
                        float best = 1e9f;
                        for (int rep = 0; rep < 3; rep++) {
                            cudaEventRecord(e0);
                            tma_stream<<<ctas, 64, smem>>>(tm_x, tm_q, mode, wide, nstage, qbytes, T, G, out);
                            cudaEventRecord(e1);
                            cudaEventSynchronize(e1);
                            float ms;
                            cudaEventElapsedTime(&ms, e0, e1);
                            if (ms < best) best = ms;
                        }
                        cudaError_t e = cudaGetLastError();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                        printf("%-6s %-5d %-7d %-7d %-5d %10.1f\n", mode ? "RIGHT" : "LEFT", wide, nstage, qbytes, ctas,
                               n * d * 4 / (best * 1e-3) / 1e9);
                    }
        }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}

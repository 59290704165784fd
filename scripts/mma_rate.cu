// Micro-benchmark (diagnostics, not product): issue rate of tcgen05.mma kind::tf32, M = 128, K = 8 per instruction,
// as a function of N, of where the A operand lives (shared memory = SS, tensor memory = TS) and of the number of
// accumulators the chain alternates between.  One CTA per SM; thread 0 issues `reps` MMAs back to back, commits to an
// mbarrier and waits; cycles = clock64 delta / reps.  Operand contents are irrelevant (zero-filled).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_rate.bin scripts/mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFF);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int kind_f16) {
    // D = F32; A = B = TF32 (format 2) or BF16 (format 1)
    return (1u << 4) | ((kind_f16 ? 1u : 2u) << 7) | ((kind_f16 ? 1u : 2u) << 10) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int form, int kind_f16>
__global__ void __launch_bounds__(128, 1) mma_rate(int N, int reps, int nacc, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar_store;
    const uint32_t bar = smem_u32(&bar_store);
    for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32 && elect_one()) {
        const uint32_t idesc = make_idesc(128, N, kind_f16);
        const uint64_t adesc = make_desc(base, 16, 1024);
        const uint64_t bdesc = make_desc(base + 16 * 1024, 16, 1024);
        const uint32_t a_tmem = tmem + 480;
        long long t0 = clock64();
        const uint32_t dstep = nacc == 2 ? 256u : 0u;
        for (int r0 = 0; r0 < reps; r0 += 8) {
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const uint32_t d = tmem + uint32_t(u & 1) * dstep;
            const int kk = u & 3;
            if (form == 0) {
                if (kind_f16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                                 ::"r"(d), "l"(adesc + uint64_t(kk * 2)), "l"(bdesc + uint64_t(kk * 2)), "r"(idesc), "r"(0u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                                 ::"r"(d), "l"(adesc + uint64_t(kk * 2)), "l"(bdesc + uint64_t(kk * 2)), "r"(idesc), "r"(0u) : "memory");
            } else {
                if (kind_f16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                                 ::"r"(d), "r"(a_tmem + uint32_t(kk * 8)), "l"(bdesc + uint64_t(kk * 2)), "r"(idesc), "r"(0u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                                 ::"r"(d), "r"(a_tmem + uint32_t(kk * 8)), "l"(bdesc + uint64_t(kk * 2)), "r"(idesc), "r"(0u) : "memory");
            }
          }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        asm volatile(
            "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
            "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar) : "memory");
        long long t2 = clock64();
        out[blockIdx.x * 2 + 0] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 2 * sizeof(long long));
    const int smem = 49 * 1024 + 1024;
    cudaFuncSetAttribute(mma_rate<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_rate<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_rate<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_rate<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 4096;
    printf("%-6s %-5s %-4s %-4s %12s %12s\n", "kind", "form", "N", "nacc", "issue clk", "done clk/MMA");
    for (int kind = 0; kind < 2; kind++)
        for (int form = 0; form < 2; form++)
            for (int N : {16, 32, 48, 64, 96, 128, 192, 256})
                for (int nacc : {1, 2}) {
                    if (nacc == 2 && N > 224) continue;            // second accumulator at column 256, A at 480
                    if (form == 1 && N > 224 && nacc == 1 && N + 0 > 480) continue;
                    for (int rep = 0; rep < 2; rep++) {
                        if (kind == 0 && form == 0) mma_rate<0, 0><<<148, 128, smem>>>(N, reps, nacc, out);
                        if (kind == 0 && form == 1) mma_rate<1, 0><<<148, 128, smem>>>(N, reps, nacc, out);
                        if (kind == 1 && form == 0) mma_rate<0, 1><<<148, 128, smem>>>(N, reps, nacc, out);
                        if (kind == 1 && form == 1) mma_rate<1, 1><<<148, 128, smem>>>(N, reps, nacc, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    long long h[296];
                    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    double issue = 0, done = 0;
                    for (int b = 0; b < 148; b++) { issue += h[2 * b]; done += h[2 * b + 1]; }
                    printf("%-6s %-5s %-4d %-4d %12.1f %12.1f\n", kind ? "bf16" : "tf32", form ? "TS" : "SS", N, nacc,
                           issue / 148 / reps, done / 148 / reps);
                }
    return 0;
}

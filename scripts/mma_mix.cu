// Micro-benchmark (diagnostics, not product): cost of the MMA issue pattern of tc_pass_kernel.  Per round: a GEMM1-like
// group (12 TS MMAs M128 N64 K8 + commit) and a GEMM2-like group (24 TS MMAs M128 N32 K8 + commit), (a) both issued by
// one thread, (b) by two warps concurrently, each waiting for its own commit of round r - 2 before issuing round r
// (so at most two rounds are in flight, like the double-buffered S / R tiles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_mix.bin scripts/mma_mix.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFF);
    d |= uint64_t((16 >> 4) & 0x3FFF) << 16;
    d |= uint64_t((1024 >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// mode 0: one warp issues both groups; mode 1: warp 0 issues the N=64 groups, warp 1 the N=32 groups
__global__ void __launch_bounds__(64, 1) mma_mix(int mode, int rounds, int depth, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bars[8];
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint64_t bdesc = make_desc(base);
    const long long t0 = clock64();
    // barrier slots: [0..1] group-1 buffers, [2..3] group-2 buffers
    auto group1 = [&](int r) {
        const uint32_t d = tmem + uint32_t((r & 1) * 64);
#pragma unroll
        for (int u = 0; u < 12; u++) mma_ts(d, tmem + 160 + (u & 3) * 8, bdesc + uint64_t((u & 3) * 2), make_idesc(128, 64));
        commit(smem_u32(&bars[r & 1]));
    };
    auto group2 = [&](int r) {
        const uint32_t d = tmem + 128u + uint32_t((u_int32_t)0);
#pragma unroll
        for (int u = 0; u < 24; u++) mma_ts(d, tmem + 256 + (r & 1) * 128 + (u & 7) * 8, bdesc + uint64_t((u & 3) * 2), make_idesc(128, 32));
        commit(smem_u32(&bars[2 + (r & 1)]));
    };
    if (mode == 0) {
        if (warp == 0) {
            for (int r = 0; r < rounds; r++) {
                if (r >= depth) { mbar_wait(smem_u32(&bars[r & 1]), ((r - 2) >> 1) & 1); mbar_wait(smem_u32(&bars[2 + (r & 1)]), ((r - 2) >> 1) & 1); }
                if (elect_one()) { group1(r); group2(r); }
                __syncwarp();
            }
        }
    } else {
        for (int r = 0; r < rounds; r++) {
            if (warp == 0) {
                if (r >= depth) mbar_wait(smem_u32(&bars[r & 1]), ((r - 2) >> 1) & 1);
                if (elect_one()) group1(r);
            } else {
                if (r >= depth) mbar_wait(smem_u32(&bars[2 + (r & 1)]), ((r - 2) >> 1) & 1);
                if (elect_one()) group2(r);
            }
            __syncwarp();
        }
    }
    // drain
    if (warp == 0 || mode == 1) {
        const int r = rounds - 1;
        if (warp == 0) mbar_wait(smem_u32(&bars[r & 1]), (r >> 1) & 1);
        if (warp == 1 || mode == 0) mbar_wait(smem_u32(&bars[2 + (r & 1)]), (r >> 1) & 1);
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    const int smem = 33 * 1024 + 1024;
    cudaFuncSetAttribute(mma_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int rounds = 512;
    printf("%-28s %12s   (tensor-pipe floor: 12 x 32 + 24 x 16 = 768 clk per round)\n", "pattern", "clk/round");
    for (int mode = 0; mode < 2; mode++)
        for (int depth : {2, 1000000}) {
            for (int rep = 0; rep < 2; rep++) {
                mma_mix<<<148, 64, smem>>>(mode, rounds, depth, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h[148];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0;
            for (int b = 0; b < 148; b++) s += h[b];
            printf("%-14s %-13s %12.1f\n", mode ? "two warps" : "one warp", depth == 2 ? "2 in flight" : "free running", s / 148 / rounds);
        }
    return 0;
}

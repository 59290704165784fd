// Micro-benchmark (diagnostics, not product): throughput of the LEGACY warp-level tensor path on B200 --
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 and m16n8k16 bf16 -- for kernels whose operands are gathered / built
// in registers (the per-row weighted Grams of the sampled Newton step), where tcgen05 with shared-memory descriptors does not
// fit directly.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/hmma_rate.bin scripts/hmma_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int KIND>
__global__ void __launch_bounds__(256) rate(int reps, float* out) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456f) out[0] = s;
}

int main() {
    float* out; cudaMalloc(&out, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int kind = 0; kind < 2; kind++)
        for (int bps = 1; bps <= 4; bps *= 2) {
            const int reps = 20000, blocks = p.multiProcessorCount * bps;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int w = 0; w < 2; w++) {
                cudaEventRecord(e0);
                if (kind == 0) rate<0><<<blocks, 256>>>(reps, out); else rate<1><<<blocks, 256>>>(reps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flop = double(blocks) * 8 /*warps*/ * reps * 8.0 * (kind == 0 ? 16.0 * 8 * 8 * 2 : 16.0 * 8 * 16 * 2);
            printf("%s  %d CTA/SM x 8 warps x 8 independent accumulators: %.3f ms  %.1f TFLOP/s\n",
                   kind == 0 ? "mma.sync m16n8k8 tf32 " : "mma.sync m16n8k16 bf16", bps, ms, flop / ms / 1e9);
        }
    return 0;
}

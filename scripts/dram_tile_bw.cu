// Micro-benchmark (diagnostics, not product): achieved HBM bandwidth when 148 x CPS persistent CTAs stream a row-major
// fp32 matrix (rows x cols, pitch = cols * 4 bytes) as tiles of TR rows x SEG contiguous bytes, each CTA walking its own
// row block left to right -- the access pattern of the tcgen05 passes (SEG = 256 B for LEFT, 512 B for RIGHT tiles).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dram_tile_bw.bin scripts/dram_tile_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// each warp owns TR / 8 rows of the tile; a lane reads 16 bytes; SEG / 16 lanes-worth of loads per row
template <int UNROLL>
__global__ void __launch_bounds__(256) stream_tiles(const float4* __restrict__ X, int64_t rows, int64_t pitch16, int tr,
                                                    int seg16, float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_blocks = rows / tr;
    const int64_t segs_per_row = pitch16 / seg16;
    float acc = 0.f;
    for (int64_t rb = blockIdx.x; rb < row_blocks; rb += gridDim.x) {
        for (int64_t sgm = 0; sgm < segs_per_row; sgm++) {
            // tile (rb, sgm): rows rb*tr .. +tr, 16-byte columns sgm*seg16 .. +seg16
            const int items = tr * seg16;                       // float4 items in the tile
            for (int base = warp * 32 * UNROLL; base < items; base += 8 * 32 * UNROLL) {
                float4 v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int it = base + u * 32 + lane;
                    const int r = it / seg16, c = it % seg16;
                    v[u] = it < items ? X[(rb * tr + r) * pitch16 + sgm * seg16 + c] : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
            }
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main() {
    const int64_t rows = 20480, cols = 5120;       // 400 MB, pitch 20480 B (a multiple of every SEG tested)
    float4* X;
    float* out;
    cudaMalloc(&X, rows * cols * 4);
    cudaMalloc(&out, 4);
    cudaMemset(X, 0, rows * cols * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%-6s %-6s %-5s %10s\n", "SEG_B", "rows", "cps", "GB/s");
    for (int cps : {1, 2, 4})
        for (int tr : {128, 64})
            for (int seg : {128, 256, 512, 1024, 2048, 4096, 20480}) {
                const int seg16 = seg / 16;
                float best = 1e9f;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(e0);
                    stream_tiles<8><<<148 * cps, 256>>>(X, rows, cols / 4, tr, seg16, out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("%-6d %-6d %-5d %10.1f\n", seg, tr, cps, rows * cols * 4 / (best * 1e-3) / 1e9);
            }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}

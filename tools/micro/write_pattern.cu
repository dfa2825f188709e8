// Micro-benchmark: HBM write bandwidth of the edge GEMM's store pattern (no compute).  148 persistent CTAs walk 128-row tiles;
// for every 64-column block 8 warps store [32 rows x 32 columns] pieces as 4 rows x 128 B per instruction, exactly as
// gemm_tc_kernel's epilogue does.  Layouts: (0) row-major [M][ncols] (row stride ncols*4 B: a block is 128 pieces of 256 B),
// (1) slice-major [ncols/128][M][128] (a tile's two blocks of a slice form one contiguous 64 KB run), (2) a plain grid-stride
// fill of the same bytes.  Optionally (PACE cycles) every block is preceded by a spin of that many cycles to mimic MMA time.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) k(float* C, long long M, int nblk, int layout, int pace) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wq = warp & 3, hh = warp >> 2, rsub = lane >> 3, cj = lane & 7;
    const long long ntile = (M + 127) / 128;
    const int ncols = nblk * 64;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long mt = blockIdx.x; mt < ntile; mt += gridDim.x) {
        for (int nb = 0; nb < nblk; nb++) {
            if (pace) { const long long t0 = clock64(); while (clock64() - t0 < pace) {} }
            const int c0 = nb * 64 + hh * 32 + cj * 4;
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const long long m = mt * 128 + wq * 32 + rr * 4 + rsub;
                if (m < M) {
                    float* p = layout == 0 ? C + m * ncols + c0 : C + ((long long)(c0 >> 7) * M + m) * 128 + (c0 & 127);
                    *reinterpret_cast<float4*>(p) = v;
                }
            }
        }
    }
}
__global__ void fill(float4* C, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) C[i] = make_float4(1, 2, 3, 4);
}
int main(int argc, char** argv) {
    const long long M = 1024LL * 30 * 29;
    float* C; cudaMalloc(&C, M * 896 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pace : {0, 600, 1100}) {
        for (int nblk : {4, 10, 14}) {
            for (int layout = 0; layout < 3; layout++) {
                if (layout == 2 && pace) continue;
                float ms = 0;
                for (int rep = 0; rep < 6; rep++) {
                    if (rep == 1) cudaEventRecord(e0);
                    if (layout < 2) k<<<148, 256>>>(C, M, nblk, layout, pace);
                    else fill<<<148 * 8, 256>>>((float4*)C, M * nblk * 16);
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                printf("pace %4d  cols %3d  layout %d: %.3f ms  %.0f GB/s   %s\n", pace, nblk * 64, layout, ms, M * nblk * 256.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}

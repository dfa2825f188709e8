// Probe: which (TMEM lane, column) does each register of tcgen05.ld.16x256b.x4 hold?  A warp writes value = lane * 100 + column
// with the 32x32b shape (thread = lane), reads the same 32 x 32 block back with two 16x256b.x4 loads (lanes 0..15 and 16..31)
// and prints the mapping of a few threads; then checks the closed form used by the attention kernels' epilogues:
//   register j of the load at lane offset 16 h  ->  lane 16 h + 8 ((j >> 1) & 1) + t / 4,  column 8 (j >> 2) + 2 (t % 4) + (j & 1)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_transpose_probe tmem_transpose_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../phoregen_b200/csrc/pg_tc.cuh"

__global__ void __launch_bounds__(128, 1) k(int* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc<512>(&slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; c++) r[c] = (warp * 32 + lane) * 100 + c;
    tc::tmem_st32(tmem + lane_base + 64, r);
    tc::tmem_st_wait();
    uint32_t v[2][16];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[h][0]), "=r"(v[h][1]), "=r"(v[h][2]), "=r"(v[h][3]), "=r"(v[h][4]), "=r"(v[h][5]), "=r"(v[h][6]), "=r"(v[h][7]),
              "=r"(v[h][8]), "=r"(v[h][9]), "=r"(v[h][10]), "=r"(v[h][11]), "=r"(v[h][12]), "=r"(v[h][13]), "=r"(v[h][14]), "=r"(v[h][15])
            : "r"(tmem + lane_base + ((uint32_t)(h * 16) << 16) + 64)
            : "memory");
    }
    tc::tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int j = 0; j < 16; j++) out[(tid * 2 + h) * 16 + j] = (int)v[h][j];
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}
int main() {
    int* d; cudaMalloc(&d, 128 * 32 * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    static int h[128 * 32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status: %s\n", cudaGetErrorString(e));
    for (int t : {0, 1, 2, 3, 4, 5, 31, 32 + 7, 96 + 30}) {
        printf("thread %3d:", t);
        for (int hh = 0; hh < 2; hh++) {
            for (int j = 0; j < 16; j++) printf(" %d.%02d", h[(t * 2 + hh) * 16 + j] / 100, h[(t * 2 + hh) * 16 + j] % 100);
            printf(" |");
        }
        printf("\n");
    }
    int bad = 0;
    for (int t = 0; t < 128; t++)
        for (int hh = 0; hh < 2; hh++)
            for (int j = 0; j < 16; j++) {
                const int l = t & 31, w = t >> 5;
                const int row = w * 32 + 16 * hh + 8 * ((j >> 1) & 1) + l / 4, col = 8 * (j >> 2) + 2 * (l % 4) + (j & 1);
                if (h[(t * 2 + hh) * 16 + j] != row * 100 + col) bad++;
            }
    printf("closed form mismatches: %d of %d\n", bad, 128 * 32);
    return 0;
}

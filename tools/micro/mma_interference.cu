// Micro-benchmark: how much does each kind of row-warp traffic slow down a stream of tcgen05.mma (kind::f16, M128 N128 K16,
// A from TMEM, B from 128B-swizzled shared memory: the second-Linear MMAs of the attention kernels)?  One CTA per SM, one
// issuing warp + NW "row" warps that hammer one resource until the issuer is done:
//   0 nothing | 1 tcgen05.ld 32x32b.x32 | 2 tcgen05.st x16 (two per iteration) | 3 LDS.128 | 4 SHFL | 5 FFMA2 | 6 STS.128 |
//   7 ld + st + shfl + ffma mix | 9 REDUX
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_interference mma_interference.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../phoregen_b200/csrc/pg_tc.cuh"

template <int MODE>
__global__ void __launch_bounds__(640, 1) k(long long* out, int iters, int nw, const float* gsrc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, tbar; __shared__ uint32_t slot; __shared__ volatile int done;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < (96 * 1024) / 4; i += 640) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (warp == 16) tc::tmem_alloc<512>(&slot);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&tbar, 1); done = 0; tc::fence_barrier_init(); }
    tc::fence_proxy_async_smem(); tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 16) {
        const uint32_t idesc = tc::umma_idesc_bf16(128, 128);
        const uint32_t b_s = tc::smem_u32(smem);
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            uint32_t acc = 0;
#pragma unroll
            for (int rep = 0; rep < 3; rep++)
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint64_t bd = tc::umma_desc_sw128(b_s + (rep & 1) * 32768 + (ks >> 2) * 16384 + (ks & 3) * 32);
                    tc::umma_bf16_ts_w(tmem, tmem + 256 + (rep == 2 ? 64 : 0) + ks * 8, bd, idesc, acc);
                    acc = 1;
                }
            tc::umma_commit_w(&bar);
            tc::mbar_wait(&bar, ph); ph ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = (t1 - t0);
        done = 1;
    } else if (warp < nw) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int cq = warp >> 2;
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; i++) r[i] = tid + i;
        float2 f = make_float2(1.0f + tid, 2.0f), g = make_float2(0.999f, 0.001f);
        float* sbuf = (float*)(smem + 65536 + warp * 1024);
        while (!done) {
#pragma unroll 1
            for (int rep = 0; rep < 16; rep++) {
                if (MODE == 1 || MODE == 7) { tc::tmem_ld32_nowait(tmem + lane_base + 128 + cq * 32, r); tc::tmem_ld_wait(); }
                if (MODE == 2 || MODE == 7) {
                    uint32_t h[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) h[i] = r[i] + rep;
                    tc::tmem_st16(tmem + lane_base + 384 + cq * 16, h);
                    tc::tmem_st16(tmem + lane_base + 384 + 64 + cq * 16, h);
                    tc::tmem_st_wait();
                }
                if (MODE == 3) {
#pragma unroll
                    for (int i = 0; i < 8; i++) { const float4 v = *reinterpret_cast<const float4*>(sbuf + ((lane * 4 + i * 128) & 255)); f.x += v.x + v.y + v.z + v.w; }
                }
                if (MODE == 6) {
#pragma unroll
                    for (int i = 0; i < 8; i++) *reinterpret_cast<float4*>(sbuf + ((lane * 4 + i * 128) & 255)) = make_float4(f.x, f.y, g.x, g.y + i);
                }
                if (MODE == 4 || MODE == 7) {
#pragma unroll
                    for (int i = 0; i < 16; i++) f.x += __shfl_xor_sync(0xffffffffu, f.x, 1 + (i & 15));
                }
                if (MODE == 5 || MODE == 7) {
#pragma unroll
                    for (int i = 0; i < 32; i++) f = tc::fma2(f, g, g);
                }
                if (MODE == 9) {
#pragma unroll
                    for (int i = 0; i < 4; i++) f.x = tc::warp_max_redux(f.x + i);
                }
            }
        }
        if (f.x == 123.f && r[0] == 77) out[1] = (long long)f.y + r[5];
    }
    (void)gsrc;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}
template <int MODE> void run(const char* name, int nw) {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    const int iters = 500;
    k<MODE><<<148, 640, 120 * 1024>>>(d, iters, nw, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s row warps %2d: %6.1f cycles per MMA (24 per commit)  %s\n", name, nw, (double)h / (iters * 24.0), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<0>("nothing", 0);
    for (int nw : {4, 16}) {
        run<1>("tcgen05.ld x32", nw);
        run<2>("tcgen05.st x16 x2", nw);
        run<3>("LDS.128", nw);
        run<6>("STS.128", nw);
        run<4>("SHFL", nw);
        run<5>("FFMA2", nw);
        run<9>("REDUX", nw);
        run<7>("ld + st + shfl + ffma", nw);
    }
    return 0;
}

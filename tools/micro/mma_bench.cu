// Micro-benchmark: issue rate of tcgen05.mma kind::f16 M128 N{64,128,256} K16, A from smem (SS) or TMEM (TS), one CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../phoregen_b200/csrc/pg_tc.cuh"
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (warp == 0) tc::tmem_alloc<512>(&slot);
    if (tid == 32) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    tc::fence_proxy_async_smem(); tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = tc::umma_idesc_bf16(128, N);
        const uint32_t a_s = tc::smem_u32(smem), b_s = tc::smem_u32(smem + 32768);
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            uint32_t acc = 0;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const uint64_t bd = tc::umma_desc_sw128(b_s + (ks >> 2) * 16384 + (ks & 3) * 32);
                if (TS) tc::umma_bf16_ts(tmem, tmem + 256 + ks * 8, bd, idesc, acc);
                else tc::umma_bf16(tmem, tc::umma_desc_sw128(a_s + (ks >> 2) * 16384 + (ks & 3) * 32), bd, idesc, acc);
                acc = 1;
            }
            tc::umma_commit(&bar);
            tc::mbar_wait(&bar, ph); ph ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0);
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}
template <int N, bool TS> void run(const char* name) {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    k<N, TS><<<148, 128, 100 * 1024>>>(d, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s N=%d: %.1f cycles per MMA (8 per commit, incl. commit+wait)  err=%s\n", name, N, (double)h / (iters * 8.0), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<64, false>("SS"); run<128, false>("SS"); run<256, false>("SS");
    run<64, true>("TS"); run<128, true>("TS"); run<256, true>("TS");
    return 0;
}

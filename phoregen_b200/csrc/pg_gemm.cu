// Dense per-node / per-edge contractions with K = 128 (hidden width): C[M, nt*128] = pro(A)[M,128] @ Wt[128, nt*128] + b.
// fp32 SIMT tile kernel (128 x 128 x 128 per CTA, 8x8 outputs per thread) with a fused A-prologue
// (sum of two inputs, or gather-add + LayerNorm + ReLU, i.e. the tail of models/common.py:99-119's first
// Linear) and a fused epilogue (bias, residual, ReLU).  The A tile is loaded once per CTA and reused for
// every 128-column block of the weight, so h_bond is read once per edge GEMM.
#include "pg_gemm.h"

namespace {
constexpr int TM = 128, TK = 128, TN = 128, AS_LD = 132;

template <int PRO>
__global__ void __launch_bounds__(256, 1) gemm_k128_kernel(GemmArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                   // [128][132]
    float* Bs = sm + TM * AS_LD;      // [128][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * TM;

    float4 g4 = make_float4(1, 1, 1, 1), b4 = make_float4(0, 0, 0, 0);
    if (PRO == PRO_LNRELU) { g4 = ldg4(a.ln_g + lane * 4); b4 = ldg4(a.ln_b + lane * 4); }
    for (int r = warp; r < TM; r += 8) {
        const long long m = m0 + r;
        float4 v = make_float4(0, 0, 0, 0);
        if (m < a.M) {
            v = ld4(a.A + m * a.lda + lane * 4);
            if (PRO == PRO_SUM2) v = f4add(v, ld4(a.A2 + m * a.lda2 + lane * 4));
            if (PRO == PRO_LNRELU) {
                if (a.A2) {
                    const long long idx = a.gidx ? (long long)a.gidx[m] : m;
                    v = f4add(v, ld4(a.A2 + idx * a.lda2 + lane * 4));
                }
                v = ln_relu_row(v, g4, b4);
            }
        }
        st4(As + r * AS_LD + lane * 4, v);
    }
    const int tx = tid & 15, ty = tid >> 4;
    for (int nt = 0; nt < a.ntiles; nt++) {
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int idx = tid + i * 256;
            const int k = idx >> 5, c4 = idx & 31;
            st4(Bs + k * TN + c4 * 4, ldg4(a.Wt + (long long)k * a.ldw + nt * TN + c4 * 4));
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 2
        for (int k4 = 0; k4 < TK / 4; k4++) {
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
                av[i] = ld4(As + row * AS_LD + k4 * 4);
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const float4 b0 = ld4(Bs + (k4 * 4 + kk) * TN + tx * 4);
                const float4 b1 = ld4(Bs + (k4 * 4 + kk) * TN + 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float s = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
                    acc[i][0] = fmaf(s, b0.x, acc[i][0]); acc[i][1] = fmaf(s, b0.y, acc[i][1]);
                    acc[i][2] = fmaf(s, b0.z, acc[i][2]); acc[i][3] = fmaf(s, b0.w, acc[i][3]);
                    acc[i][4] = fmaf(s, b1.x, acc[i][4]); acc[i][5] = fmaf(s, b1.y, acc[i][5]);
                    acc[i][6] = fmaf(s, b1.z, acc[i][6]); acc[i][7] = fmaf(s, b1.w, acc[i][7]);
                }
            }
        }
        const int c0 = nt * TN + tx * 4, c1 = c0 + 64;
        float4 bb0 = make_float4(0, 0, 0, 0), bb1 = bb0;
        if (a.bias) { bb0 = ldg4(a.bias + c0); bb1 = ldg4(a.bias + c1); }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
            const long long m = m0 + row;
            if (m >= a.M) continue;
            float4 v0 = make_float4(acc[i][0] + bb0.x, acc[i][1] + bb0.y, acc[i][2] + bb0.z, acc[i][3] + bb0.w);
            float4 v1 = make_float4(acc[i][4] + bb1.x, acc[i][5] + bb1.y, acc[i][6] + bb1.z, acc[i][7] + bb1.w);
            if (a.resid) {
                v0 = f4add(v0, ld4(a.resid + m * a.ldr + c0));
                v1 = f4add(v1, ld4(a.resid + m * a.ldr + c1));
            }
            if (a.relu) {
                v0 = make_float4(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f), fmaxf(v0.z, 0.f), fmaxf(v0.w, 0.f));
                v1 = make_float4(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f), fmaxf(v1.z, 0.f), fmaxf(v1.w, 0.f));
            }
            st4(a.C + m * a.ldc + c0, v0);
            st4(a.C + m * a.ldc + c1, v1);
        }
    }
}
}  // namespace

int pg_launch_gemm(const GemmArgs& a, int pro, cudaStream_t stream) {
    if (a.M <= 0) return PG_OK;
    if (a.csplit > 0) { pg_set_error("fp32 GEMM: split outputs are a feature of the tcgen05 kernel"); return PG_EINVAL; }
    const size_t smem = (size_t)(TM * AS_LD + TK * TN) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_k128_kernel<PRO_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_k128_kernel<PRO_SUM2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_k128_kernel<PRO_LNRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const unsigned grid = (unsigned)((a.M + TM - 1) / TM);
    switch (pro) {
        case PRO_PLAIN: gemm_k128_kernel<PRO_PLAIN><<<grid, 256, smem, stream>>>(a); break;
        case PRO_SUM2: gemm_k128_kernel<PRO_SUM2><<<grid, 256, smem, stream>>>(a); break;
        case PRO_LNRELU: case PRO_LNRELU_MF: gemm_k128_kernel<PRO_LNRELU><<<grid, 256, smem, stream>>>(a); break;
        default: pg_set_error("bad gemm prologue"); return PG_EINVAL;
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

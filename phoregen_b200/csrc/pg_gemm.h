#pragma once
#include "pg_common.cuh"

// PRO_LNRELU_MF: LayerNorm + ReLU prologue for inputs that are already mean-free (first-Linear blocks centred at pack time,
// weights._center_first_linears): only the second moment is reduced.  The fp32 kernel runs it as PRO_LNRELU.
enum { PRO_PLAIN = 0, PRO_SUM2 = 1, PRO_LNRELU = 2, PRO_LNRELU_MF = 3 };

struct GemmArgs {
    long long M;
    const float* A; long long lda;       // input rows (first 128 columns starting at A)
    const float* A2; long long lda2;     // SUM2: second input; LNRELU: optional gather-add rows A2[gidx[m]]
    const int* gidx;                     // optional row gather for A2
    const float* ln_g; const float* ln_b;
    const float* Wt; long long ldw;      // [128][ldw] (k-major); column block nt starts at nt*128   (fp32 SIMT kernel)
    const float* Wbf;                    // bf16 hi|lo split of the same weight, [2][ntiles*128][128] K-major (tcgen05 kernel)
    const float* bias;                   // [ntiles*128] or null
    float* C; long long ldc;
    float* C2; long long ldc2; int csplit;   // optional second output (tcgen05 kernel): columns >= csplit go to C2[m * ldc2 + c - csplit]
    int ntiles;
    const float* resid; long long ldr;   // optional residual added in the epilogue (may alias C)
    int relu;
};

int pg_launch_gemm(const GemmArgs& a, int pro, cudaStream_t stream);      // fp32 FFMA reference kernel (PG_GEMM=simt)
int pg_launch_gemm_tc(const GemmArgs& a, int pro, cudaStream_t stream);   // tcgen05 / TMEM bf16x3 kernel (default)

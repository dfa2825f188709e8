// Shared device helpers and internal structs for the phoregen_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/phoregen_b200.h"

#define PG_FULL 0xffffffffu

void pg_set_error(const char* fmt, ...);
#define PG_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            pg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PG_ECUDA;                                                             \
        }                                                                                \
    } while (0)
#define PG_LAUNCH_CHECK() PG_CUDA_CHECK(cudaGetLastError())

// ---------------------------------------------------------------- batch plan (device view)
// Context order per graph g: [phore nodes (p_g) | ligand atoms (n_g)].
// Internal bond-edge order per graph: dst-major, src ascending, self excluded:
//   e(j -> i) = eoff[g] + i*(n-1) + (j - (j > i)),  i = dst, j = src (ligand-local indices).
struct PlanDev {
    int G, N, Nl, P;
    long long Eb, Ek, E3;
    int max_n, max_p, max_ng, min_n;
    const int* ctx_off;     // [G+1] context node offsets
    const int* lig_off;     // [G+1]
    const int* ph_off;      // [G+1]
    const long long* eoff;  // [G+1] bond-edge offsets
    const long long* koff;  // [G+1] kNN-edge offsets (graph g has N_g * min(32, N_g-1))
    const long long* t3off; // [G+1] triplet offsets
    const int* g_n;         // [G]
    const int* g_p;         // [G]
    const int* node_graph;  // [N]  context node -> graph
    const int* lig_graph;   // [Nl] ligand atom -> graph
    const int* ph_graph;    // [P]
    const int* perm;        // [Eb] reference edge order -> internal order
    const int* edge_graph;  // [Eb] reference-order edge -> graph
    const int* esrc_node;   // [Eb] internal edge -> context node of src (j)
    const int* edst_node;   // [Eb] internal edge -> context node of dst (i)
    // bond-graph attention tiles of the molecules whose segments are longer than one 32-row TMEM lane quarter
    // (pg_bond_tc.cu): a graph whose segments take C = ceil((n-1)/32) > 1 quarters places 4/C atoms in a tile,
    // ceil(n / (4/C)) tiles per graph; molecules with n-1 <= 32 have no entry (they run on the packed instantiation)
    int nbt;                // number of such tiles in the batch
    const int* btile_off;   // [G+1] first tile of graph g
    const int* btile_graph; // [nbt] tile -> graph
};
__host__ __device__ inline int pg_bond_chunks(int n) { return (n - 1 + 31) >> 5; }                // quarters per segment
__host__ __device__ inline int pg_bond_atoms_per_tile(int n) { const int c = pg_bond_chunks(n); return c <= 4 ? 4 / c : 0; }

// ---------------------------------------------------------------- small vector helpers
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 process two fp32 lanes per issued instruction; every lane is
// a correctly rounded fp32 operation, so results equal the scalar ones).  The 4-wide row helpers are built on them.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float2 up2(unsigned long long v) {
    return make_float2(__uint_as_float((unsigned)v), __uint_as_float((unsigned)(v >> 32)));
}
__device__ __forceinline__ unsigned long long fma2_raw(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2_raw(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2_raw(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// (measured: routing these 4-wide helpers through FFMA2 does not help the latency/LSU-bound fp32 kernels)
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PG_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PG_FULL, v, o));
    return v;
}

// LayerNorm(128, eps 1e-5, affine) + ReLU over a row spread as 4 channels per lane (models/common.py:113-115).
__device__ __forceinline__ float4 ln_relu_row(float4 v, float4 g, float4 b) {
    float mu = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / 128.0f);
    float4 d = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    float var = warp_sum(fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, d.w * d.w)))) * (1.0f / 128.0f);
    float rstd = rsqrtf(var + 1e-5f);
    return make_float4(fmaxf(fmaf(d.x * rstd, g.x, b.x), 0.f), fmaxf(fmaf(d.y * rstd, g.y, b.y), 0.f),
                       fmaxf(fmaf(d.z * rstd, g.z, b.z), 0.f), fmaxf(fmaf(d.w * rstd, g.w, b.w), 0.f));
}

// Butterfly transpose-reduce: every lane holds 16 partial sums p[0..15]; afterwards the full sum of
// p[h] over the warp is returned on lanes 2h and 2h+1  (h = (lane >> 1) & 15).  16 shuffles.
__device__ __forceinline__ float transpose_reduce16(float (&p)[16], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float send = up ? p[i] : p[i + 8];
            float keep = up ? p[i + 8] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float send = up ? p[i] : p[i + 4];
            float keep = up ? p[i + 4] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float send = up ? p[i] : p[i + 2];
            float keep = up ? p[i + 2] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        float send = up ? p[0] : p[1];
        float keep = up ? p[1] : p[0];
        p[0] = keep + __shfl_xor_sync(PG_FULL, send, 2);
    }
    return p[0] + __shfl_xor_sync(PG_FULL, p[0], 1);
}

// 32 partial sums per lane -> lane L returns the warp-wide sum of p[L].  31 shuffles.
__device__ __forceinline__ float transpose_reduce32(float (&p)[32], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float send = up ? p[i] : p[i + 16];
            float keep = up ? p[i + 16] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float send = up ? p[i] : p[i + 8];
            float keep = up ? p[i + 8] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float send = up ? p[i] : p[i + 4];
            float keep = up ? p[i + 4] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 4);
        }
    }
    {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float send = up ? p[i] : p[i + 2];
            float keep = up ? p[i + 2] : p[i];
            p[i] = keep + __shfl_xor_sync(PG_FULL, send, 2);
        }
    }
    {
        const bool up = lane & 1;
        float send = up ? p[0] : p[1];
        float keep = up ? p[1] : p[0];
        return keep + __shfl_xor_sync(PG_FULL, send, 1);
    }
}

// GaussianSmearing offsets (models/common.py:18), coeff -0.5.
__constant__ float c_smear_off[20] = {0.f, 1.f, 1.25f, 1.5f, 1.75f, 2.f, 2.25f, 2.5f, 2.75f, 3.f,
                                      3.5f, 4.f, 4.5f, 5.f, 5.5f, 6.f, 7.f, 8.f, 9.f, 10.f};
__device__ __forceinline__ float smear_val(float dist, int g) {
    float d = dist - c_smear_off[g];
    return expf(-0.5f * d * d);
}

// Philox4x32-10 (counter-based RNG; Salmon et al. 2011).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

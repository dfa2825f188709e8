// BondUpdateLayer (uni_denoiser.py:123-165) on the 5th-gen tensor cores.
//
// One persistent CTA per SM walks the ligand atoms j ("units").  For a unit the per-edge partials
// P[k->j] (n-1 rows x 256 channels, k|v) are staged once in shared memory; the n-1 segments (j->i) are processed
// four at a time as a 128-row tile: TMEM lane = thread = triplet row (32 lanes per segment, rows k ascending).
//
//   1. angular encoding (13 values per triplet) -> bf16 hi/lo A tile in smem; one tcgen05.mma (M128 N256 K16, x3 for
//      bf16x3) against the angle slice of the first Linear gives the triplet-specific part of both pre-activations
//      in TMEM columns [0,256).
//   2. every thread reads its row (tcgen05.ld), adds P[k->j] (smem) and R[j->i] (the r_ji slice, per segment),
//      applies LayerNorm + ReLU in registers (no shuffles: the row is thread-local), splits to bf16 hi/lo and writes
//      the result back to TMEM (tcgen05.st) as the A operand of the second Linear.
//   3. second Linear of the key and value MLPs: A from TMEM, B = W2 (bf16 hi/lo, resident in smem, 128B swizzle),
//      24 tcgen05.mma (M128 N128 K16) each, fp32 accumulators in TMEM.
//   4. epilogue: logits = q . k / sqrt(8) per head (thread-local dot), segment softmax across the 32 lanes of the
//      warp, alpha-weighted sum of v over the lanes with a butterfly transpose-reduce, residual add into h_bond.
//
// TMEM map (512 columns): [0,128) pre_k -> reused for hid_v ; [128,256) pre_v -> reused for out_v ;
//                         [256,384) out_k ; [384,512) hid_k.       (hid = 64 columns hi + 64 columns lo)
// Precision: bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate) on every contraction, fp32 everywhere else.
#include <algorithm>
#include "pg_attn.h"
#include "pg_tc.cuh"

namespace {
constexpr int PS_LD = 260;                 // padded row stride of the staged P rows (conflict-free LDS.128 across rows)
constexpr int W_TILE = 32768;              // one [128 x 128] bf16 matrix in two 128B-swizzled K blocks
constexpr int SM_W = 4 * W_TILE;           // (k,v) x (hi,lo)
constexpr int SM_WA = 2 * 8192;            // angle slice of the first Linear: (hi,lo) x [256 x 16] bf16, no swizzle
constexpr int SM_FEAT = 2 * 4096;          // (hi,lo) x [128 x 16] bf16, no swizzle
constexpr int SM_FIXED = SM_W + SM_WA + SM_FEAT + 4 * 128 * 4 /*q*/ + 6 * 128 * 4 /*ln + b2*/ + 64 /*barriers*/;
constexpr float kInvSqrtD = 0.35355339059327373f;

__global__ void __launch_bounds__(256, 1) trip_tc_kernel(TripTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;
    uint8_t* sWa = sW + SM_W;
    uint8_t* sFeat = sWa + SM_WA;
    float* sQ = (float*)(sFeat + SM_FEAT);          // [4][128]
    float* sLn = sQ + 4 * 128;                      // gk, bk, gv, bv
    float* sB2 = sLn + 4 * 128;                     // b2k, b2v
    uint64_t* bars = (uint64_t*)(sB2 + 2 * 128);    // 3 mbarriers
    uint32_t* tmem_slot = (uint32_t*)(bars + 4);
    float* sPs = (float*)(smem + SM_FIXED);         // [(maxn-1)][260]
    float* sSmr = sPs + (size_t)(a.maxn - 1) * PS_LD;   // [(maxn-1)][20]
    const PlanDev& d = a.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp & 3, half = warp >> 2;

    if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 32) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::mbar_init(&bars[2], 1); tc::fence_barrier_init(); }
    // ---- resident weights
    for (int idx = tid; idx < 4 * 128 * 16; idx += 256) {     // 16-byte chunks: [mat 4][n 128][chunk 16]
        const int mat = idx >> 11, n = (idx >> 4) & 127, c = idx & 15;
        const uint16_t* src = ((mat >> 1) ? a.w2v_bf : a.w2k_bf) + ((size_t)(mat & 1) * 128 + n) * 128 + c * 8;
        const uint32_t dst = tc::smem_u32(sW) + mat * W_TILE + (c >> 3) * 16384 + tc::sw128_chunk(n, c & 7);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    for (int idx = tid; idx < 2 * 256 * 2; idx += 256) {      // [part 2][n 256][kc 2]
        const int part = idx >> 9, n = (idx >> 1) & 255, kc = idx & 1;
        const uint16_t* src = a.wa_bf + ((size_t)part * 256 + n) * 16 + kc * 8;
        const uint32_t dst = tc::smem_u32(sWa) + part * 8192 + (n >> 3) * 256 + kc * 128 + (n & 7) * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    if (tid < 128) {
        sLn[tid] = a.lnk_g[tid]; sLn[128 + tid] = a.lnk_b[tid]; sLn[256 + tid] = a.lnv_g[tid]; sLn[384 + tid] = a.lnv_b[tid];
        sB2[tid] = a.b2k[tid]; sB2[128 + tid] = a.b2v[tid];
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    constexpr uint32_t idesc_feat = tc::umma_idesc_bf16(128, 256);
    constexpr uint32_t idesc_w2 = tc::umma_idesc_bf16(128, 128);
    const uint32_t sW_u32 = tc::smem_u32(sW), sWa_u32 = tc::smem_u32(sWa), sFeat_u32 = tc::smem_u32(sFeat);
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;

    for (int u = blockIdx.x; u < d.Nl; u += gridDim.x) {
        const int g = d.lig_graph[u];
        const int n = d.g_n[g], jl = u - d.lig_off[g];
        if (n < 3 || n - 2 > 32) continue;                  // larger molecules take the fp32 kernel
        const int ctx0 = d.ctx_off[g] + d.g_p[g];
        const int cj = ctx0 + jl;
        const long long eoff = d.eoff[g];
        const float xj0 = a.x[(size_t)cj * 3], xj1 = a.x[(size_t)cj * 3 + 1], xj2 = a.x[(size_t)cj * 3 + 2];
        // ---- stage P rows of the edges k -> j
        for (int t = tid; t < n - 1; t += 256) {
            const int ck = ctx0 + t + (t >= jl);
            const float d0 = xj0 - a.x[(size_t)ck * 3], d1 = xj1 - a.x[(size_t)ck * 3 + 1], d2 = xj2 - a.x[(size_t)ck * 3 + 2];
            const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
#pragma unroll
            for (int gg = 0; gg < 20; gg++) sSmr[t * 20 + gg] = smear_val(dist, gg);
        }
        __syncthreads();
        {
            const int c = tid, cc = c & 127;
            const bool isv = c >= 128;
            const float hjb = __ldg(a.H + (size_t)cj * a.ldh + (isv ? a.hj_v : a.hj_k) + cc);
            float wr[20];
#pragma unroll
            for (int gg = 0; gg < 20; gg++) wr[gg] = __ldg(a.wrkj + gg * 256 + c);
            const int tcol = (isv ? a.t_v : a.t_k) + cc, hcol = (isv ? a.hk_v : a.hk_k) + cc;
            for (int t = 0; t < n - 1; t++) {
                const int ck = ctx0 + t + (t >= jl);
                const long long e = eoff + (long long)jl * (n - 1) + t;
                float val = __ldg(a.T + (size_t)e * a.ldt + tcol) + __ldg(a.H + (size_t)ck * a.ldh + hcol) + hjb;
#pragma unroll
                for (int gg = 0; gg < 20; gg++) val = fmaf(sSmr[t * 20 + gg], wr[gg], val);
                sPs[t * PS_LD + c] = val;
            }
        }
        __syncthreads();

        const int R = n - 2;
        const int ntile = (n - 1 + 3) >> 2;
        for (int tile = 0; tile < ntile; tile++) {
            const int sidx = tile * 4 + wq;
            const bool segvalid = sidx < n - 1;
            const int il = segvalid ? sidx + (sidx >= jl) : (jl == 0 ? 1 : 0);
            const int ti = il - (il > jl);
            const bool rowvalid = segvalid && lane < R;
            const int trow = rowvalid ? lane + (lane >= ti) : 0;
            const long long eji = eoff + (long long)il * (n - 1) + (jl - (jl > il));
            // ---- phase 0: angular features (warps 0-3), query rows (warps 4-7)
            if (half == 0) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; i++) f[i] = 0.f;
                if (rowvalid) {
                    const int ci = ctx0 + il, ck = ctx0 + trow + (trow >= jl);
                    const float xi0 = a.x[(size_t)ci * 3], xi1 = a.x[(size_t)ci * 3 + 1], xi2 = a.x[(size_t)ci * 3 + 2];
                    const float pj0 = xj0 - xi0, pj1 = xj1 - xi1, pj2 = xj2 - xi2;
                    const float pk0 = a.x[(size_t)ck * 3] - xi0, pk1 = a.x[(size_t)ck * 3 + 1] - xi1, pk2 = a.x[(size_t)ck * 3 + 2] - xi2;
                    const float dotv = pj0 * pk0 + pj1 * pk1 + pj2 * pk2;
                    const float c0 = pj1 * pk2 - pj2 * pk1, c1 = pj2 * pk0 - pj0 * pk2, c2 = pj0 * pk1 - pj1 * pk0;
                    const float th = atan2f(sqrtf(c0 * c0 + c1 * c1 + c2 * c2), dotv);
                    float s1, k1, s2, k2, s3, k3, sh, kh, st, kt;
                    sincosf(th, &s1, &k1); sincosf(th * 2.0f, &s2, &k2); sincosf(th * 3.0f, &s3, &k3);
                    sincosf(th * 0.5f, &sh, &kh); sincosf(th * (1.0f / 3.0f), &st, &kt);
                    f[0] = th; f[1] = s1; f[2] = s2; f[3] = s3; f[4] = s1; f[5] = sh; f[6] = st;
                    f[7] = k1; f[8] = k2; f[9] = k3; f[10] = k1; f[11] = kh; f[12] = kt;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    tc::split_bf16(f[2 * i], h0, l0); tc::split_bf16(f[2 * i + 1], h1, l1);
                    hi[i] = tc::pack_bf16(h0, h1); lo[i] = tc::pack_bf16(l0, l1);
                }
                const int r = wq * 32 + lane;
                uint8_t* ph = sFeat + (r >> 3) * 256 + (r & 7) * 16;
                *reinterpret_cast<uint4*>(ph) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(ph + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<uint4*>(ph + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(ph + 4096 + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            } else {
                st4(sQ + wq * 128 + lane * 4, segvalid ? ldg4(a.q + (size_t)eji * 128 + lane * 4) : make_float4(0, 0, 0, 0));
            }
            tc::fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc::tc_fence_after();
                const uint64_t fh = tc::umma_desc_k16_noswizzle(sFeat_u32), fl = tc::umma_desc_k16_noswizzle(sFeat_u32 + 4096);
                const uint64_t wh = tc::umma_desc_k16_noswizzle(sWa_u32), wl = tc::umma_desc_k16_noswizzle(sWa_u32 + 8192);
                tc::umma_bf16(tmem, fh, wh, idesc_feat, 0);
                tc::umma_bf16(tmem, fh, wl, idesc_feat, 1);
                tc::umma_bf16(tmem, fl, wh, idesc_feat, 1);
                tc::umma_commit(&bars[0]);
            }
            tc::mbar_wait(&bars[0], ph0); ph0 ^= 1;
            tc::tc_fence_after();
            // ---- phase 1: pre-activation row -> LayerNorm + ReLU -> bf16 hi/lo A operand in TMEM
            {
                const int mlp = half;                       // warps 0-3: key MLP, warps 4-7: value MLP
                float xr[128];
#pragma unroll
                for (int c4 = 0; c4 < 4; c4++) {
                    float v[32];
                    tc::tmem_ld32(tmem + lane_base + mlp * 128 + c4 * 32, v);
#pragma unroll
                    for (int q8 = 0; q8 < 8; q8++) {
                        const int c = c4 * 32 + q8 * 4;
                        const float4 p = ld4(sPs + trow * PS_LD + mlp * 128 + c);
                        const float4 rr = ldg4(a.R + (size_t)eji * 256 + mlp * 128 + c);
                        xr[c] = v[q8 * 4] + p.x + rr.x; xr[c + 1] = v[q8 * 4 + 1] + p.y + rr.y;
                        xr[c + 2] = v[q8 * 4 + 2] + p.z + rr.z; xr[c + 3] = v[q8 * 4 + 3] + p.w + rr.w;
                    }
                }
                float s1 = 0.f;
#pragma unroll
                for (int c = 0; c < 128; c++) s1 += xr[c];
                const float mu = s1 * (1.0f / 128.0f);
                float s2 = 0.f;
#pragma unroll
                for (int c = 0; c < 128; c++) { const float dd = xr[c] - mu; s2 = fmaf(dd, dd, s2); }
                const float rstd = rsqrtf(s2 * (1.0f / 128.0f) + 1e-5f);
                const float nmr = -mu * rstd;
                tc::tc_fence_before();
                __syncthreads();                             // every pre_* column has been read: safe to overwrite pre_k with hid_v
                tc::tc_fence_after();
                const float* gam = sLn + mlp * 256;
                const float* bet = gam + 128;
                const uint32_t hid = tmem + lane_base + (mlp == 0 ? 384 : 0);
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {             // channels [32*q4, 32*q4+32)
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int c = q4 * 32 + 4 * i;
                        const float4 g4 = ld4(gam + c), b4 = ld4(bet + c);
                        const float y0 = fmaxf(fmaf(fmaf(xr[c], rstd, nmr), g4.x, b4.x), 0.f);
                        const float y1 = fmaxf(fmaf(fmaf(xr[c + 1], rstd, nmr), g4.y, b4.y), 0.f);
                        const float y2 = fmaxf(fmaf(fmaf(xr[c + 2], rstd, nmr), g4.z, b4.z), 0.f);
                        const float y3 = fmaxf(fmaf(fmaf(xr[c + 3], rstd, nmr), g4.w, b4.w), 0.f);
                        __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
                        tc::split_bf16(y0, h0, l0); tc::split_bf16(y1, h1, l1); tc::split_bf16(y2, h2, l2); tc::split_bf16(y3, h3, l3);
                        hi[2 * i] = tc::pack_bf16(h0, h1); hi[2 * i + 1] = tc::pack_bf16(h2, h3);
                        lo[2 * i] = tc::pack_bf16(l0, l1); lo[2 * i + 1] = tc::pack_bf16(l2, l3);
                    }
                    tc::tmem_st16(hid + q4 * 16, hi);
                    tc::tmem_st16(hid + 64 + q4 * 16, lo);
                }
                tc::tmem_st_wait();
            }
            tc::tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc::tc_fence_after();
#pragma unroll
                for (int mlp = 0; mlp < 2; mlp++) {
                    const uint32_t hid = tmem + (mlp == 0 ? 384 : 0);
                    const uint32_t dcol = tmem + (mlp == 0 ? 256 : 128);
                    uint32_t acc = 0;
#pragma unroll
                    for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                        const uint32_t abase = hid + (combo == 2 ? 64 : 0);
                        const uint32_t bbase = sW_u32 + (mlp * 2 + (combo == 1 ? 1 : 0)) * W_TILE;
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint64_t bd = tc::umma_desc_sw128(bbase + (ks >> 2) * 16384 + (ks & 3) * 32);
                            tc::umma_bf16_ts(dcol, abase + ks * 8, bd, idesc_w2, acc);
                            acc = 1;
                        }
                    }
                    tc::umma_commit(&bars[1 + mlp]);
                }
            }
            // ---- phase 2: logits, segment softmax (rows = lanes), alpha-weighted sum of values
            float alpha[8];
            {
                tc::mbar_wait(&bars[1], ph1); ph1 ^= 1;
                tc::tc_fence_after();
                float kk[64];
                {
                    float v[32];
                    tc::tmem_ld32(tmem + lane_base + 256 + half * 64, v);
#pragma unroll
                    for (int i = 0; i < 32; i++) kk[i] = v[i];
                    tc::tmem_ld32(tmem + lane_base + 256 + half * 64 + 32, v);
#pragma unroll
                    for (int i = 0; i < 32; i++) kk[32 + i] = v[i];
                }
                const float* qrow = sQ + wq * 128 + half * 64;
                const float* b2 = sB2 + half * 64;
#pragma unroll
                for (int h = 0; h < 8; h++) {
                    float s = 0.f;
#pragma unroll
                    for (int dd = 0; dd < 8; dd++) s = fmaf(qrow[h * 8 + dd], kk[h * 8 + dd] + b2[h * 8 + dd], s);
                    const float lg = rowvalid ? s * kInvSqrtD : -INFINITY;
                    const float m = warp_max(lg);
                    const float e = rowvalid ? expf(lg - m) : 0.f;
                    const float tot = warp_sum(e);
                    alpha[h] = e / tot;
                }
            }
            {
                tc::mbar_wait(&bars[2], ph2); ph2 ^= 1;
                tc::tc_fence_after();
                const float* b2 = sB2 + 128 + half * 64;
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    float v[32];
                    tc::tmem_ld32(tmem + lane_base + 128 + half * 64 + ch * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = rowvalid ? alpha[ch * 4 + (i >> 3)] * v[i] : 0.f;
                    const float o = transpose_reduce32(v, lane);
                    if (segvalid) {
                        const int c = half * 64 + ch * 32 + lane;
                        a.hb[(size_t)eji * 128 + c] += o + b2[ch * 32 + lane];     // sum(alpha) = 1; residual (uni_denoiser.py:285)
                    }
                }
            }
            tc::tc_fence_before();
            __syncthreads();                                 // TMEM columns, sFeat and sQ are reused by the next tile
        }
    }
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}

// R[e] = smear(|x_dst - x_src|) @ Wrji  for every bond edge (the r_ji slice of the triplet MLPs' first Linear)
__global__ void __launch_bounds__(256) trip_r_kernel(PlanDev d, const float* __restrict__ x, const float* __restrict__ wrji,
                                                     float* __restrict__ R) {
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= d.Eb) return;
    const int s = d.esrc_node[e], t = d.edst_node[e];
    const float d0 = x[(size_t)t * 3] - x[(size_t)s * 3], d1 = x[(size_t)t * 3 + 1] - x[(size_t)s * 3 + 1], d2 = x[(size_t)t * 3 + 2] - x[(size_t)s * 3 + 2];
    const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float mine = lane < 20 ? smear_val(dist, lane) : 0.f;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
#pragma unroll
    for (int gg = 0; gg < 20; gg++) {
        const float sg = __shfl_sync(PG_FULL, mine, gg);
        a0 = f4fma(sg, ldg4(wrji + gg * 256 + lane * 4), a0);
        a1 = f4fma(sg, ldg4(wrji + gg * 256 + 128 + lane * 4), a1);
    }
    st4(R + (size_t)e * 256 + lane * 4, a0);
    st4(R + (size_t)e * 256 + 128 + lane * 4, a1);
}
}  // namespace

size_t pg_trip_tc_smem(int maxn) { return (size_t)SM_FIXED + (size_t)(maxn - 1) * (PS_LD + 20) * sizeof(float) + 1024; }

int pg_launch_trip_tc(const TripTcArgs& a, int num_sms, cudaStream_t s) {
    if (a.d.Nl <= 0) return PG_OK;
    const size_t smem = pg_trip_tc_smem(a.maxn);
    if (smem > 227 * 1024) { pg_set_error("trip_tc: shared memory budget exceeded"); return PG_ELIMIT; }
    static size_t cur = 0;
    if (smem > cur) { PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
    trip_r_kernel<<<(unsigned)((a.d.Eb + 7) / 8), 256, 0, s>>>(a.d, a.x, a.wrji, a.R);
    PG_LAUNCH_CHECK();
    const unsigned grid = (unsigned)std::min<long long>(a.d.Nl, num_sms);
    trip_tc_kernel<<<grid, 256, smem, s>>>(a);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

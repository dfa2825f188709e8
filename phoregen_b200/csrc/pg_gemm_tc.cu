// tcgen05 / TMEM version of the K = 128 contractions:  C[M, nt*64] = pro(A)[M,128] @ W[nt*64,128]^T + b.
//
// Precision: "bf16x3" — both operands are split x = hi + lo into two bf16 values and the product is evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (three bf16-in / fp32-accumulate MMAs per K slice), which
// keeps ~16 mantissa bits (rel. error ~2^-16) — needed for the 1e-3 / 1e-4 parity bar, which plain bf16
// (2^-9) does not meet through six residual layers.
//
// Persistent, warp-specialised CTA (one per SM) walking 128-row tiles:
//   warps 0-7  epilogue      : accumulator (TMEM, double buffered) -> registers -> bias / residual / ReLU -> global;
//                              warp w owns lane quarter w & 3 and column half w >> 2 of every 64-column block
//   warps 8-15 A producer    : fused prologue (sum of two inputs, or gather-add + LayerNorm + ReLU), bf16 hi/lo split,
//                              canonical K-major 128B-swizzled UMMA layout, double buffered (next tile while MMAs run)
//   warp  16   B loader      : weights are stored pre-swizzled, one 32 KB image per 64-column block -> a single
//                              cp.async.bulk (TMA unit) per block into a 2-stage ring, mbarrier transaction counts
//   warp  17   MMA issuer    : 24 tcgen05.mma (M=128, N=64, K=16) per block, tcgen05.commit releases smem / signals epilogue
#include <algorithm>
#include "pg_gemm.h"
#include "pg_tc.cuh"

namespace {
constexpr int TM = 128, TN = 64;
constexpr int A_KBLK = TM * 128;             // one 64-wide K block of the A tile: 128 rows x 128 B
constexpr int A_BUF = 4 * A_KBLK;            // (hi,lo) x 2 K blocks = 64 KB
constexpr int B_KBLK = TN * 128;             // 8 KB
constexpr int B_BUF = 4 * B_KBLK;            // (hi,lo) x 2 K blocks = 32 KB  (== one pre-swizzled weight image)
constexpr int NB = 2;                        // B ring stages
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;   // per epilogue warp: staging tile [32 rows x 32 cols], 16-byte chunks XOR-swizzled by row
constexpr int SMEM_TOTAL = 2 * A_BUF + NB * B_BUF + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int PROD_WARPS = 8;
constexpr int NTHREADS = (EPI_WARPS + PROD_WARPS + 2) * 32;
constexpr int PROD_WARP0 = EPI_WARPS, LOAD_WARP = EPI_WARPS + PROD_WARPS, MMA_WARP = LOAD_WARP + 1;

enum { A_FULL = 0, A_EMPTY = 2, B_FULL = 4, B_EMPTY = 4 + NB, ACC_FULL = 4 + 2 * NB, ACC_EMPTY = 6 + 2 * NB, NBARS = 8 + 2 * NB };

template <int PRO>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(GemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment; offset arithmetic keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * A_BUF;
    float* sEpi = (float*)(sB + NB * B_BUF);
    uint64_t* bars = (uint64_t*)(sB + NB * B_BUF + EPI_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(bars + NBARS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the role branches and the MMA issuer's descriptor arithmetic stay on the uniform datapath
    const long long n_mtiles = (a.M + TM - 1) / TM;

    if (warp == MMA_WARP) tc::tmem_alloc<2 * TN>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&bars[A_FULL + i], PROD_WARPS * 32); tc::mbar_init(&bars[A_EMPTY + i], 1);
            tc::mbar_init(&bars[ACC_FULL + i], 1); tc::mbar_init(&bars[ACC_EMPTY + i], EPI_WARPS * 32);
        }
        for (int i = 0; i < NB; i++) { tc::mbar_init(&bars[B_FULL + i], 1); tc::mbar_init(&bars[B_EMPTY + i], 1); }
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EPI_WARPS) {
        // ================= epilogue =================
        // TMEM row (thread = row) -> per-warp smem tile -> row-contiguous global stores (4 rows x 128 B per instruction)
        long long cnt = 0;
        float* stg = sEpi + warp * 32 * 32;
        const int wq = warp & 3, hh = warp >> 2;
        const int rsub = lane >> 3, cj = lane & 7, csub = cj * 4;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            const long long mw = mt * TM + wq * 32;                 // first row of this warp
            for (int nt = 0; nt < a.ntiles; nt++, cnt++) {
                const int ab = cnt & 1;
                tc::mbar_wait(&bars[ACC_FULL + ab], (cnt >> 1) & 1);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32_nowait(tmem_base + ((uint32_t)(wq * 32) << 16) + ab * TN + hh * 32, v);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars[ACC_EMPTY + ab]);            // accumulator is in registers: the MMA warp may reuse it
                const int c0 = nt * TN + hh * 32 + csub;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    st4(stg + lane * 32 + ((q ^ (lane & 7)) << 2), make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                                             __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])));
                __syncwarp();
                float4 bb = make_float4(0, 0, 0, 0);
                if (a.bias) bb = ldg4(a.bias + c0);
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    const int r = rr * 4 + rsub;
                    const long long m = mw + r;
                    if (m < a.M) {
                        float4 o = f4add(ld4(stg + r * 32 + ((cj ^ (r & 7)) << 2)), bb);
                        if (a.resid) o = f4add(o, ld4(a.resid + m * a.ldr + c0));
                        if (a.relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                        if (a.csplit > 0 && c0 >= a.csplit) st4(a.C2 + m * a.ldc2 + (c0 - a.csplit), o);
                        else st4(a.C + m * a.ldc + c0, o);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp < LOAD_WARP) {
        // ================= A producer: fused prologue, bf16 hi/lo split, swizzled K-major layout =================
        const int pw = warp - PROD_WARP0;
        float4 g4 = make_float4(1, 1, 1, 1), b4 = make_float4(0, 0, 0, 0);
        constexpr bool LN = PRO == PRO_LNRELU || PRO == PRO_LNRELU_MF;
    if (LN) { g4 = ldg4(a.ln_g + lane * 4); b4 = ldg4(a.ln_b + lane * 4); }
        long long it = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x, it++) {
            const int buf = it & 1;
            tc::mbar_wait(&bars[A_EMPTY + buf], ((it >> 1) & 1) ^ 1);   // MMAs that read this buffer two tiles ago are done
            uint8_t* dstA = sA + buf * A_BUF;
            const long long m0 = mt * TM;
            // rows pw, pw+PROD_WARPS, ...: loads of a whole batch are issued before any is consumed (enough bytes in flight per SM
            // to cover HBM latency: Little's law needs ~28 KB at 23 B/clk/SM)
            constexpr int BATCH = (PRO == PRO_PLAIN) ? 16 : 8;
            for (int rb = 0; rb < TM / PROD_WARPS; rb += BATCH) {
                float4 v[BATCH], w[BATCH];
                // gather indices of the whole batch first: one round trip instead of one per row (the warp issues in order)
                long long gi[BATCH];
                if (LN) {
#pragma unroll
                    for (int i = 0; i < BATCH; i++) {
                        const long long m = min(m0 + pw + (long long)(rb + i) * PROD_WARPS, a.M - 1);
                        gi[i] = (a.A2 && a.gidx) ? (long long)__ldg(a.gidx + m) : m;
                    }
                }
#pragma unroll
                for (int i = 0; i < BATCH; i++) {
                    const long long m = m0 + pw + (rb + i) * PROD_WARPS;
                    const long long mc = min(m, a.M - 1);           // rows past the end re-read the last row and are zeroed below
                    v[i] = ld4(a.A + mc * a.lda + lane * 4);
                    w[i] = make_float4(0, 0, 0, 0);
                    if (PRO == PRO_SUM2) w[i] = ld4(a.A2 + mc * a.lda2 + lane * 4);
                    if (LN && a.A2) w[i] = ld4(a.A2 + gi[i] * a.lda2 + lane * 4);
                    if (m >= a.M) { v[i] = make_float4(0, 0, 0, 0); w[i] = v[i]; }
                }
                if (PRO != PRO_PLAIN) {
#pragma unroll
                    for (int i = 0; i < BATCH; i++) v[i] = f4add(v[i], w[i]);
                }
                if (LN) {
                    // LayerNorm + ReLU of the whole batch: the BATCH rows walk through the shuffle butterflies together (ILP)
                    float s1[BATCH], s2[BATCH];
                    if (PRO != PRO_LNRELU_MF) {        // mean-free inputs skip the first butterfly and the subtraction
#pragma unroll
                        for (int i = 0; i < BATCH; i++) s1[i] = (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                            for (int i = 0; i < BATCH; i++) s1[i] += __shfl_xor_sync(PG_FULL, s1[i], o);
                    }
#pragma unroll
                    for (int i = 0; i < BATCH; i++) {
                        if (PRO != PRO_LNRELU_MF) {
                            const float mu = s1[i] * (1.0f / 128.0f);
                            v[i] = make_float4(v[i].x - mu, v[i].y - mu, v[i].z - mu, v[i].w - mu);
                        }
                        s2[i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, v[i].w * v[i].w)));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int i = 0; i < BATCH; i++) s2[i] += __shfl_xor_sync(PG_FULL, s2[i], o);
#pragma unroll
                    for (int i = 0; i < BATCH; i++) {
                        const float rstd = rsqrtf(s2[i] * (1.0f / 128.0f) + 1e-5f);
                        v[i] = make_float4(fmaxf(fmaf(v[i].x * rstd, g4.x, b4.x), 0.f), fmaxf(fmaf(v[i].y * rstd, g4.y, b4.y), 0.f),
                                           fmaxf(fmaf(v[i].z * rstd, g4.z, b4.z), 0.f), fmaxf(fmaf(v[i].w * rstd, g4.w, b4.w), 0.f));
                        if (m0 + pw + (rb + i) * PROD_WARPS >= a.M) v[i] = make_float4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int i = 0; i < BATCH; i++) {
                    const int r = pw + (rb + i) * PROD_WARPS;
                    const float4 x = v[i];
                    uint32_t h0, l0, h1, l1;
                    tc::split_pair_trunc(x.x, x.y, h0, l0);
                    tc::split_pair_trunc(x.z, x.w, h1, l1);
                    // lane covers k = 4*lane .. 4*lane+3: K block kb = lane/16, 16-byte chunk j = (lane%16)/2, half = lane%2
                    const int kb = lane >> 4, j = (lane & 15) >> 1, half = lane & 1;
                    const uint32_t off = kb * A_KBLK + tc::sw128_chunk(r, j) + half * 8;
                    *reinterpret_cast<uint2*>(dstA + off) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(dstA + 2 * A_KBLK + off) = make_uint2(l0, l1);
                }
            }
            tc::fence_proxy_async_smem();                               // generic-proxy writes -> async proxy (tensor core reads)
            tc::mbar_arrive(&bars[A_FULL + buf]);
        }
    } else if (warp == LOAD_WARP) {
        // ================= B loader: one bulk copy of a pre-swizzled 32 KB weight image per 64-column block =================
        long long cnt = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            for (int nt = 0; nt < a.ntiles; nt++, cnt++) {
                const int s = cnt % NB;
                tc::mbar_wait_wd(&bars[B_EMPTY + s], ((cnt / NB) & 1) ^ 1);
                if (lane == 0) {
                    tc::mbar_arrive_expect_tx(&bars[B_FULL + s], B_BUF);
                    tc::bulk_copy_g2s(sB + s * B_BUF, reinterpret_cast<const uint8_t*>(a.Wbf) + (size_t)nt * B_BUF, B_BUF, &bars[B_FULL + s]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = tc::umma_idesc_bf16(TM, TN);
        const uint32_t sA_u32 = tc::smem_u32(sA), sB_u32 = tc::smem_u32(sB);
        long long cnt = 0, it = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x, it++) {
            const int buf = it & 1;
            tc::mbar_wait_wd(&bars[A_FULL + buf], (it >> 1) & 1);
            for (int nt = 0; nt < a.ntiles; nt++, cnt++) {
                const int s = cnt % NB, ab = cnt & 1;
                tc::mbar_wait_wd(&bars[B_FULL + s], (cnt / NB) & 1);
                tc::mbar_wait_wd(&bars[ACC_EMPTY + ab], ((cnt >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                // warp-collective issue (pg_tc.cuh): descriptor arithmetic on the uniform datapath, one elected lane issues
                uint32_t acc = 0;
#pragma unroll
                for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                    const uint32_t abase = sA_u32 + buf * A_BUF + (combo == 2 ? 2 * A_KBLK : 0);
                    const uint32_t bbase = sB_u32 + s * B_BUF + (combo == 1 ? 2 * B_KBLK : 0);
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {           // 4 x K=16 slices (32 B) inside the 128-byte swizzle row
                            const uint64_t ad = tc::umma_desc_sw128(abase + kb * A_KBLK + k * 32);
                            const uint64_t bd = tc::umma_desc_sw128(bbase + kb * B_KBLK + k * 32);
                            tc::umma_bf16_w(tmem_base + ab * TN, ad, bd, idesc, acc);
                            acc = 1;
                        }
                    }
                }
                tc::umma_commit_w(&bars[ACC_FULL + ab]);          // accumulator ready for the epilogue
                tc::umma_commit_w(&bars[B_EMPTY + s]);            // weight stage may be refilled
                if (nt == a.ntiles - 1) tc::umma_commit_w(&bars[A_EMPTY + buf]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc::tc_fence_after(); tc::tmem_dealloc<2 * TN>(tmem_base); }
}
}  // namespace

int pg_launch_gemm_tc(const GemmArgs& a, int pro, cudaStream_t stream) {
    if (a.M <= 0) return PG_OK;
    if (!a.Wbf) { pg_set_error("tcgen05 GEMM needs the bf16 hi/lo weights"); return PG_EINVAL; }
    static int sms = 0;
    if (!sms) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<PRO_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<PRO_SUM2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<PRO_LNRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<PRO_LNRELU_MF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        int dev = 0;
        PG_CUDA_CHECK(cudaGetDevice(&dev));
        PG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long n_mtiles = (a.M + TM - 1) / TM;
    const unsigned grid = (unsigned)std::min<long long>(n_mtiles, sms);
    GemmArgs b = a;
    b.ntiles = a.ntiles * 2;      // callers count 128-column blocks; this kernel walks 64-column blocks
    switch (pro) {
        case PRO_PLAIN: gemm_tc_kernel<PRO_PLAIN><<<grid, NTHREADS, SMEM_TOTAL, stream>>>(b); break;
        case PRO_SUM2: gemm_tc_kernel<PRO_SUM2><<<grid, NTHREADS, SMEM_TOTAL, stream>>>(b); break;
        case PRO_LNRELU: gemm_tc_kernel<PRO_LNRELU><<<grid, NTHREADS, SMEM_TOTAL, stream>>>(b); break;
        case PRO_LNRELU_MF: gemm_tc_kernel<PRO_LNRELU_MF><<<grid, NTHREADS, SMEM_TOTAL, stream>>>(b); break;
        default: pg_set_error("bad gemm prologue"); return PG_EINVAL;
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

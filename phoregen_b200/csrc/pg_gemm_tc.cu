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
// Hand-offs from a warp role to the MMA issuer are ONE mbarrier arrival per warp (fence by every lane, __syncwarp, lane 0
// arrives): an mbarrier.arrive is a shared-memory atomic, and 256 of them on one barrier serialise.
//
// TSA = true (plain prologue, i.e. the wide edge GEMMs): the A operand lives in TMEM instead of shared memory.
// With both operands in shared memory an M128 N64 K16 MMA reads 4 KB of A and 2 KB of B: 144 KB per 64-column block, which
// together with the weight refill (32 KB) and the epilogue's staging tile (64 KB) is ~1,900 cycles of the 128 B/clk shared
// memory pipe per block against 24 x 46 = 1,100 cycles of MMA issue; measured (tools/gemm_bw.py): 896 output columns per
// row ran at 3.4 TB/s written while a plain fill of the same buffer reaches 7.4 TB/s.  The producer warps therefore read
// their rows thread-per-row (256-bit loads), split to bf16 hi/lo in registers and tcgen05.st the pairs into TMEM columns
// (lane = row, two K elements per 32-bit column; 2 tiles x (64 hi + 64 lo) columns next to the 2 x 64 accumulator
// columns); the MMAs take A from TMEM and the freed 128 KB of shared memory deepen the weight ring from 2 to 4 stages.
#include <algorithm>
#include <cstdlib>
#include "pg_gemm.h"
#include "pg_tc.cuh"

namespace {
constexpr int TM = 128, TN = 64;
constexpr int A_KBLK = TM * 128;             // one 64-wide K block of the A tile: 128 rows x 128 B
constexpr int A_BUF = 4 * A_KBLK;            // (hi,lo) x 2 K blocks = 64 KB
constexpr int B_KBLK = TN * 128;             // 8 KB
constexpr int B_BUF = 4 * B_KBLK;            // (hi,lo) x 2 K blocks = 32 KB  (== one pre-swizzled weight image)
constexpr int NB_SS = 2, NB_TS = 4;          // B ring stages (A in shared memory / A in TMEM)
constexpr int NBMAX = 4;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;   // per epilogue warp: staging tile [32 rows x 32 cols], 16-byte chunks XOR-swizzled by row
constexpr int SMEM_SS = 2 * A_BUF + NB_SS * B_BUF + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int SMEM_TS = NB_TS * B_BUF + EPI_BYTES + 1024 + 256;
constexpr int TM_ACC = 0, TM_A = 2 * TN;     // TMEM columns (TSA): accumulators 2 x 64, then per A buffer 64 hi + 64 lo
constexpr int PROD_WARPS = 8;
constexpr int NTHREADS = (EPI_WARPS + PROD_WARPS + 2) * 32;
constexpr int PROD_WARP0 = EPI_WARPS, LOAD_WARP = EPI_WARPS + PROD_WARPS, MMA_WARP = LOAD_WARP + 1;

enum { A_FULL = 0, A_EMPTY = 2, B_FULL = 4, B_EMPTY = 4 + NBMAX, ACC_FULL = 4 + 2 * NBMAX, ACC_EMPTY = 6 + 2 * NBMAX, NBARS = 8 + 2 * NBMAX };

// 256-bit global load (sm_100: LDG.E.256): one full 32-byte sector per thread
__device__ __forceinline__ void ld8(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

template <int PRO, bool TSA>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(GemmArgs a) {
    constexpr int NB = TSA ? NB_TS : NB_SS;
    constexpr int TMEM_COLS = TSA ? 512 : 2 * TN;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment; offset arithmetic keeps the shared address space (LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + (TSA ? 0 : 2 * A_BUF);
    float* sEpi = (float*)(sB + NB * B_BUF);
    uint64_t* bars = (uint64_t*)(sB + NB * B_BUF + EPI_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(bars + NBARS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the role branches and the MMA issuer's descriptor arithmetic stay on the uniform datapath
    const long long n_mtiles = (a.M + TM - 1) / TM;

    if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&bars[A_FULL + i], PROD_WARPS); tc::mbar_init(&bars[A_EMPTY + i], 1);
            tc::mbar_init(&bars[ACC_FULL + i], 1); tc::mbar_init(&bars[ACC_EMPTY + i], EPI_WARPS);
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
                const int c0 = nt * TN + hh * 32 + csub;
                float4 bb = make_float4(0, 0, 0, 0);
                if (a.bias) bb = ldg4(a.bias + c0);                 // requested before the wait: its latency is not on the block's critical path
                tc::mbar_wait(&bars[ACC_FULL + ab], (cnt >> 1) & 1);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32_nowait(tmem_base + ((uint32_t)(wq * 32) << 16) + TM_ACC + ab * TN + hh * 32, v);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars[ACC_EMPTY + ab]);   // accumulator is in registers: the MMA warp may reuse it (one arrival per warp, see the header)
#pragma unroll
                for (int q = 0; q < 8; q++)
                    st4(stg + lane * 32 + ((q ^ (lane & 7)) << 2), make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                                             __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])));
                __syncwarp();
                // The output buffer, its row pitch and the fused-epilogue flavour are per-block facts: decide them once, then a
                // row costs LDS + 4 FADD + STG + one pointer step.  (With the decisions inside the row loop the epilogue warps
                // executed 53 instructions per 512-byte store and were the limiter of the wide edge GEMMs: ncu, source view.)
                const bool second = a.csplit > 0 && c0 >= a.csplit;
                const long long ldo = second ? a.ldc2 : a.ldc;
                float* prow = (second ? a.C2 + (c0 - a.csplit) : a.C + c0) + (mw + rsub) * ldo;
                const int rows_left = (int)min((long long)32, a.M - mw) - rsub;        // rows rr * 4 + rsub < 32 of this warp that exist
                const float* srow = stg + rsub * 32;
                if (!a.resid && !a.relu) {
                    float4 o[8];
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) o[rr] = ld4(srow + rr * 128 + ((cj ^ ((rr * 4 + rsub) & 7)) << 2));     // all eight reads in flight
#pragma unroll
                    for (int rr = 0; rr < 8; rr++)
                        if (rr * 4 < rows_left) st4(prow + (long long)(rr * 4) * ldo, f4add(o[rr], bb));
                } else {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) {
                        const int r = rr * 4 + rsub;
                        if (rr * 4 < rows_left) {
                            float4 o = f4add(ld4(srow + rr * 128 + ((cj ^ (r & 7)) << 2)), bb);
                            if (a.resid) o = f4add(o, ld4(a.resid + (mw + r) * a.ldr + c0));
                            if (a.relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                            st4(prow + (long long)(rr * 4) * ldo, o);
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (TSA && warp < LOAD_WARP) {
        // ================= A producer, operand in TMEM: thread = (row, K half); 256-bit loads, split, tcgen05.st =================
        const int pw = warp - PROD_WARP0;
        const int q = pw & 3, kh = pw >> 2;                 // a warp reaches the TMEM lanes of quarter (warp % 4); PROD_WARP0 % 4 == 0
        long long it = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x, it++) {
            const int buf = it & 1;
            const long long m = mt * TM + q * 32 + lane;
            const long long mc = min(m, a.M - 1);           // rows past the end re-read the last row and are zeroed below
            const float* src = a.A + mc * a.lda + kh * 64;
            float v[64];
#pragma unroll
            for (int i = 0; i < 8; i++) ld8(src + i * 8, v + i * 8);
            if (m >= a.M) {
#pragma unroll
                for (int i = 0; i < 64; i++) v[i] = 0.f;
            }
            tc::mbar_wait_sleep(&bars[A_EMPTY + buf], ((it >> 1) & 1) ^ 1);   // MMAs that read this TMEM buffer two tiles ago are done
            tc::tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TM_A + buf * 128 + kh * 32;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; i++) tc::split_pair_trunc(v[c * 32 + 2 * i], v[c * 32 + 2 * i + 1], hi[i], lo[i]);
                tc::tmem_st16(ta + c * 16, hi);
                tc::tmem_st16(ta + 64 + c * 16, lo);
            }
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[A_FULL + buf]);
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
            tc::mbar_wait_sleep(&bars[A_EMPTY + buf], ((it >> 1) & 1) ^ 1);   // MMAs that read this buffer two tiles ago are done
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
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[A_FULL + buf]);
        }
    } else if (warp == LOAD_WARP) {
        // ================= B loader: one bulk copy of a pre-swizzled 32 KB weight image per 64-column block =================
        long long cnt = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            for (int nt = 0; nt < a.ntiles; nt++, cnt++) {
                const int s = cnt % NB;
                tc::mbar_wait_wd(&bars[B_EMPTY + s], ((cnt / NB) & 1) ^ 1);
                if (lane == 0) {
#if defined(PG_GEMM_EXP) && (PG_GEMM_EXP & 1)      // timing experiment (wrong results): no weight traffic after the first tile
                    if (mt != blockIdx.x) tc::mbar_arrive(&bars[B_FULL + s]); else
#endif
                    {
                    tc::mbar_arrive_expect_tx(&bars[B_FULL + s], B_BUF);
                    tc::bulk_copy_g2s(sB + s * B_BUF, reinterpret_cast<const uint8_t*>(a.Wbf) + (size_t)nt * B_BUF, B_BUF, &bars[B_FULL + s]);
                    }
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
#if defined(PG_GEMM_EXP) && (PG_GEMM_EXP & 4)      // timing experiment (wrong results): 3 MMAs per block instead of 24
#define PG_EXP_KB 1
#define PG_EXP_K 1
#else
#define PG_EXP_KB 2
#define PG_EXP_K 4
#endif
                for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                    const uint32_t abase = sA_u32 + buf * A_BUF + (combo == 2 ? 2 * A_KBLK : 0);
                    const uint32_t atmem = tmem_base + TM_A + buf * 128 + (combo == 2 ? 64 : 0);
                    const uint32_t bbase = sB_u32 + s * B_BUF + (combo == 1 ? 2 * B_KBLK : 0);
#pragma unroll
                    for (int kb = 0; kb < PG_EXP_KB; kb++) {
#pragma unroll
                        for (int k = 0; k < PG_EXP_K; k++) {    // 4 x K=16 slices (32 B) inside the 128-byte swizzle row
                            const uint64_t bd = tc::umma_desc_sw128(bbase + kb * B_KBLK + k * 32);
                            if (TSA) {
                                tc::umma_bf16_ts_w(tmem_base + TM_ACC + ab * TN, atmem + (kb * 4 + k) * 8, bd, idesc, acc);
                            } else {
                                const uint64_t ad = tc::umma_desc_sw128(abase + kb * A_KBLK + k * 32);
                                tc::umma_bf16_w(tmem_base + ab * TN, ad, bd, idesc, acc);
                            }
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
    if (warp == MMA_WARP) { tc::tc_fence_after(); tc::tmem_dealloc<TMEM_COLS>(tmem_base); }
}
}  // namespace

template <int PRO, bool TSA>
static int launch(const GemmArgs& b, unsigned grid, cudaStream_t stream) {
    constexpr int smem = TSA ? SMEM_TS : SMEM_SS;
    static bool configured = false;
    if (!configured) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<PRO, TSA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    gemm_tc_kernel<PRO, TSA><<<grid, NTHREADS, smem, stream>>>(b);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

int pg_launch_gemm_tc(const GemmArgs& a, int pro, cudaStream_t stream) {
    if (a.M <= 0) return PG_OK;
    if (!a.Wbf) { pg_set_error("tcgen05 GEMM needs the bf16 hi/lo weights"); return PG_EINVAL; }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        PG_CUDA_CHECK(cudaGetDevice(&dev));
        PG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    static const char* ss_env = getenv("PG_GEMM_SS");
    static const bool force_ss = ss_env && *ss_env;      // A/B switch: A operand in shared memory for every prologue
    const long long n_mtiles = (a.M + TM - 1) / TM;
    const unsigned grid = (unsigned)std::min<long long>(n_mtiles, sms);
    GemmArgs b = a;
    b.ntiles = a.ntiles * 2;      // callers count 128-column blocks; this kernel walks 64-column blocks
    // the TMEM-operand producer reads its rows with 256-bit loads
    const bool ts = !force_ss && ((uintptr_t)a.A % 32 == 0) && a.lda % 8 == 0;
    switch (pro) {
        case PRO_PLAIN: return ts ? launch<PRO_PLAIN, true>(b, grid, stream) : launch<PRO_PLAIN, false>(b, grid, stream);
        case PRO_SUM2: return launch<PRO_SUM2, false>(b, grid, stream);        // small node GEMMs only
        case PRO_LNRELU: return launch<PRO_LNRELU, false>(b, grid, stream);
        case PRO_LNRELU_MF: return launch<PRO_LNRELU_MF, false>(b, grid, stream);
        default: pg_set_error("bad gemm prologue"); return PG_EINVAL;
    }
}

// NodeUpdateLayer / PosUpdateLayer over the complete ligand bond graph (uni_denoiser.py:284,294) on the 5th-gen
// tensor cores.  Same tile shape and thread mapping as the triplet kernel (pg_trip_tc.cu):
//
//   tile   = 4 ligand atoms (segments) x 32 rows; TMEM lane = triplet row = incoming bond edge j -> i (n-1 <= 32 rows)
//   thread = (row, 32-channel quarter): 16 row warps, warp w -> lane quarter w & 3 (segment), channel quarter w >> 2
//
//   1. pre-activation of the key / value MLP = dst-node partial + src-node partial + per-edge partial (three global
//      rows, the first Linear was applied by the node / edge GEMMs), LayerNorm + ReLU thread-locally (statistics of
//      the four channel quarters meet in shared memory), bf16 hi/lo split, tcgen05.st -> A operand in TMEM
//   2. second Linear: 24 tcgen05.mma (M128 N128 K16; value MLP of the position layer: N16), B = W2 bf16 hi/lo resident
//      in shared memory (128B swizzle), fp32 accumulators in TMEM
//   3. logits = q . k per head (thread-local), segment softmax across the 32 lanes, then
//        node layer: alpha-weighted sum of the value rows (butterfly transpose-reduce)          -> out[i, 128]
//        pos  layer: sum_h alpha_h v_h per row, times rel_x, summed over the rows, mean over heads -> out[i, 3]
//
// Software pipeline across tiles, as in the triplet kernel:  LN-k(t) | epilogue(t-1) | LN-v(t) | logits(t)  on the row
// warps while the tensor pipe runs W2k(t) | W2v(t).  TMEM: hid_k [0,128) out_k [128,256) hid_v [256,384) out_v [384,512).
// MULTI: batches with segments longer than one lane quarter (n-1 > 32).  A segment then takes C = ceil((n-1)/32) quarters
// of the same tile (4/C atoms per tile, tile table in the plan); the softmax statistics and the per-quarter outputs of
// the C row warps of a (segment, channel quarter) meet in shared memory.
#include <algorithm>
#include "pg_attn.h"
#include "pg_tc.cuh"

namespace {
constexpr int W_TILE = 32768;               // one [128 x 128] bf16 matrix in two 128B-swizzled K blocks
constexpr int SM_W = 4 * W_TILE;            // (k,v) x (hi,lo)
constexpr int SM_STAT = 128 * 16 * 4;       // [2 mlp][4 quarters][128 rows] partial second moments of the LayerNorm (half used)
constexpr int SM_RED = 2 * 4 * 4 * 4 * 4;   // pos layer: [tile parity][segment][quarter][4] partial coordinate updates
constexpr int SM_EX = 4 * 4 * 8 * 4;        // MULTI: [channel quarter][lane quarter][max 4 | sum 4] softmax statistics of a chunk
constexpr int SM_OX = 4 * 128 * 4;          // MULTI, node layer: [lane quarter][128] per-chunk outputs
constexpr int XP_LD = 36;                   // padded row stride (floats) of a warp's transpose tile [32 rows][32 channels]
constexpr int SM_XP = 16 * 32 * XP_LD * 4;  // one tile per row warp
constexpr int SM_TOTAL = SM_W + SM_XP + SM_STAT + SM_RED + SM_EX + SM_OX + 5 * 128 * 4 /*ln + b2v*/ + 128 /*barriers*/ + 1024 /*alignment*/;
constexpr float kInvSqrtD = 0.35355339059327373f;
constexpr int ROW_WARPS = 16;
constexpr int MMA_WARP = ROW_WARPS;
constexpr int NTHREADS = (ROW_WARPS + 4) * 32;
constexpr int ROW_THREADS = ROW_WARPS * 32;
enum { B_HIDK = 0, B_HIDV, B_OUTK, B_OUTV, B_COUNT };

struct SegInfo {
    bool valid;
    int n, il, v, ctx0;
    int r0, nch, wq0;       // first row of this quarter's chunk, chunks of the segment, lane quarter of its first chunk
    long long e0;
};
template <bool MULTI>
__device__ __forceinline__ SegInfo seg_info(const PlanDev& d, long long tile, int wq) {
    SegInfo s;
    s.valid = false; s.n = 2; s.il = 0; s.v = 0; s.ctx0 = 0; s.e0 = 0; s.r0 = 0; s.nch = 1; s.wq0 = wq;
    if (!MULTI) {
        const long long u = tile * 4 + wq;
        if (u < d.Nl) {
            const int g = d.lig_graph[u];
            s.n = d.g_n[g]; s.il = (int)(u - d.lig_off[g]);
            s.ctx0 = d.ctx_off[g] + d.g_p[g];
            s.v = s.ctx0 + s.il;
            s.e0 = d.eoff[g] + (long long)s.il * (s.n - 1);
            s.valid = s.n - 1 >= 1 && s.n - 1 <= 32;
        }
    } else if (tile < d.nbt) {
        const int g = d.btile_graph[tile];
        const int n = d.g_n[g], nch = pg_bond_chunks(n), apt = pg_bond_atoms_per_tile(n);
        const int slot = wq / nch;
        const int il = (int)(tile - d.btile_off[g]) * apt + slot;
        if (slot < apt && il < n) {
            s.n = n; s.il = il; s.nch = nch; s.wq0 = slot * nch; s.r0 = (wq - s.wq0) * 32;
            s.ctx0 = d.ctx_off[g] + d.g_p[g];
            s.v = s.ctx0 + il;
            s.e0 = d.eoff[g] + (long long)il * (n - 1);
            s.valid = true;
        }
    }
    return s;
}

// KF16: the key MLP's second Linear as a single-pass fp16 contraction (8 MMAs; pg_trip_tc.cu explains the error budget)
template <int POS, bool MULTI, bool KF16>
__global__ void __launch_bounds__(NTHREADS, 1) bond_tc_kernel(BondTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    float* sXp = (float*)(sW + SM_W);
    float* sStat = sXp + SM_XP / 4;
    float* sRed = sStat + SM_STAT / 4;
    float* sEx = sRed + SM_RED / 4;
    float* sOx = sEx + SM_EX / 4;
    float* sLn = sOx + SM_OX / 4;                   // gk, bk, gv, bv
    float* sB2 = sLn + 4 * 128;                     // b2v (128 or 16)
    uint64_t* bars = (uint64_t*)(sB2 + 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + B_COUNT);
    volatile int* sProg = (volatile int*)(tmem_slot + 1);   // tiles whose value LayerNorm is done (pacing of the prefetch warps)
    const PlanDev& d = a.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the role branches and the MMA issuer's descriptor arithmetic stay on the uniform datapath
    const int wq = warp & 3;
    constexpr int NV = POS ? 16 : 128;              // outputs of the value MLP's second Linear

    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bars[B_HIDK], ROW_THREADS); tc::mbar_init(&bars[B_HIDV], ROW_THREADS);
        tc::mbar_init(&bars[B_OUTK], 1); tc::mbar_init(&bars[B_OUTV], 1);
        tc::fence_barrier_init();
        *sProg = 0;
    }
    // ---- resident weights: 16-byte chunks [mat 4][n][chunk 16] -> 128B-swizzled K-major tiles
    for (int idx = tid; idx < 4 * 128 * 16; idx += NTHREADS) {
        const int mat = idx >> 11, n = (idx >> 4) & 127, c = idx & 15;
        const bool isv = mat >> 1;
        if ((isv && n >= NV) || (KF16 && mat == 1)) continue;
        const uint16_t* src = (isv ? a.w2v_bf : (KF16 ? a.w2k_h : a.w2k_bf)) + ((size_t)(mat & 1) * (isv ? NV : 128) + n) * 128 + c * 8;
        const uint32_t dst = tc::smem_u32(sW) + mat * W_TILE + (c >> 3) * 16384 + tc::sw128_chunk(n, c & 7);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    if (tid < 128) {
        sLn[tid] = a.w.lnk_g[tid]; sLn[128 + tid] = a.w.lnk_bf[tid]; sLn[256 + tid] = a.w.lnv_g[tid]; sLn[384 + tid] = a.w.lnv_bf[tid];
        if (tid < NV) sB2[tid] = a.w.b2v[tid];
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    constexpr uint32_t C_HIDK = 0, C_OUTK = 128, C_HIDV = 256, C_OUTV = 384;
    // a CTA walks a CONTIGUOUS range of tiles [tile_begin, ntiles): consecutive tiles are atoms of the same molecule, whose
    // node partial rows are then re-read from this SM's L1 instead of L2
    const long long ntiles_all = MULTI ? d.nbt : (d.Nl + 3) / 4;
    const long long tile_begin = ntiles_all * blockIdx.x / gridDim.x, ntiles = ntiles_all * (blockIdx.x + 1) / gridDim.x;

    if (warp >= MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == MMA_WARP) {
            // ================= MMA issue =================
            constexpr uint32_t idesc_k = tc::umma_idesc_bf16(128, 128);
            constexpr uint32_t idesc_v = tc::umma_idesc_bf16(128, NV);
            const uint32_t sW_u32 = tc::smem_u32(sW);
            uint32_t ph = 0;
            for (long long tile = tile_begin; tile < ntiles; tile++, ph ^= 1) {
#pragma unroll
                for (int mlp = 0; mlp < 2; mlp++) {
                    tc::mbar_wait_wd(&bars[mlp == 0 ? B_HIDK : B_HIDV], ph);
                    tc::tc_fence_after();
                    // warp-collective issue (pg_tc.cuh): descriptor arithmetic on the uniform datapath, one elected lane issues
                    if (KF16 && mlp == 0) {
                        constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A, B = F16
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint64_t bd = tc::umma_desc_sw128(sW_u32 + (ks >> 2) * 16384 + (ks & 3) * 32);
                            tc::umma_bf16_ts_w(tmem + C_OUTK, tmem + C_HIDK + ks * 8, bd, idesc16, ks > 0);
                        }
                        tc::umma_commit_w(&bars[B_OUTK]);
                    } else {
                        const uint32_t hid = tmem + (mlp == 0 ? C_HIDK : C_HIDV);
                        const uint32_t dcol = tmem + (mlp == 0 ? C_OUTK : C_OUTV);
                        uint32_t acc = 0;
#pragma unroll
                        for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                            const uint32_t abase = hid + (combo == 2 ? 64 : 0);
                            const uint32_t bbase = sW_u32 + (mlp * 2 + (combo == 1 ? 1 : 0)) * W_TILE;
#pragma unroll
                            for (int ks = 0; ks < 8; ks++) {
                                const uint64_t bd = tc::umma_desc_sw128(bbase + (ks >> 2) * 16384 + (ks & 3) * 32);
                                tc::umma_bf16_ts_w(dcol, abase + ks * 8, bd, mlp == 0 ? idesc_k : idesc_v, acc);
                                acc = 1;
                            }
                        }
                        tc::umma_commit_w(&bars[mlp == 0 ? B_OUTK : B_OUTV]);
                    }
                    __syncwarp();
                }
            }
        } else {
            // ================= idle warps: pull the next tile's per-edge rows (the only HBM stream) into L2 =================
            const int pt = tid - (MMA_WARP + 1) * 32;      // 0..95
            uint32_t ph = 0;
            int k = 0;
            for (long long tile = tile_begin; tile < ntiles; tile++, ph ^= 1, k++) {
                const long long nt = tile + 1;
                if (nt < ntiles) {
                    for (int s4 = 0; s4 < 4; s4++) {
                        const SegInfo sg = seg_info<MULTI>(d, nt, s4);
                        if (!sg.valid) continue;
                        for (int r = sg.r0 + pt; r < min(sg.n - 1, sg.r0 + 32); r += 96) {
                            const float* row = a.B + (size_t)(sg.e0 + r) * a.ldb + a.b_k;
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"((a.b_v - a.b_k + 128) * 4) : "memory");
                        }
                    }
                }
                // pace: one tile ahead of the row warps.  A monotonic counter, not an mbarrier parity wait: a prefetch warp
                // that falls two phases behind would wait for a phase that never completes.
                while (*sProg < k + 1) __nanosleep(100);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= row warps: thread = (row, channel quarter) =================
        const int cq = warp >> 2;
        const bool fold[2] = {a.w.fold[0] > 0.5f, a.w.fold[1] > 0.5f};
        float al[4] = {0.f, 0.f, 0.f, 0.f};
        bool prev_valid = false;
        int prev_v = 0;
        float rel0 = 0.f, rel1 = 0.f, rel2 = 0.f;   // pos layer: x[dst] - x[src] of the previous tile's row
        // pre-activation slice (dst node + src node + bond edge partials) -> LayerNorm + ReLU -> bf16 hi/lo A operand
        // Coalesced gather: 8 lanes cover the 128-byte slice of one row (4 rows per instruction); the rows meet in a per-warp
        // shared tile and are read back row-per-lane, the layout of TMEM.  (A row-per-lane global read would cost 32 L1
        // wavefronts per instruction instead of 4.)  The per-edge rows - the only HBM stream - are requested one phase
        // ahead with cp.async straight into the tile; padded rows re-read row 0 and are masked by alpha = 0.
        const int sr = lane >> 3, ch = (lane & 7) * 4;
        float* xp = sXp + warp * (32 * XP_LD);
        auto request_edge_rows = [&](int mlp, const SegInfo& sg) {
            const int nrow = sg.valid ? sg.n - 1 : 0;
            const float* pb = a.B + (mlp == 0 ? a.b_k : a.b_v) + cq * 32 + ch;
            const uint32_t dst = tc::smem_u32(xp + sr * XP_LD + ch);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = (sg.r0 + i * 4 + sr) < nrow ? (sg.r0 + i * 4 + sr) : 0;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 4 * XP_LD * 4), "l"(pb + (size_t)(sg.e0 + row) * a.ldb) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // pre-activation slice (dst node + src node + bond edge partials) -> LayerNorm + ReLU -> bf16 hi/lo A operand;
        // `nsg` / `nmlp`: the phase whose edge rows are requested as soon as this one has left the shared tile
        auto layer_norm = [&](int mlp, const SegInfo& sg, uint32_t hid, bool has_next, int nmlp, const SegInfo& nsg) {
            {
                const int c0 = cq * 32 + ch;
                const int nrow = sg.valid ? sg.n - 1 : 0;
                const float4 d4 = ldg4(a.nc.A + (size_t)sg.v * a.nc.lda + (mlp == 0 ? a.nc.dst_k : a.nc.dst_v) + c0);
                const float* ps = a.nc.A + (mlp == 0 ? a.nc.src_k : a.nc.src_v) + c0;
                float4 u[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = (sg.r0 + i * 4 + sr) < nrow ? (sg.r0 + i * 4 + sr) : 0;
                    u[i] = ldg4(ps + (size_t)(sg.ctx0 + row + (row >= sg.il)) * a.nc.lda);
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float* pxy = xp + (i * 4 + sr) * XP_LD + ch;
                    st4(pxy, f4add(f4add(ld4(pxy), u[i]), d4));
                }
                __syncwarp();
            }
            const float* xrow = xp + lane * XP_LD;
            float2 x2[16];
            // the pre-activation is mean-free (every first-Linear block has its channel mean removed at pack time,
            // weights._center_first_linears): the LayerNorm only needs the second moment
            float2 s2 = make_float2(0.f, 0.f), s2b = s2;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float4 e4 = ld4(xrow + 2 * i);
                x2[i] = make_float2(e4.x, e4.y);
                x2[i + 1] = make_float2(e4.z, e4.w);
                s2 = tc::fma2(x2[i], x2[i], s2); s2b = tc::fma2(x2[i + 1], x2[i + 1], s2b);
            }
            __syncwarp();
            if (has_next) request_edge_rows(nmlp, nsg);
            s2 = tc::add2(s2, s2b);
            // combine with the other three channel quarters of the same row (warps w +- 4k, same lane).  The barrier also
            // orders this lane quarter's reads of the previous tile's accumulators before the hid columns are rewritten.
            // quarter-major layout [mlp][quarter][row]: consecutive lanes touch consecutive 8-byte words (no bank conflicts)
            float* st = sStat + (size_t)(mlp * 4) * 128 + wq * 32 + lane;
            st[cq * 128] = s2.x + s2.y;
            asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
            const float rstd = rsqrtf(((st[0] + st[128]) + (st[256] + st[384])) * (1.0f / 128.0f) + 1e-5f);
            const float2 rs2 = make_float2(rstd, rstd);
            const float* gam = sLn + mlp * 256 + cq * 32;
            const float* bet = gam + 128;
            if (KF16 && mlp == 0) {
                // key MLP: one fp16 value per activation (its output only feeds the softmax logits; see pg_trip_tc.cu)
                uint32_t hh[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 b4 = ld4(bet + 2 * i);
                    float2 y0, y1;
                    if (fold[0]) {
                        y0 = tc::fma2(x2[i], rs2, make_float2(b4.x, b4.y)); y1 = tc::fma2(x2[i + 1], rs2, make_float2(b4.z, b4.w));
                    } else {
                        const float4 g4 = ld4(gam + 2 * i);
                        y0 = tc::fma2(tc::mul2(x2[i], rs2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
                        y1 = tc::fma2(tc::mul2(x2[i + 1], rs2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
                    }
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hh[i]) : "f"(y0.y), "f"(y0.x));
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hh[i + 1]) : "f"(y1.y), "f"(y1.x));
                }
                tc::tmem_st16(hid + lane_base + cq * 16, hh);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars[B_HIDK]);
                return;
            }
            uint32_t hi[16], lo[16];
            if (fold[mlp]) {
                // gamma > 0 everywhere: it lives in the columns of W2, only beta / gamma is added here (half the broadcast loads)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 b4 = ld4(bet + 2 * i);
                    tc::split_pair_relu(tc::fma2(x2[i], rs2, make_float2(b4.x, b4.y)), hi[i], lo[i]);
                    tc::split_pair_relu(tc::fma2(x2[i + 1], rs2, make_float2(b4.z, b4.w)), hi[i + 1], lo[i + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 g4 = ld4(gam + 2 * i), b4 = ld4(bet + 2 * i);
                    float2 y0 = tc::fma2(tc::mul2(x2[i], rs2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
                    float2 y1 = tc::fma2(tc::mul2(x2[i + 1], rs2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
                    tc::split_pair_relu(y0, hi[i], lo[i]);
                    tc::split_pair_relu(y1, hi[i + 1], lo[i + 1]);
                }
            }
            tc::tmem_st16(hid + lane_base + cq * 16, hi);
            tc::tmem_st16(hid + lane_base + 64 + cq * 16, lo);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&bars[mlp == 0 ? B_HIDK : B_HIDV]);
        };
        int prev_nch = 1, prev_wq0 = 0, prev_r0 = 0;
        auto epilogue = [&](uint32_t parity) {
            tc::mbar_wait(&bars[B_OUTV], parity);
            tc::tc_fence_after();
            if (POS == 0) {
                uint32_t vu[32];
                tc::tmem_ld32_nowait(tmem + C_OUTV + lane_base + cq * 32, vu);
                tc::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {                                           // alpha is 0 on padded rows
                    const float2 pr = tc::mul2(make_float2(al[i >> 3], al[i >> 3]), make_float2(__uint_as_float(vu[i]), __uint_as_float(vu[i + 1])));
                    v[i] = pr.x; v[i + 1] = pr.y;
                }
                const float o = transpose_reduce32(v, lane);
                const int c = cq * 32 + lane;
                if (!MULTI) {
                    if (prev_valid) a.out[(size_t)prev_v * 128 + c] = o + sB2[c];          // sum(alpha) = 1
                } else {
                    // the chunks of a segment sit in consecutive lane quarters: their partial outputs meet in shared memory
                    sOx[wq * 128 + c] = o;
                    asm volatile("bar.sync %0, 128;" ::"r"(7 + cq) : "memory");
                    if (prev_valid && prev_r0 == 0) {
                        float t = sOx[prev_wq0 * 128 + c];
                        for (int k = 1; k < prev_nch; k++) t += sOx[(prev_wq0 + k) * 128 + c];
                        a.out[(size_t)prev_v * 128 + c] = t + sB2[c];
                    }
                }
            } else {
                float vh[4];
                tc::tmem_ld4(tmem + C_OUTV + lane_base + cq * 4, vh);
                float c = 0.f;
#pragma unroll
                for (int h = 0; h < 4; h++) c = fmaf(al[h], vh[h] + sB2[cq * 4 + h], c);     // alpha is 0 on padded rows
                float p0 = c * rel0, p1 = c * rel1, p2 = c * rel2;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    p0 += __shfl_xor_sync(PG_FULL, p0, o); p1 += __shfl_xor_sync(PG_FULL, p1, o); p2 += __shfl_xor_sync(PG_FULL, p2, o);
                }
                float* red = sRed + (parity & 1) * 64;      // double-buffered by tile parity (readers of other quarters, MULTI)
                float* rd = red + (wq * 4 + cq) * 4;
                if (lane == 0) { rd[0] = p0; rd[1] = p1; rd[2] = p2; }
                if (!MULTI) {
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
                    if (cq == 0 && lane < 3 && prev_valid) {
                        const float* r4 = red + wq * 16 + lane;
                        a.out[(size_t)prev_v * 3 + lane] = ((r4[0] + r4[4]) + (r4[8] + r4[12])) * (1.0f / 16.0f);
                    }
                } else {
                    asm volatile("bar.sync 11, %0;" ::"n"(ROW_THREADS) : "memory");
                    if (cq == 0 && lane < 3 && prev_valid && prev_r0 == 0) {
                        float t = 0.f;
                        for (int k = 0; k < prev_nch; k++) {
                            const float* r4 = red + (prev_wq0 + k) * 16 + lane;
                            t += (r4[0] + r4[4]) + (r4[8] + r4[12]);
                        }
                        a.out[(size_t)prev_v * 3 + lane] = t * (1.0f / 16.0f);
                    }
                }
            }
        };
        uint32_t ph = 0;
        int tiles_done = 0;
        bool any = false;
        SegInfo sg = seg_info<MULTI>(d, tile_begin, wq);
        request_edge_rows(0, sg);
        for (long long tile = tile_begin; tile < ntiles; tile++, ph ^= 1) {
            const bool rowvalid = sg.valid && sg.r0 + lane < sg.n - 1;
            const int r = rowvalid ? sg.r0 + lane : 0;
            const bool more = tile + 1 < ntiles;
            const SegInfo nsg = seg_info<MULTI>(d, more ? tile + 1 : tile, wq);
            // ---- key MLP
            layer_norm(0, sg, tmem + C_HIDK, true, 1, sg);
            // ---- value epilogue of the previous tile (its W2v MMA ran during that tile's logits and the LayerNorm above)
            if (any) epilogue(ph ^ 1);
            // ---- value MLP
            layer_norm(1, sg, tmem + C_HIDV, more, 0, nsg);
            if (warp == 0 && lane == 0) *sProg = ++tiles_done;
            // ---- logits of this thread's 4 heads, segment softmax across the 32 lanes (rows) of the warp
            {
                // the key bias b2k shifts all logits of a (segment, head) by the same q . b: softmax-invariant, dropped
                const float* qrow = a.q + (size_t)sg.v * 128 + cq * 32;
                float4 qv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) qv[i] = ldg4(qrow + i * 4);
                tc::mbar_wait(&bars[B_OUTK], ph);
                tc::tc_fence_after();
                uint32_t vv[32];
                tc::tmem_ld32_nowait(tmem + C_OUTK + lane_base + cq * 32, vv);
                tc::tmem_ld_wait();
                constexpr float kScale = kInvSqrtD * 1.4426950408889634f;      // logits in log2 units -> ex2 directly
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int o = h * 8;
                    const float4 qa = qv[2 * h], qb = qv[2 * h + 1];
                    float2 s0 = tc::mul2(make_float2(qa.x, qa.y), make_float2(__uint_as_float(vv[o]), __uint_as_float(vv[o + 1])));
                    float2 s1 = tc::mul2(make_float2(qb.x, qb.y), make_float2(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5])));
                    s0 = tc::fma2(make_float2(qa.z, qa.w), make_float2(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3])), s0);
                    s1 = tc::fma2(make_float2(qb.z, qb.w), make_float2(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7])), s1);
                    s0 = tc::add2(s0, s1);
                    al[h] = rowvalid ? (s0.x + s0.y) * kScale : -INFINITY;
                }
                // segment softmax across the 32 lanes: max and sum on the REDUX unit
                float mx[4], ssum[4];
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    mx[h] = tc::warp_max_redux(al[h]);
                    al[h] = rowvalid ? tc::ex2_approx(al[h] - mx[h]) : 0.f;
                }
#pragma unroll
                for (int h = 0; h < 4; h++) ssum[h] = tc::warp_sum01_redux(al[h]);   // every lane takes part in the reduction
                if (!MULTI) {
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] = rowvalid ? al[h] * tc::rcp_approx(ssum[h]) : 0.f;
                } else {
                    // (max, sum) of every chunk of the segment -> global max M and sum L; alpha = p 2^(m - M) / L
                    float* ex = sEx + (cq * 4 + wq) * 8;
                    if (lane == 0) { st4(ex, make_float4(mx[0], mx[1], mx[2], mx[3])); st4(ex + 4, make_float4(ssum[0], ssum[1], ssum[2], ssum[3])); }
                    asm volatile("bar.sync %0, 128;" ::"r"(7 + cq) : "memory");
                    float M[4] = {mx[0], mx[1], mx[2], mx[3]}, L[4] = {0.f, 0.f, 0.f, 0.f};
                    const float* e0 = sEx + (cq * 4 + sg.wq0) * 8;
                    for (int k = 0; k < sg.nch; k++) {
                        const float4 m4 = ld4(e0 + k * 8);
                        M[0] = fmaxf(M[0], m4.x); M[1] = fmaxf(M[1], m4.y); M[2] = fmaxf(M[2], m4.z); M[3] = fmaxf(M[3], m4.w);
                    }
                    for (int k = 0; k < sg.nch; k++) {
                        const float4 m4 = ld4(e0 + k * 8), l4 = ld4(e0 + k * 8 + 4);
                        L[0] = fmaf(l4.x, tc::ex2_approx(m4.x - M[0]), L[0]); L[1] = fmaf(l4.y, tc::ex2_approx(m4.y - M[1]), L[1]);
                        L[2] = fmaf(l4.z, tc::ex2_approx(m4.z - M[2]), L[2]); L[3] = fmaf(l4.w, tc::ex2_approx(m4.w - M[3]), L[3]);
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] = rowvalid ? al[h] * __fdividef(tc::ex2_approx(mx[h] - M[h]), L[h]) : 0.f;
                }
            }
            prev_valid = sg.valid; prev_v = sg.v; prev_nch = sg.nch; prev_wq0 = sg.wq0; prev_r0 = sg.r0;
            if (POS) {
                const int sj = sg.ctx0 + r + (r >= sg.il);
                rel0 = a.x[(size_t)sg.v * 3] - a.x[(size_t)sj * 3];                        // rel_x = x[dst] - x[src]
                rel1 = a.x[(size_t)sg.v * 3 + 1] - a.x[(size_t)sj * 3 + 1];
                rel2 = a.x[(size_t)sg.v * 3 + 2] - a.x[(size_t)sj * 3 + 2];
            }
            any = true;
            sg = nsg;
        }
        if (any) epilogue(ph ^ 1);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}
}  // namespace

int pg_launch_bond_tc(const BondTcArgs& a, int pos, int num_sms, cudaStream_t s) {
    if (a.d.Nl <= 0) return PG_OK;
    static bool init = false;
    if (!init) {
#define PG_BOND_ATTR(P, M, K) PG_CUDA_CHECK(cudaFuncSetAttribute(bond_tc_kernel<P, M, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL))
        PG_BOND_ATTR(0, false, false); PG_BOND_ATTR(1, false, false); PG_BOND_ATTR(0, true, false); PG_BOND_ATTR(1, true, false);
        PG_BOND_ATTR(0, false, true); PG_BOND_ATTR(1, false, true); PG_BOND_ATTR(0, true, true); PG_BOND_ATTR(1, true, true);
#undef PG_BOND_ATTR
        init = true;
    }
    // Which instantiation an atom runs on depends on its molecule alone (n-1 <= 32: atoms packed four to a tile across the
    // batch; longer segments: the per-graph tile table, which lists only those molecules), so results do not depend on
    // the batch composition.  A mixed batch launches both.
    const bool have_single = a.d.min_n - 1 <= PG_BOND_TC_SINGLE_CHUNK_ROWS, have_multi = a.d.nbt > 0;
#define PG_BOND_GO(P, M, K) bond_tc_kernel<P, M, K><<<grid, NTHREADS, SM_TOTAL, s>>>(a)
    if (have_single) {
        const unsigned grid = (unsigned)std::min<long long>((a.d.Nl + 3) / 4, num_sms);
        if (a.key_bf16x3) { if (pos == 0) PG_BOND_GO(0, false, false); else PG_BOND_GO(1, false, false); }
        else { if (pos == 0) PG_BOND_GO(0, false, true); else PG_BOND_GO(1, false, true); }
    }
    if (have_multi) {
        const unsigned grid = (unsigned)std::min<long long>(a.d.nbt, num_sms);
        if (a.key_bf16x3) { if (pos == 0) PG_BOND_GO(0, true, false); else PG_BOND_GO(1, true, false); }
        else { if (pos == 0) PG_BOND_GO(0, true, true); else PG_BOND_GO(1, true, true); }
    }
#undef PG_BOND_GO
    PG_LAUNCH_CHECK();
    return PG_OK;
}

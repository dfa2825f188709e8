// NodeUpdateLayer / PosUpdateLayer over the joint ligand + pharmacophore kNN graph (uni_denoiser.py:264-281,291) on the
// 5th-gen tensor cores.  Tile = 4 destination nodes (segments) x 32 neighbour rows; thread = (row, 32-channel quarter)
// exactly as in pg_trip_tc.cu / pg_bond_tc.cu.
//
// The key and the value MLP run as two passes of the same kernel (PASS 0 / 1): shared memory holds one second-Linear
// weight (bf16 hi/lo, 64 KB), one edge-feature table (48 KB) and the edge-feature operand (48 KB) at a time.
//   pre-activation = dst-node partial + src-node partial (gathered) + Table[edge type] . feat(edge)
//     feat = 20 Gaussian smearings of the distance, 1, 3 direction dot products (uni_denoiser.py:373-387); the one-hot
//     edge type (x) feat outer product is a K = 96 operand, so the table product is 18 tcgen05.mma (M128 N128 K16, bf16x3)
//     into TMEM; three otherwise idle warps compute the features of the next tile while the row warps work.
//   LayerNorm + ReLU thread-locally -> bf16 hi/lo A operand in TMEM -> second Linear (24 tcgen05.mma).
//   PASS 0: logits = q . k per head, segment softmax over the 32 lanes, alpha * e_w -> global scratch [Ek,16] (+ its
//           per-head sums [N,16]).
//   PASS 1: node layer  out[i] = sum_r alpha'_r v_r + b2v * sum_r alpha'_r           (butterfly transpose-reduce)
//           pos  layer  out[i] = mean_h sum_r alpha'_rh (v_rh + b2v_h) rel_x_r
// Row-warp order per tile: LayerNorm(t) | logits or epilogue (t-1), so the W2 MMA of tile t and the table MMA of tile
// t+1 run under the previous tile's post-processing.  TMEM: pre [0,128) hid [128,256) out [256,384) | [384,512).
#include <algorithm>
#include "pg_attn.h"
#include "pg_tc.cuh"

// Optional phase tracing (debug builds only, -DPG_TRIP_TRACE): cycle stamps of CTA 0, key pass
#ifdef PG_TRIP_TRACE
__device__ long long g_knn_trace[2 * 64 * 16];
extern "C" int pg_debug_knn_trace(long long* h_out) { return cudaMemcpyFromSymbol(h_out, g_knn_trace, sizeof(g_knn_trace)) == cudaSuccess ? 0 : -2; }
#define KTRACE(role, slot) do { if (PASS == 0 && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 16) && tcount < 64) g_knn_trace[((role) * 64 + tcount) * 16 + (slot)] = clock64(); } while (0)
// per row warp: 0 loop top, 1 PRE wait done, 2 HID arrival, 3 post-processing done
__device__ long long g_knn_wtrace[16 * 64 * 4];
extern "C" int pg_debug_knn_wtrace(long long* h_out) { return cudaMemcpyFromSymbol(h_out, g_knn_wtrace, sizeof(g_knn_wtrace)) == cudaSuccess ? 0 : -2; }
#define KWTRACE(slot) do { if (PASS == 0 && blockIdx.x == 0 && lane == 0 && tcount < 64) g_knn_wtrace[(warp * 64 + tcount) * 4 + (slot)] = clock64(); } while (0)
#else
#define KTRACE(role, slot) do {} while (0)
#define KWTRACE(slot) do {} while (0)
#endif

namespace {
constexpr int W_TILE = 32768;               // one [128 x 128] bf16 matrix in two 128B-swizzled K blocks
constexpr int SM_W = 2 * W_TILE;            // hi, lo
constexpr int KF = 96;                      // 4 edge types x 24 features
constexpr int TAB_PART = KF * 128 * 2;      // 24 KB: [k-step 6][n/8][kc 2][n%8][8 bf16]
constexpr int SM_TAB = 2 * TAB_PART;
constexpr int FEAT_PART = 128 * KF * 2;     // 24 KB: [k-step 6][row/8][kc 2][row%8][8 bf16]
constexpr int SM_FEAT = 2 * FEAT_PART;
constexpr int XP_LD = 36;
constexpr int SM_XP = 16 * 16 * XP_LD * 4;  // per row warp: transpose tile of 16 rows x 32 channels (two rounds per tile)
constexpr int SM_STAT = 128 * 4 * 2 * 4;    // [4 quarters][128 rows] partial second moments of the LayerNorm (half used)
constexpr int SM_RED = 4 * 4 * 4 * 4;
constexpr int SM_Q = 16 * 2 * 32 * 4;       // per row warp: two 128-byte slots for the query slice of the pending / current tile
constexpr int SM_TOTAL = SM_W + SM_TAB + SM_FEAT + SM_XP + SM_STAT + SM_RED + SM_Q + 3 * 128 * 4 + 128 + 1024;
constexpr float kInvSqrtD = 0.35355339059327373f;
constexpr int ROW_WARPS = 16;
constexpr int MMA_WARP = ROW_WARPS;
constexpr int NTHREADS = (ROW_WARPS + 4) * 32;
constexpr int ROW_THREADS = ROW_WARPS * 32;
enum { B_FEAT = 0, B_PRE, B_HID, B_OUT, B_COUNT };

struct KSeg {
    bool valid, dl, inr;
    int v, R, ctx0, gp;
    long long e0;
};
// graph of the node that lane quarter wq handles in `tile` (first level of the dependent index loads; -1 past the end)
__device__ __forceinline__ int kseg_graph(const PlanDev& d, long long tile, int wq) {
    const long long v = tile * 4 + wq;
    return v < d.N ? d.node_graph[v] : -1;
}
// second level, given the graph (the row warps fetch the graph one tile earlier, so neither level's latency is exposed)
__device__ __forceinline__ KSeg kseg_of_graph(const PlanDev& d, long long tile, int wq, int g) {
    KSeg s;
    const long long v = tile * 4 + wq;
    s.valid = false; s.dl = false; s.inr = v < d.N; s.v = 0; s.R = 0; s.ctx0 = 0; s.gp = 0; s.e0 = 0;
    if (v < d.N) {
        const int ng = d.g_n[g] + d.g_p[g];
        s.v = (int)v; s.ctx0 = d.ctx_off[g]; s.gp = d.g_p[g];
        s.R = min(PG_KNN, ng - 1);
        s.e0 = d.koff[g] + (long long)(s.v - s.ctx0) * s.R;
        s.dl = (s.v - s.ctx0) >= s.gp;
        s.valid = s.R >= 1;
    }
    return s;
}
__device__ __forceinline__ KSeg kseg(const PlanDev& d, long long tile, int wq) {
    KSeg s;
    const long long v = tile * 4 + wq;
    s.valid = false; s.dl = false; s.inr = v < d.N; s.v = 0; s.R = 0; s.ctx0 = 0; s.gp = 0; s.e0 = 0;
    if (v < d.N) {
        const int g = d.node_graph[v];
        const int ng = d.g_n[g] + d.g_p[g];
        s.v = (int)v; s.ctx0 = d.ctx_off[g]; s.gp = d.g_p[g];
        s.R = min(PG_KNN, ng - 1);
        s.e0 = d.koff[g] + (long long)(s.v - s.ctx0) * s.R;
        s.dl = (s.v - s.ctx0) >= s.gp;
        s.valid = s.R >= 1;
    }
    return s;
}

// KF16 (key pass only): the second Linear as a single-pass fp16 contraction (8 MMAs; pg_trip_tc.cu explains the error budget)
template <int PASS, int POS, bool KF16>
__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(KnnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sTab = sW + SM_W;
    uint8_t* sFeat = sTab + SM_TAB;
    float* sXp = (float*)(sFeat + SM_FEAT);
    float* sStat = sXp + SM_XP / 4;
    float* sRed = sStat + SM_STAT / 4;
    float* sQ = sRed + SM_RED / 4;
    float* sLn = sQ + SM_Q / 4;                 // gamma, beta of this pass's LayerNorm
    float* sB2 = sLn + 2 * 128;                     // b2v (128 or 16), value pass only
    uint64_t* bars = (uint64_t*)(sB2 + 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + B_COUNT);
    const PlanDev& d = a.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the role branches and the MMA issuer's descriptor arithmetic stay on the uniform datapath
    const int wq = warp & 3;
    constexpr int NOUT = (PASS == 1 && POS) ? 16 : 128;     // outputs of this pass's second Linear

    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bars[B_FEAT], 128); tc::mbar_init(&bars[B_PRE], 1);
        tc::mbar_init(&bars[B_HID], ROW_THREADS); tc::mbar_init(&bars[B_OUT], 1);
        tc::fence_barrier_init();
    }
    // ---- resident operands of this pass
    {
        const uint16_t* w2 = PASS == 0 ? (KF16 ? a.w2k_h : a.w2k_bf) : a.w2v_bf;
        for (int idx = tid; idx < 2 * 128 * 16; idx += NTHREADS) {        // [part 2][n][chunk 16] -> 128B-swizzled K-major
            const int part = idx >> 11, n = (idx >> 4) & 127, c = idx & 15;
            if (n >= NOUT || (KF16 && part == 1)) continue;
            const uint16_t* src = w2 + ((size_t)part * NOUT + n) * 128 + c * 8;
            const uint32_t dst = tc::smem_u32(sW) + part * W_TILE + (c >> 3) * 16384 + tc::sw128_chunk(n, c & 7);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        const uint16_t* tb = PASS == 0 ? a.tabk_bf : a.tabv_bf;
        for (int idx = tid; idx < 2 * 128 * 12; idx += NTHREADS) {        // [part 2][n 128][chunk 12] -> no-swizzle K-major
            const int part = idx / (128 * 12), rem = idx - part * (128 * 12), n = rem / 12, c = rem - n * 12;
            const uint16_t* src = tb + ((size_t)part * 128 + n) * KF + c * 8;
            const uint32_t dst = tc::smem_u32(sTab) + part * TAB_PART + (c >> 1) * 4096 + (n >> 3) * 256 + (c & 1) * 128 + (n & 7) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        for (int i = tid; i < SM_FEAT / 16; i += NTHREADS) reinterpret_cast<uint4*>(sFeat)[i] = make_uint4(0, 0, 0, 0);
        if (tid < 128) {
            sLn[tid] = (PASS == 0 ? a.w.lnk_g : a.w.lnv_g)[tid]; sLn[128 + tid] = (PASS == 0 ? a.w.lnk_bf : a.w.lnv_bf)[tid];
            if (PASS == 1 && tid < NOUT) sB2[tid] = a.w.b2v[tid];
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    constexpr uint32_t C_PRE = 0, C_HID = 128, C_OUT = 256;     // out: two buffers of 128 columns
    // a CTA walks a CONTIGUOUS range of tiles: consecutive tiles are nodes of the same molecule, so the gathered partial
    // rows, coordinates and neighbour lists of a molecule are re-read from this SM's L1 instead of L2
    const long long ntiles_all = (d.N + 3) / 4;
    const long long tile_begin = ntiles_all * blockIdx.x / gridDim.x, ntiles = ntiles_all * (blockIdx.x + 1) / gridDim.x;   // [tile_begin, ntiles)

    if (warp >= MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ================= auxiliary warpgroup: edge features of the next tile (one row per thread) + MMA issue (warp 16) =======
        // feature row r = tid - 512: segment r / 32, neighbour row r % 32 -> bf16 hi/lo one-hot-typed operand row
        const int r = tid - MMA_WARP * 32, frow = r & 31;
        int prev_type = -1;
        // The inputs of a feature row hang off a four-level chain of dependent loads (node -> graph -> offsets -> neighbour
        // index -> coordinates / direction vectors).  Fetched inside the per-tile feature step the chain was fully exposed
        // (~4,000 cycles per tile, the critical path of the whole kernel: table MMA(t) -> features(t+1) -> table MMA(t+1));
        // here every level runs one tile ahead of the next one, so a tile's feature step finds its inputs in registers.
        struct FIn { float xd0, xd1, xd2, xs0, xs1, xs2, cs0, cs1, cs2, cd0, cd1, cd2; int type; };    // type -1: padded row
        const long long fstep = 1;
        auto graph_at = [&](long long t) -> int { return kseg_graph(d, t < ntiles ? t : tile_begin, r >> 5); };
        auto seg_at = [&](long long t, int g) -> KSeg { return kseg_of_graph(d, t < ntiles ? t : tile_begin, r >> 5, g); };
        auto src_at = [&](long long t, const KSeg& sg) -> int { return (t < ntiles && sg.valid && frow < sg.R) ? a.knn_src[sg.e0 + frow] : -1; };
        auto load_in = [&](int s, const KSeg& sg) -> FIn {
            FIn f;
            f.xd0 = f.xd1 = f.xd2 = f.xs0 = f.xs1 = f.xs2 = f.cs0 = f.cs1 = f.cs2 = f.cd0 = f.cd1 = f.cd2 = 0.f;
            f.type = -1;
            if (s >= 0) {
                f.xd0 = a.x[(size_t)sg.v * 3]; f.xd1 = a.x[(size_t)sg.v * 3 + 1]; f.xd2 = a.x[(size_t)sg.v * 3 + 2];
                f.xs0 = a.x[(size_t)s * 3]; f.xs1 = a.x[(size_t)s * 3 + 1]; f.xs2 = a.x[(size_t)s * 3 + 2];
                const float* c1 = a.comb + (size_t)s * 3;                             // vec_1 = comb[src]
                const float* c2 = a.comb + (size_t)sg.v * 3;                          // vec_2 = comb[dst]
                f.cs0 = c1[0]; f.cs1 = c1[1]; f.cs2 = c1[2]; f.cd0 = c2[0]; f.cd1 = c2[1]; f.cd2 = c2[2];
                const bool sl = (s - sg.ctx0) >= sg.gp;
                f.type = sl ? (sg.dl ? 0 : 1) : (sg.dl ? 2 : 3);                      // uni_denoiser.py:373-378
            }
            return f;
        };
        auto features = [&](const FIn& in) {
            const int type = in.type;
            uint32_t hi[12], lo[12];
            if (type >= 0) {
                const float r0 = in.xd0 - in.xs0, r1 = in.xd1 - in.xs1, r2 = in.xd2 - in.xs2;
                const float dist = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
                float f[24];
#pragma unroll
                for (int gg = 0; gg < 20; gg++) { const float dd = dist - c_smear_off[gg]; f[gg] = __expf(-0.5f * dd * dd); }
                f[20] = 1.0f;
                f[21] = in.cs0 * in.cd0 + in.cs1 * in.cd1 + in.cs2 * in.cd2;
                f[22] = -(in.cs0 * r0 + in.cs1 * r1 + in.cs2 * r2);                   // vec_3 = x[src] - x[dst]
                f[23] = -(in.cd0 * r0 + in.cd1 * r1 + in.cd2 * r2);
#pragma unroll
                for (int i = 0; i < 12; i++) tc::split_pair_trunc(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
            }
            uint8_t* base = sFeat + (r >> 3) * 256 + (r & 7) * 16;
            auto chunk = [&](int c) { return base + (c >> 1) * 4096 + (c & 1) * 128; };
            if (prev_type >= 0 && prev_type != type) {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    uint8_t* p = chunk(prev_type * 3 + j);
                    *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(p + FEAT_PART) = make_uint4(0, 0, 0, 0);
                }
            }
            if (type >= 0) {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    uint8_t* p = chunk(type * 3 + j);
                    *reinterpret_cast<uint4*>(p) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    *reinterpret_cast<uint4*>(p + FEAT_PART) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
            prev_type = type;
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&bars[B_FEAT]);
        };
        constexpr uint32_t idesc_t = tc::umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_w = tc::umma_idesc_bf16(128, NOUT);
        const uint32_t sW_u32 = tc::smem_u32(sW), sTab_u32 = tc::smem_u32(sTab), sFeat_u32 = tc::smem_u32(sFeat);
        auto table_mma = [&]() {                                // MMA warp only
            // warp-collective issue (pg_tc.cuh): descriptor arithmetic on the uniform datapath, one elected lane issues
            uint32_t acc = 0;
#pragma unroll
            for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                const uint32_t fb = sFeat_u32 + (combo == 2 ? FEAT_PART : 0), wb = sTab_u32 + (combo == 1 ? TAB_PART : 0);
#pragma unroll
                for (int ks = 0; ks < KF / 16; ks++) {
                    tc::umma_bf16_w(tmem + C_PRE, tc::umma_desc_k16_noswizzle(fb + ks * 4096), tc::umma_desc_k16_noswizzle(wb + ks * 4096), idesc_t, acc);
                    acc = 1;
                }
            }
            tc::umma_commit_w(&bars[B_PRE]);
        };
        // fill the index pipeline (the only place where the chain is walked in one go)
        const long long f0 = tile_begin;
        FIn in_cur;
        int s1;
        KSeg seg1, seg2;
        int g3;
        {
            const KSeg seg0 = seg_at(f0, graph_at(f0));
            in_cur = load_in(src_at(f0, seg0), seg0);
            seg1 = seg_at(f0 + fstep, graph_at(f0 + fstep));
            s1 = src_at(f0 + fstep, seg1);
            seg2 = seg_at(f0 + 2 * fstep, graph_at(f0 + 2 * fstep));
            g3 = graph_at(f0 + 3 * fstep);
        }
        // Rotated loop: iteration n computes the features and issues the table product of the CTA's n-th tile, and issues the
        // second Linear of tile n-1.  Every tile goes through the SAME feature / MMA code (one call site each): a peeled
        // first tile is compiled separately and may round differently, which would make a molecule's result depend on its
        // position in the batch (tests: bit-exact batch independence).
        int tcount = 0;
        for (long long ft = f0;; ft += fstep, tcount++) {
            const bool have = ft < ntiles;
            const uint32_t pp = (uint32_t)(tcount - 1) & 1;         // phase parity of tile n-1
            KTRACE(1, 0);
            // the table MMA of tile n-1 has consumed the feature operand
            if (tcount > 0) tc::mbar_wait_wd(&bars[B_PRE], pp);
            KTRACE(1, 1);
            if (have) {
                features(in_cur);                                   // inputs fetched while the previous tile was processed
                in_cur = load_in(s1, seg1);                         // tile n+1
                s1 = src_at(ft + 2 * fstep, seg2);                  // neighbour index of tile n+2
                seg1 = seg2;
                seg2 = seg_at(ft + 3 * fstep, g3);                  // offsets of tile n+3
                g3 = graph_at(ft + 4 * fstep);                      // graph of tile n+4
            }
            KTRACE(1, 2);
            if (warp == MMA_WARP) {
                if (tcount > 0) {
                    tc::mbar_wait_wd(&bars[B_HID], pp);
                    KTRACE(1, 3);
                    tc::tc_fence_after();
                    if (KF16) {
                        constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A, B = F16
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint64_t bd = tc::umma_desc_sw128(sW_u32 + (ks >> 2) * 16384 + (ks & 3) * 32);
                            tc::umma_bf16_ts_w(tmem + C_OUT + pp * 128, tmem + C_HID + ks * 8, bd, idesc16, ks > 0);
                        }
                        tc::umma_commit_w(&bars[B_OUT]);
                    } else {
                        const uint32_t dcol = tmem + C_OUT + pp * 128;
                        uint32_t acc = 0;
#pragma unroll
                        for (int combo = 0; combo < 3; combo++) {
                            const uint32_t abase = tmem + C_HID + (combo == 2 ? 64 : 0);
                            const uint32_t bbase = sW_u32 + (combo == 1 ? W_TILE : 0);
#pragma unroll
                            for (int ks = 0; ks < 8; ks++) {
                                const uint64_t bd = tc::umma_desc_sw128(bbase + (ks >> 2) * 16384 + (ks & 3) * 32);
                                tc::umma_bf16_ts_w(dcol, abase + ks * 8, bd, idesc_w, acc);
                                acc = 1;
                            }
                        }
                        tc::umma_commit_w(&bars[B_OUT]);
                    }
                }
                KTRACE(1, 4);
                if (have) {
                    tc::mbar_wait_wd(&bars[B_FEAT], (uint32_t)tcount & 1);   // features of tile n (pre columns are free: HID(n-1) has fired)
                    tc::tc_fence_after();
                    KTRACE(1, 5);
                    table_mma();
                    KTRACE(1, 6);
                }
            }
            if (!have) break;
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= row warps: thread = (row, channel quarter) =================
        const int cq = warp >> 2;
        const bool fold = a.w.fold[PASS] > 0.5f;
        const int sr = lane >> 3, ch = (lane & 7) * 4;
        float* xp = sXp + warp * (16 * XP_LD);
        float al[4] = {0.f, 0.f, 0.f, 0.f};
        KSeg psg = kseg(d, tile_begin, wq);                 // segment of the tile whose post-processing is pending
        bool prow = false;
        float rel0 = 0.f, rel1 = 0.f, rel2 = 0.f;
        float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);       // per-row inputs of the pending tile's post-processing, fetched a tile ahead:
        float pfs = 0.f;                                    //   key pass: e_w (pf.x); value pass: alpha' of the 4 heads, per-head alpha' sum

        int tcount = 0;
        // ---- post-processing of a finished tile (its out columns: buffer `ob`)
        // `wait`: only the drain call waits for OUT itself.  Inside the loop the LayerNorm of the current tile has already
        // waited for OUT(t-1); waiting again here, after this thread's HID(t) arrival, could see OUT(t) completed as well and
        // then block on the parity of a phase that needs this very thread (two-phase aliasing of the parity wait).
        auto post = [&](uint32_t parity, uint32_t ob, bool wait) {
            const uint32_t out = tmem + C_OUT + ob * 128 + lane_base;
            if (PASS == 0) {
                // logits of this thread's 4 heads (key bias dropped: softmax-invariant), softmax over the lanes, times e_w
                const float* qrow = sQ + (warp * 2 + ob) * 32;          // staged by this warp one tile ago (cp.async)
                float4 qv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) qv[i] = ld4(qrow + i * 4);
                const float ew = pf.x;
                KTRACE(0, 8);
                if (wait) tc::mbar_wait(&bars[B_OUT], parity);
                tc::tc_fence_after();
                KTRACE(0, 9);
                uint32_t vv[32];
                tc::tmem_ld32_nowait(out + cq * 32, vv);
                tc::tmem_ld_wait();
                KTRACE(0, 10);
                constexpr float kScale = kInvSqrtD * 1.4426950408889634f;
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int o = h * 8;
                    const float4 qa = qv[2 * h], qb = qv[2 * h + 1];
                    float2 s0 = tc::mul2(make_float2(qa.x, qa.y), make_float2(__uint_as_float(vv[o]), __uint_as_float(vv[o + 1])));
                    float2 s1 = tc::mul2(make_float2(qb.x, qb.y), make_float2(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5])));
                    s0 = tc::fma2(make_float2(qa.z, qa.w), make_float2(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3])), s0);
                    s1 = tc::fma2(make_float2(qb.z, qb.w), make_float2(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7])), s1);
                    s0 = tc::add2(s0, s1);
                    al[h] = prow ? (s0.x + s0.y) * kScale : -INFINITY;
                }
                KTRACE(0, 11);
                // segment softmax across the 32 lanes on the REDUX unit (max; fixed-point sums of values in [0, 1]) instead
                // of three 5-round shuffle butterflies: the dependent shuffle chains were a quarter of the tile time
                float sw[4];
#if defined(PG_KNN_EXP) && (PG_KNN_EXP & 1)
#pragma unroll
                for (int h = 0; h < 4; h++) { al[h] = prow ? tc::ex2_approx(al[h] * 1e-3f) * 0.03f * ew : 0.f; sw[h] = al[h]; }     // knock-out: no reductions (wrong results)
                if (false)
#endif
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const float mx = tc::warp_max_redux(al[h]);
                    al[h] = prow ? tc::ex2_approx(al[h] - mx) : 0.f;
                }
#if !(defined(PG_KNN_EXP) && (PG_KNN_EXP & 1))
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const float sm = tc::warp_sum01_redux(al[h]);
                    al[h] = prow ? al[h] * tc::rcp_approx(sm) * ew : 0.f;      // e_w is a sigmoid: alpha' stays in [0, 1]
                }
#pragma unroll
                for (int h = 0; h < 4; h++) sw[h] = tc::warp_sum01_redux(al[h]);
#endif
                KTRACE(0, 12);
                if (prow) st4(a.alpha + (size_t)(psg.e0 + lane) * 16 + cq * 4, make_float4(al[0], al[1], al[2], al[3]));
                if (psg.valid && lane == 0) st4(a.alpha_sum + (size_t)psg.v * 16 + cq * 4, make_float4(sw[0], sw[1], sw[2], sw[3]));
            } else {
                const float4 a4 = pf;
                if (wait) tc::mbar_wait(&bars[B_OUT], parity);
                tc::tc_fence_after();
                if (POS == 0) {
                    const float swc = pfs;
                    uint32_t vu[32];
                    tc::tmem_ld32_nowait(out + cq * 32, vu);
                    tc::tmem_ld_wait();
                    const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 pr = tc::mul2(make_float2(aa[i >> 3], aa[i >> 3]), make_float2(__uint_as_float(vu[i]), __uint_as_float(vu[i + 1])));
                        v[i] = pr.x; v[i + 1] = pr.y;
                    }
                    const float o = transpose_reduce32(v, lane);
                    if (psg.inr) {                                                          // isolated node: no messages
                        const int c = cq * 32 + lane;
                        a.out[(size_t)psg.v * 128 + c] = psg.valid ? fmaf(sB2[c], swc, o) : 0.f;
                    }
                } else {
                    float vh[4];
                    tc::tmem_ld4(out + cq * 4, vh);
                    float c = a4.x * (vh[0] + sB2[cq * 4]) + a4.y * (vh[1] + sB2[cq * 4 + 1]) + a4.z * (vh[2] + sB2[cq * 4 + 2]) + a4.w * (vh[3] + sB2[cq * 4 + 3]);
                    float p0 = c * rel0, p1 = c * rel1, p2 = c * rel2;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        p0 += __shfl_xor_sync(PG_FULL, p0, o); p1 += __shfl_xor_sync(PG_FULL, p1, o); p2 += __shfl_xor_sync(PG_FULL, p2, o);
                    }
                    float* rd = sRed + (wq * 4 + cq) * 4;
                    if (lane == 0) { rd[0] = p0; rd[1] = p1; rd[2] = p2; }
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
                    if (cq == 0 && lane < 3 && psg.inr) {
                        const float* r4 = sRed + wq * 16 + lane;
                        a.out[(size_t)psg.v * 3 + lane] = psg.valid ? ((r4[0] + r4[4]) + (r4[8] + r4[12])) * (1.0f / 16.0f) : 0.f;
                    }
                    // the next writer of sRed may be the drain call right after the loop, with no LayerNorm barrier in between
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
                }
            }
        };

        uint32_t ph = 0;
        bool any = false;
        KSeg sg = kseg(d, tile_begin, wq);
        // index pipeline: graph of the segment two tiles ahead -> its offsets one tile ahead -> neighbour indices one tile
        // ahead -> gathers at the start of the tile; every level is requested a phase before its first use
        // (tiles past the end fall back to the CTA's first tile: always a valid address, never used for real work)
        auto tile_or_first = [&](long long t) { return t < ntiles ? t : tile_begin; };
        KSeg nsg = kseg(d, tile_or_first(tile_begin + 1), wq);
        int g2 = kseg_graph(d, tile_or_first(tile_begin + 2), wq);
        // neighbour index of this lane's row (padded rows re-read row 0, an isolated node reads node 0); the gather below
        // needs the indices of rows 4 i + (lane >> 3) and takes them from their lanes by shuffle (2 registers instead of 16)
        int myidx = sg.valid ? a.knn_src[sg.e0 + (lane < sg.R ? lane : 0)] : 0;
        for (long long tile = tile_begin; tile < ntiles; tile++, ph ^= 1, tcount++) {
            const bool rowvalid = sg.valid && lane < sg.R;
            KTRACE(0, 0);
            KWTRACE(0);
            // ---- gather the src-node partial rows: 8 lanes cover the 128-byte slice of one row (coalesced), the dst-node
            //      partial is added in that layout, and the rows are transposed to row-per-lane through the warp's shared
            //      tile (two rounds of 16 rows).  Neighbour indices were fetched one tile ahead.
            const int c0 = cq * 32;
            const float* ps = a.nc.A + (PASS == 0 ? a.nc.src_k : a.nc.src_v) + c0 + ch;
            float4 u[8];
#pragma unroll
#if defined(PG_KNN_EXP) && (PG_KNN_EXP & 2)
            for (int i = 0; i < 8; i++) u[i] = make_float4(0.1f * i, 0.2f, 0.3f * lane, 0.4f);      // knock-out: no gather (wrong results)
            const float4 d4 = make_float4(0.f, 1.f, 0.f, -1.f); (void)ps;
#else
            for (int i = 0; i < 8; i++) u[i] = ldg4(ps + (size_t)__shfl_sync(PG_FULL, myidx, i * 4 + sr) * a.nc.lda);
            const float4 d4 = ldg4(a.nc.A + (size_t)sg.v * a.nc.lda + (PASS == 0 ? a.nc.dst_k : a.nc.dst_v) + c0 + ch);
#endif
            // this tile's post-processing inputs (consumed one iteration later) and the next tile's neighbour indices
            float4 nf = make_float4(0.f, 0.f, 0.f, 0.f);
            float nfs = 0.f;
            if (PASS == 0) { if (rowvalid) nf.x = a.ew[sg.e0 + lane]; }
            else {
                if (rowvalid) nf = ld4(a.alpha + (size_t)(sg.e0 + lane) * 16 + cq * 4);
                if (POS == 0 && sg.valid) nfs = a.alpha_sum[(size_t)sg.v * 16 + cq * 4 + (lane >> 3)];
            }
            if (PASS == 0 && lane < 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(sQ + (warp * 2 + ph) * 32 + lane * 4)),
                             "l"(a.q + (size_t)sg.v * 128 + c0 + lane * 4) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            KTRACE(0, 13);
            // next tile's neighbour indices (its offsets were fetched during the previous tile), the offsets of the tile after
            // it (graph fetched during the previous tile) and the graph of the tile after that
            const int nmyidx = nsg.valid ? a.knn_src[nsg.e0 + (lane < nsg.R ? lane : 0)] : 0;
            const long long t2 = tile_or_first(tile + 2), t3 = tile_or_first(tile + 3);
            const KSeg nsg2 = kseg_of_graph(d, t2, wq, g2);
            const int g3 = kseg_graph(d, t3, wq);
            KTRACE(0, 14);
            float2 x2[16];
            KTRACE(0, 1);
#pragma unroll
            for (int round = 0; round < 2; round++) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; i++) st4(xp + (i * 4 + sr) * XP_LD + ch, f4add(u[round * 4 + i], d4));
                __syncwarp();
                if ((lane >> 4) == round) {
                    const float* xrow = xp + (lane & 15) * XP_LD;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float4 e4 = ld4(xrow + 2 * i);
                        x2[i] = make_float2(e4.x, e4.y);
                        x2[i + 1] = make_float2(e4.z, e4.w);
                    }
                }
            }
            // ---- table product from TMEM, LayerNorm + ReLU, bf16 hi/lo A operand
            KTRACE(0, 2);
            tc::mbar_wait(&bars[B_PRE], ph);
            tc::tc_fence_after();
            KTRACE(0, 3);
            KWTRACE(1);
            {
                uint32_t xu[32];
                tc::tmem_ld32_nowait(tmem + C_PRE + lane_base + cq * 32, xu);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) x2[i] = tc::add2(x2[i], make_float2(__uint_as_float(xu[2 * i]), __uint_as_float(xu[2 * i + 1])));
            }
            // the pre-activation is mean-free (every first-Linear block and table row has its channel mean removed at pack
            // time, weights._center_first_linears): the LayerNorm only needs the second moment
            float2 s2 = make_float2(0.f, 0.f), s2b = s2;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                s2 = tc::fma2(x2[i], x2[i], s2); s2b = tc::fma2(x2[i + 1], x2[i + 1], s2b);
            }
            s2 = tc::add2(s2, s2b);
            // quarter-major layout [quarter][row]: consecutive lanes touch consecutive words (no bank conflicts)
            float* st = sStat + wq * 32 + lane;
            st[cq * 128] = s2.x + s2.y;
            asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
            KTRACE(0, 4);
            const float rstd = rsqrtf(((st[0] + st[128]) + (st[256] + st[384])) * (1.0f / 128.0f) + 1e-5f);
            const float2 rs2 = make_float2(rstd, rstd);
            const float* gam = sLn + cq * 32;
            const float* bet = gam + 128;
            if (KF16) {
                // key MLP: one fp16 value per activation (its output only feeds the softmax logits; see pg_trip_tc.cu)
                uint32_t hh[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 b4 = ld4(bet + 2 * i);
                    float2 y0, y1;
                    if (fold) {
                        y0 = tc::fma2(x2[i], rs2, make_float2(b4.x, b4.y)); y1 = tc::fma2(x2[i + 1], rs2, make_float2(b4.z, b4.w));
                    } else {
                        const float4 g4 = ld4(gam + 2 * i);
                        y0 = tc::fma2(tc::mul2(x2[i], rs2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
                        y1 = tc::fma2(tc::mul2(x2[i + 1], rs2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
                    }
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hh[i]) : "f"(y0.y), "f"(y0.x));
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hh[i + 1]) : "f"(y1.y), "f"(y1.x));
                }
                if (any) tc::mbar_wait(&bars[B_OUT], ph ^ 1);
                tc::tc_fence_after();
                tc::tmem_st16(tmem + C_HID + lane_base + cq * 16, hh);
            } else {
            uint32_t hi[16], lo[16];
            if (fold) {
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
            // the W2 MMA of the previous tile must be done with the hid columns before they are rewritten
            if (any) tc::mbar_wait(&bars[B_OUT], ph ^ 1);
            tc::tc_fence_after();
            tc::tmem_st16(tmem + C_HID + lane_base + cq * 16, hi);
            tc::tmem_st16(tmem + C_HID + lane_base + 64 + cq * 16, lo);
            }
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&bars[B_HID]);
            KTRACE(0, 5);
            KWTRACE(2);
            // ---- post-processing of the previous tile while the tensor pipe works on this one
            if (any) { asm volatile("cp.async.wait_group 1;" ::: "memory"); __syncwarp(); post(ph ^ 1, ph ^ 1, false); }
            KTRACE(0, 6);
            KWTRACE(3);
            psg = sg; prow = rowvalid; pf = nf; pfs = nfs;
            if (PASS == 1 && POS) {
                const int s = rowvalid ? a.knn_src[sg.e0 + lane] : sg.v;
                rel0 = a.x[(size_t)sg.v * 3] - a.x[(size_t)s * 3];                          // rel_x = x[dst] - x[src]
                rel1 = a.x[(size_t)sg.v * 3 + 1] - a.x[(size_t)s * 3 + 1];
                rel2 = a.x[(size_t)sg.v * 3 + 2] - a.x[(size_t)s * 3 + 2];
            }
            any = true;
            sg = nsg; nsg = nsg2; g2 = g3;
            myidx = nmyidx;
        }
        if (any) { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); post(ph ^ 1, ph ^ 1, true); }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}
}  // namespace

int pg_launch_knn_tc(const KnnTcArgs& a, int pos, int num_sms, cudaStream_t s) {
    if (a.d.N <= 0) return PG_OK;
    static bool init = false;
    if (!init) {
        PG_CUDA_CHECK((cudaFuncSetAttribute(knn_tc_kernel<0, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)));
        PG_CUDA_CHECK((cudaFuncSetAttribute(knn_tc_kernel<0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)));
        PG_CUDA_CHECK((cudaFuncSetAttribute(knn_tc_kernel<1, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)));
        PG_CUDA_CHECK((cudaFuncSetAttribute(knn_tc_kernel<1, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)));
        init = true;
    }
    const long long ntiles = (a.d.N + 3) / 4;
    const unsigned grid = (unsigned)std::min<long long>(ntiles, num_sms);
    if (a.key_bf16x3) knn_tc_kernel<0, 0, false><<<grid, NTHREADS, SM_TOTAL, s>>>(a);          // key pass: alpha * e_w -> scratch
    else knn_tc_kernel<0, 0, true><<<grid, NTHREADS, SM_TOTAL, s>>>(a);
    PG_LAUNCH_CHECK();
    if (pos == 0) knn_tc_kernel<1, 0, false><<<grid, NTHREADS, SM_TOTAL, s>>>(a);
    else knn_tc_kernel<1, 1, false><<<grid, NTHREADS, SM_TOTAL, s>>>(a);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

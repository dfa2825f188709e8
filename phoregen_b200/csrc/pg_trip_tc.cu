// BondUpdateLayer (uni_denoiser.py:123-165) on the 5th-gen tensor cores.
//
// One persistent CTA per SM (640 threads) walks the ligand atoms j ("units").  For a unit the per-edge partials
// P[k->j] (n-1 rows x 256 channels, k|v; written by trip_pr_kernel below) are staged once in shared memory by bulk copies;
// the n-1 segments (j->i) are processed four at a time as a 128-row tile: TMEM lane = triplet row (32 lanes per segment,
// rows k ascending).
//
// Roles: 16 row warps, thread = (row, 32-channel quarter), warp w -> lane quarter w & 3, channel quarter w >> 2 (four row
//        warps per scheduler); warp 16 issues every MMA and the cp.async / bulk copies; warps 17-19 compute the angular
//        features of the tiles ahead.  setmaxnreg moves registers from the auxiliary warpgroup to the row warps (104 / 64).
//
//   1. angular encoding (13 values per triplet) -> bf16 hi/lo A tile in smem (double buffered); per MLP three
//      tcgen05.mma (M128 N128 K16, bf16x3) against the angle slice of the first Linear -> pre-activation columns in TMEM.
//   2. every thread reads its 32-channel slice (tcgen05.ld), adds P[k->j] (smem) and R[j->i] (per segment), applies
//      LayerNorm + ReLU in registers with packed fp32x2 math (the four quarter statistics of a row meet in shared memory),
//      splits to bf16 hi/lo with the ReLU fused into the conversions and writes the A operand of the second Linear back to
//      TMEM (tcgen05.st).  Positive LayerNorm gains are folded into W2 at pack time (weights.py), so only beta/gamma is added.
//   3. second Linear of the key and value MLPs: A from TMEM, B = W2 (bf16 hi/lo, resident in smem, 128B swizzle),
//      24 tcgen05.mma (M128 N128 K16) each, fp32 accumulators in TMEM.
//   4. logits = q . k per head (thread-local dot; the key bias is softmax-invariant and dropped), segment softmax across
//      the 32 lanes on the REDUX unit, alpha-weighted sum of v over the lanes with a butterfly transpose-reduce, residual
//      add into h_bond.
//
// Software pipeline across tiles: the row warps run  LN-k(t) | epilogue(t-1) | LN-v(t) | logits(t)  while the tensor pipe
// runs  W2k(t), angle-k(t+1) | W2v(t), angle-v(t+1).  The four 128-column TMEM blocks alternate roles per tile (pre/out of
// one MLP in one block, hid in the other), so two tiles share the 512 columns.  All hand-offs are mbarriers; the ordering
// argument behind every parity wait is in DESIGN.md ("mbarrier parity waits").
// Precision: bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate) on every contraction, fp32 everywhere else.
#include <algorithm>
#include "pg_attn.h"
#include "pg_tc.cuh"

// Optional phase tracing (debug builds only): cycle stamps of CTA 0 into a.trace[role][tile][slot]
#ifdef PG_TRIP_TRACE
__device__ long long g_trip_trace[3 * 64 * 16];
extern "C" int pg_debug_trip_trace(long long* h_out) { return cudaMemcpyFromSymbol(h_out, g_trip_trace, sizeof(g_trip_trace)) == cudaSuccess ? 0 : -2; }
#define TRACE(role, slot) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 16) && tcount < 64) g_trip_trace[((role) * 64 + tcount) * 16 + (slot)] = clock64(); } while (0)
#else
#define TRACE(role, slot) do {} while (0)
#endif

namespace {
// Staged P rows.  trip_pr_kernel writes the key and the value halves of P to two arrays [Eb][132]: 128 channels + 4 floats of
// padding per row, i.e. already in the shared-memory layout.  The rows of a unit (or of one 32-row chunk of it) are then ONE
// contiguous block per half, staged by a single bulk copy instead of one per row, and a warp's row-per-lane LDS.128 (row
// stride 528 B = 33 x 16 B) touches every bank exactly once per 8 lanes.
constexpr int PS_HALF = PG_TRIP_P_STRIDE;  // floats per row of one half (132)
// staged P rows: all n-1 rows of a unit while a segment is one chunk (n-1 <= 33), else the 33 rows of one chunk
__host__ __device__ constexpr int ps_rows(int maxn) { return maxn - 1 < 33 ? maxn - 1 : 33; }
constexpr int W_TILE = 32768;              // one [128 x 128] bf16 matrix in two 128B-swizzled K blocks
constexpr int SM_W = 4 * W_TILE;           // (k,v) x (hi,lo)
constexpr int SM_WA = 2 * 8192;            // angle slice of the first Linear: (hi,lo) x [256 x 16] bf16, no swizzle
constexpr int SM_FEAT1 = 2 * 4096;         // (hi,lo) x [128 x 16] bf16, no swizzle
constexpr int SM_FEAT = 2 * SM_FEAT1;      // double-buffered: the feature warps run up to two tiles ahead
constexpr int SM_QR = 2 * (4 * 128 + 4 * 256) * 4;   // double-buffered query rows + r_ji rows of a tile's 4 segments
constexpr int SM_ALPHA = 128 * 16 * 4;     // attention weights of the tile [row][head]
constexpr int SM_FIXED = SM_W + SM_WA + SM_FEAT + SM_QR + SM_ALPHA + 6 * 128 * 4 /*ln + b2*/ + 128 /*barriers*/;
constexpr float kInvSqrtD = 0.35355339059327373f;
constexpr int ROW_WARPS = 16;               // 4 warps per TMEM lane quarter, each owning a 32-channel slice of the row
constexpr int MMA_WARP = ROW_WARPS;         // MMA issue + q/R/P loaders; the 3 warps after it compute angular features
constexpr int NTHREADS = (ROW_WARPS + 4) * 32;
constexpr int ROW_THREADS = ROW_WARPS * 32;

// mbarrier slots
// A parity wait is only safe while the barrier cannot complete a second time before the waiter looks at it.  B_PREV has
// two kinds of waiters (row warps and the feature warps, which run ahead on their own clock), so it alternates between
// two barriers by tile parity, like B_FEAT.
enum { B_FEAT0 = 0, B_FEAT1, B_PREK, B_PREV0, B_PREV1, B_HIDK, B_HIDV, B_OUTK, B_OUTV, B_PSK, B_PSV, B_COUNT };

// the sequence of (unit, tile) a CTA walks; every role steps through it redundantly
// MULTI: a segment longer than 32 rows is cut into chunks of 32 rows; a tile is then (group of 4 segments) x (chunk c),
// the chunks of a group on consecutive tiles (chunk-inner order), and the softmax runs on-line across them.
// `stage` counts the refills of the staged P rows: once per unit when a segment is a single chunk (all n-1 rows of the
// unit stay resident), once per tile otherwise (the 33 rows k -> j of chunk c).
struct TileIter {
    int u, tile, ntile, n, jl, ctx0, nchunk, chunk, grp, stage;
    long long eoff;
    bool valid;
};
template <bool MULTI>
__device__ __forceinline__ void iter_load_unit(const PlanDev& d, TileIter& it) {
    it.valid = false;
    while (it.u < d.Nl) {
        const int g = d.lig_graph[it.u];
        const int n = d.g_n[g];
        if (n >= 3 && (MULTI ? n - 2 > 32 : n - 2 <= 32)) {       // each instantiation takes its own molecules (see the launcher)
            it.n = n; it.jl = it.u - d.lig_off[g]; it.ctx0 = d.ctx_off[g] + d.g_p[g]; it.eoff = d.eoff[g];
            it.nchunk = MULTI ? (n - 2 + 31) >> 5 : 1;
            it.ntile = ((n - 1 + 3) >> 2) * it.nchunk; it.tile = 0; it.chunk = 0; it.grp = 0; it.stage++; it.valid = true;
            return;
        }
        it.u += gridDim.x;
    }
}
template <bool MULTI>
__device__ __forceinline__ void iter_next(const PlanDev& d, TileIter& it) {
    if (++it.tile < it.ntile) {
        if (MULTI && it.nchunk > 1) {
            it.stage++;
            if (++it.chunk == it.nchunk) { it.chunk = 0; it.grp++; }
        } else {
            it.grp = it.tile;
        }
        return;
    }
    it.u += gridDim.x;
    iter_load_unit<MULTI>(d, it);
}
// segment handled by lane quadrant wq in this tile
struct Seg { bool valid; int il, ti; long long eji; };
__device__ __forceinline__ Seg seg_of(const TileIter& it, int wq) {
    Seg s;
    const int sidx = it.grp * 4 + wq;
    s.valid = sidx < it.n - 1;
    s.il = s.valid ? sidx + (sidx >= it.jl) : (it.jl == 0 ? 1 : 0);
    s.ti = s.il - (s.il > it.jl);
    s.eji = it.eoff + (long long)s.il * (it.n - 1) + (it.jl - (it.jl > s.il));
    return s;
}

// angular encoding of this thread's triplet row -> bf16 hi/lo A tile (common.py:67-87; uni_denoiser.py:131-135)
// xs: coordinates of the molecule's ligand atoms, [n][4] floats (shared memory copy)
// `wq` = segment slot of the tile (0..3), `row` = row inside the segment (0..31)
__device__ __forceinline__ void write_features(const float* xs, const TileIter& it, int wq, int row, uint8_t* sFeat) {
    const int lane = row;
    const Seg sg = seg_of(it, wq);
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = 0.f;
    const int srow = it.chunk * 32 + lane;          // row inside the segment
    if (sg.valid && srow < it.n - 2) {
        const int trow = srow + (srow >= sg.ti);
        const float4 xj = ld4(xs + it.jl * 4), xi = ld4(xs + sg.il * 4), xk = ld4(xs + (trow + (trow >= it.jl)) * 4);
        const float xi0 = xi.x, xi1 = xi.y, xi2 = xi.z;
        const float pj0 = xj.x - xi0, pj1 = xj.y - xi1, pj2 = xj.z - xi2;
        const float pk0 = xk.x - xi0, pk1 = xk.y - xi1, pk2 = xk.z - xi2;
        const float dotv = pj0 * pk0 + pj1 * pk1 + pj2 * pk2;
        const float c0 = pj1 * pk2 - pj2 * pk1, c1 = pj2 * pk0 - pj0 * pk2, c2 = pj0 * pk1 - pj1 * pk0;
        const float th = atan2f(sqrtf(c0 * c0 + c1 * c1 + c2 * c2), dotv);
        // th in [0, pi]: fast sincos is accurate to ~5e-7 there; double / triple angle by identities
        float s1, k1, sh, kh, st, kt;
        __sincosf(th, &s1, &k1); __sincosf(th * 0.5f, &sh, &kh); __sincosf(th * (1.0f / 3.0f), &st, &kt);
        const float s2 = 2.0f * s1 * k1, k2 = fmaf(-2.0f * s1, s1, 1.0f);
        const float s3 = s1 * fmaf(-4.0f * s1, s1, 3.0f), k3 = k1 * fmaf(4.0f * k1, k1, -3.0f);
        f[0] = th; f[1] = s1; f[2] = s2; f[3] = s3; f[4] = s1; f[5] = sh; f[6] = st;
        f[7] = k1; f[8] = k2; f[9] = k3; f[10] = k1; f[11] = kh; f[12] = kt;
    }
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) tc::split_pair_trunc(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
    const int r = wq * 32 + lane;
    uint8_t* ph = sFeat + (r >> 3) * 256 + (r & 7) * 16;
    *reinterpret_cast<uint4*>(ph) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(ph + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(ph + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(ph + 4096 + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

// KF16: precision of the key MLP's second Linear, whose output only feeds the softmax logits.
//   false (default) = bf16x3 like every other contraction (24 MMAs).
//   true  (PG_KEY=trip16 | fp16, opt-in) = ONE fp16 value per activation and per weight: 8 MMAs and one conversion per
//   channel pair, 7.0 -> 6.2 ms per launch at configs[1].  Measured on the fixtures: model outputs stay inside the parity
//   bar (worst 0.50 x tolerance against 0.41), but the denoiser's internal h_bond reaches 1.07 x tolerance (one element
//   in 66 k; the per-row rounding of the activations does not cancel in the softmax, a hi/lo weight pair does not help),
//   so it is not the default.  A single-pass VALUE path misses the bar by 1.5x (measured) and is not offered.
template <bool MULTI, bool KF16>
__global__ void __launch_bounds__(NTHREADS, 1) trip_tc_kernel(TripTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // offset arithmetic on the __shared__ array (not a uintptr_t round trip) so the compiler keeps emitting LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sWa = sW + SM_W;
    uint8_t* sFeat = sWa + SM_WA;
    float* sQR = (float*)(sFeat + SM_FEAT);         // [2][ q: 4x128 | R: 4x256 ]
    float* sStat = sQR + SM_QR / 4;                 // [2 mlp][4 quarters][128 rows][2]  partial LayerNorm sums (quarter-major: conflict-free)
    float* sLn = sStat + SM_ALPHA / 4;              // gk, bk, gv, bv
    float* sB2 = sLn + 4 * 128;                     // b2k, b2v
    uint64_t* bars = (uint64_t*)(sB2 + 2 * 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + B_COUNT);
    float* sPs = (float*)(smem + SM_FIXED);         // [k|v][min(maxn-1, 33)][132] (see PS_HALF)
    const int psr = ps_rows(a.maxn);
    float* sX = sPs + (size_t)2 * psr * PS_HALF;    // [2][maxn][4] ligand coordinates of the current / next unit
    const PlanDev& d = a.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp & 3;

    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bars[B_FEAT0], 96 + 32); tc::mbar_init(&bars[B_FEAT1], 96 + 32);
        tc::mbar_init(&bars[B_PREK], 1); tc::mbar_init(&bars[B_PREV0], 1); tc::mbar_init(&bars[B_PREV1], 1);
        tc::mbar_init(&bars[B_HIDK], ROW_THREADS); tc::mbar_init(&bars[B_HIDV], ROW_THREADS);
        tc::mbar_init(&bars[B_OUTK], 1); tc::mbar_init(&bars[B_OUTV], 1);
        tc::mbar_init(&bars[B_PSK], 1); tc::mbar_init(&bars[B_PSV], 1);
        tc::fence_barrier_init();
    }
    // ---- resident weights
    for (int idx = tid; idx < 4 * 128 * 16; idx += NTHREADS) {     // 16-byte chunks: [mat 4][n 128][chunk 16]
        const int mat = idx >> 11, n = (idx >> 4) & 127, c = idx & 15;
        if (KF16 && mat == 1) continue;                               // the key MLP has a single fp16 image (tile 0)
        const uint16_t* src = ((mat >> 1) ? a.w2v_bf : (KF16 ? a.w2k_h : a.w2k_bf)) + ((size_t)(mat & 1) * 128 + n) * 128 + c * 8;
        const uint32_t dst = tc::smem_u32(sW) + mat * W_TILE + (c >> 3) * 16384 + tc::sw128_chunk(n, c & 7);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    for (int idx = tid; idx < 2 * 256 * 2; idx += NTHREADS) {      // [part 2][n 256][kc 2]
        const int part = idx >> 9, n = (idx >> 1) & 255, kc = idx & 1;
        const uint16_t* src = a.wa_bf + ((size_t)part * 256 + n) * 16 + kc * 8;
        const uint32_t dst = tc::smem_u32(sWa) + part * 8192 + (n >> 3) * 256 + kc * 128 + (n & 7) * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    if (tid < 128) {
        sLn[tid] = a.lnk_g[tid]; sLn[128 + tid] = a.lnk_bf[tid]; sLn[256 + tid] = a.lnv_g[tid]; sLn[384 + tid] = a.lnv_bf[tid];
        sB2[tid] = a.b2k[tid]; sB2[128 + tid] = a.b2v[tid];
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;

    TileIter it;
    it.u = blockIdx.x; it.stage = 0;
    iter_load_unit<MULTI>(d, it);
    if (!it.valid) {
        __syncthreads();
        if (warp == MMA_WARP) tc::tmem_dealloc<512>(tmem);
        return;
    }

    if (warp >= MMA_WARP) {
        // registers move from this warpgroup to the four row warpgroups
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
      if (warp == MMA_WARP) {
        // ================= MMA issue + loaders (query / r_ji rows via cp.async, P rows via bulk copy) =================
        constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 128);
        const uint32_t sW_u32 = tc::smem_u32(sW), sWa_u32 = tc::smem_u32(sWa), sFeat_u32 = tc::smem_u32(sFeat);
        auto load_qr = [&](const TileIter& t, int buf) {
            const uint32_t q = tc::smem_u32(sQR + buf * (SM_QR / 8));
            const uint32_t r = q + 4 * 128 * 4;
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const Seg sg = seg_of(t, s4);
                const float* qs = a.q + (size_t)sg.eji * 128 + lane * 4;
                const float* rs = a.R + (size_t)sg.eji * 256 + lane * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(q + (s4 * 128 + lane * 4) * 4), "l"(qs) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(r + (s4 * 256 + lane * 4) * 4), "l"(rs) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(r + (s4 * 256 + 128 + lane * 4) * 4), "l"(rs + 128) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // P rows of the edges k -> j of a unit (n-1 contiguous 1 KB rows) -> padded smem rows.  The key and value halves
        // travel separately (one mbarrier transaction each): each half is dead as soon as its LayerNorm pass of the
        // unit's last tile is done, so the next unit's half is requested a whole phase before it is needed.
        auto load_ps = [&](const TileIter& t, int mlp) {
            uint64_t* bar = &bars[mlp == 0 ? B_PSK : B_PSV];
            // rows k -> j of chunk c: segment rows [32c, 32c+32) map to unit rows [32c, 32c+33) (the row k = i is skipped)
            const int r0 = t.chunk * 32, nr = min(33, t.n - 1 - r0);
            if (lane == 0) tc::mbar_arrive_expect_tx(bar, (uint32_t)nr * (PS_HALF * 4u));
            __syncwarp();
            const float* src = a.P + (size_t)mlp * a.d.Eb * PS_HALF + (size_t)(t.eoff + (long long)t.jl * (t.n - 1) + r0) * PS_HALF;
            if (lane == 0) tc::bulk_copy_g2s(sPs + (size_t)mlp * psr * PS_HALF, src, (uint32_t)nr * (PS_HALF * 4u), bar);   // one copy per half
        };
        // angle slice of the first Linear for one MLP of tile `tl` (operand buffer tl & 1) -> pre-activation columns `dcol`
        auto feat_mma = [&](int tl, int mlp, uint32_t dcol, uint64_t* bar) {
            if (lane == 0) {
                const uint32_t fb = sFeat_u32 + (tl & 1) * SM_FEAT1;
                const uint64_t fh = tc::umma_desc_k16_noswizzle(fb), fl = tc::umma_desc_k16_noswizzle(fb + 4096);
                const uint64_t wh = tc::umma_desc_k16_noswizzle(sWa_u32 + mlp * 4096), wl = tc::umma_desc_k16_noswizzle(sWa_u32 + 8192 + mlp * 4096);
                tc::umma_bf16(dcol, fh, wh, idesc, 0);
                tc::umma_bf16(dcol, fh, wl, idesc, 1);
                tc::umma_bf16(dcol, fl, wh, idesc, 1);
                tc::umma_commit(bar);
            }
            __syncwarp();
        };
        // second Linear of one MLP: A = bf16 hi/lo activations in TMEM columns `hid`, D = `dcol`
        auto w2_mma = [&](int mlp, uint32_t hid, uint32_t dcol, uint64_t* bar) {
            if (KF16 && mlp == 0) {
                if (lane == 0) {
                    constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A, B = F16
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t bd = tc::umma_desc_sw128(sW_u32 + (ks >> 2) * 16384 + (ks & 3) * 32);
                        tc::umma_bf16_ts(dcol, hid + ks * 8, bd, idesc16, ks > 0);
                    }
                    tc::umma_commit(bar);
                }
                __syncwarp();
                return;
            }
            if (lane == 0) {
                uint32_t acc = 0;
#pragma unroll
                for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                    const uint32_t abase = hid + (combo == 2 ? 64 : 0);
                    const uint32_t bbase = sW_u32 + (mlp * 2 + (combo == 1 ? 1 : 0)) * W_TILE;
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t bd = tc::umma_desc_sw128(bbase + (ks >> 2) * 16384 + (ks & 3) * 32);
                        tc::umma_bf16_ts(dcol, abase + ks * 8, bd, idesc, acc);
                        acc = 1;
                    }
                }
                tc::umma_commit(bar);
            }
            __syncwarp();
        };
        load_ps(it, 0);
        load_ps(it, 1);
        load_qr(it, 0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        tc::mbar_arrive(&bars[B_FEAT0]);
        tc::mbar_wait_wd(&bars[B_FEAT0], 0);
        tc::tc_fence_after();
        feat_mma(0, 0, tmem + 0, &bars[B_PREK]);
        feat_mma(0, 1, tmem + 128, &bars[B_PREV0]);
        int tcount = 0;
        while (it.valid) {
            const uint32_t ph = tcount & 1;
            // column roles alternate per tile: even tiles pre/out in [0,128)|[128,256), hid in [256,384)|[384,512); odd swapped
            const uint32_t preK = tmem + (ph ? 256 : 0), hidK = tmem + (ph ? 0 : 256);
            const uint32_t preV = tmem + (ph ? 384 : 128), hidV = tmem + (ph ? 128 : 384);
            TileIter nx = it;
            iter_next<MULTI>(d, nx);
            const bool newu = nx.valid && nx.stage != it.stage;     // the staged P rows change with the next tile
            uint64_t* featbar = &bars[(tcount + 1) & 1 ? B_FEAT1 : B_FEAT0];
            TRACE(2, 0);
            tc::mbar_wait_wd(&bars[B_HIDK], ph);          // key activations of tile t are in TMEM; logits of tile t-1 are done
            TRACE(2, 1);
            tc::tc_fence_after();
            if (nx.valid) load_qr(nx, (tcount + 1) & 1);
            if (newu) load_ps(nx, 0);
            w2_mma(0, hidK, preK, &bars[B_OUTK]);
            TRACE(2, 2);
            if (nx.valid) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                tc::mbar_arrive(featbar);
                tc::mbar_wait_wd(&bars[B_OUTK], ph);      // hid_k(t) has been consumed: its columns take pre_k(t+1)
                tc::mbar_wait_wd(featbar, ((tcount + 1) >> 1) & 1);
                tc::tc_fence_after();
                TRACE(2, 3);
                feat_mma(tcount + 1, 0, hidK, &bars[B_PREK]);
            }
            TRACE(2, 4);
            tc::mbar_wait_wd(&bars[B_HIDV], ph);
            TRACE(2, 5);
            tc::tc_fence_after();
            if (newu) load_ps(nx, 1);
            w2_mma(1, hidV, preV, &bars[B_OUTV]);
            TRACE(2, 6);
            if (nx.valid) {
                tc::mbar_wait_wd(&bars[B_OUTV], ph);
                tc::tc_fence_after();
                TRACE(2, 7);
                feat_mma(tcount + 1, 1, hidV, &bars[(tcount + 1) & 1 ? B_PREV1 : B_PREV0]);
            }
            TRACE(2, 8);
            it = nx; tcount++;
        }
      } else {
        // ================= feature warps: angular features, up to two tiles ahead (double-buffered operand) =================
        // 96 threads cover the 128 rows of a tile in two passes; they also stage the next unit's coordinates.
        const int ft = tid - (MMA_WARP + 1) * 32;       // 0..95
        auto stage_x = [&](const TileIter& t, int b) {
            for (int i = ft; i < t.n; i += 96) {
                const float* src = a.x + (size_t)(t.ctx0 + i) * 3;
                st4(sX + ((size_t)b * a.maxn + i) * 4, make_float4(src[0], src[1], src[2], 0.f));
            }
            asm volatile("bar.sync 2, 96;" ::: "memory");
        };
        auto features = [&](const TileIter& t, int b, int tl) {
            for (int r = ft; r < 128; r += 96) write_features(sX + (size_t)b * a.maxn * 4, t, r >> 5, r & 31, sFeat + (tl & 1) * SM_FEAT1);
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&bars[tl & 1 ? B_FEAT1 : B_FEAT0]);
        };
        int xb = 0, s = 0;
        stage_x(it, 0);
        features(it, 0, 0);
        while (it.valid) {
            TileIter nx = it;
            iter_next<MULTI>(d, nx);
            // operand buffer (s+1)&1 was last read by the value-side angle MMA of tile s-1
            if (s >= 1) tc::mbar_wait_wd(&bars[(s - 1) & 1 ? B_PREV1 : B_PREV0], ((s - 1) >> 1) & 1);
            if (nx.valid) {
                if (nx.u != it.u) { xb ^= 1; stage_x(nx, xb); }
                features(nx, xb, s + 1);
            }
            it = nx; s++;
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= row warps: thread = (row, channel quarter) =================
        // warp w: rows 32*(w&3)..+31, channels [32*(w>>2), +32) of the key MLP and of the value MLP; heads 4*(w>>2)..+3
        // (four row warps per scheduler: enough independent work to cover the TMEM / shared-memory latencies)
        // Software pipeline across tiles (no warp ever waits for the tensor pipe in steady state):
        //   LN-k(t) | epilogue(t-1) | LN-v(t) | logits(t)      while the tensor pipe runs   W2k(t), angle-k(t+1) | W2v(t), angle-v(t+1)
        const int cq = warp >> 2;
        const bool fold[2] = {a.fold[0] > 0.5f, a.fold[1] > 0.5f};
        uint32_t psk = 0, psv = 0;
        int staged_k = -1, staged_v = -1;
        int tcount = 0;
        const int role = cq; (void)role;
        float al[4] = {0.f, 0.f, 0.f, 0.f};
        bool prev_valid = false;
        long long prev_eji = 0;
        // MULTI: on-line softmax state of the open segment (warp-uniform: running maximum and sum per head, the factor
        // the running output is rescaled by when the previous tile's chunk is added) and this lane's output channel
        float mrun[4] = {0.f, 0.f, 0.f, 0.f}, lrun[4] = {0.f, 0.f, 0.f, 0.f}, psc[4] = {0.f, 0.f, 0.f, 0.f}, oacc = 0.f;
        bool prev_first = true, prev_last = true;
        // pre-activation slice -> LayerNorm + ReLU -> bf16 hi/lo A operand of the second Linear
        auto layer_norm = [&](int mlp, uint32_t pre, uint32_t hid, int trow, const float* sR) {
            const int c0 = mlp * 128 + cq * 32;          // first of this thread's 32 channels inside the 256-wide (k|v) row
            float2 x2[16];
            {
                uint32_t xu[32];
                tc::tmem_ld32_nowait(pre + lane_base + cq * 32, xu);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) x2[i] = make_float2(__uint_as_float(xu[2 * i]), __uint_as_float(xu[2 * i + 1]));
            }
            float2 s1 = make_float2(0.f, 0.f), s2 = s1, s1b = s1, s2b = s1;
            const float* prow = sPs + ((size_t)mlp * psr + trow) * PS_HALF + cq * 32;
            const float* rrow = sR + wq * 256 + c0;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float4 p = ld4(prow + 2 * i), rr = ld4(rrow + 2 * i);
                x2[i] = tc::add2(x2[i], tc::add2(make_float2(p.x, p.y), make_float2(rr.x, rr.y)));
                x2[i + 1] = tc::add2(x2[i + 1], tc::add2(make_float2(p.z, p.w), make_float2(rr.z, rr.w)));
                s1 = tc::add2(s1, x2[i]); s1b = tc::add2(s1b, x2[i + 1]);
                s2 = tc::fma2(x2[i], x2[i], s2); s2b = tc::fma2(x2[i + 1], x2[i + 1], s2b);
            }
            s1 = tc::add2(s1, s1b); s2 = tc::add2(s2, s2b);
            // combine with the other three channel quarters of the same row (warps w +- 4k, same lane).  The barrier also
            // orders this lane quarter's reads of the previous tile's accumulators before the hid columns overwrite them.
            // quarter-major layout [mlp][quarter][row]: consecutive lanes touch consecutive 8-byte words (no bank conflicts)
            float* st = sStat + ((size_t)(mlp * 4) * 128 + wq * 32 + lane) * 2;
            *reinterpret_cast<float2*>(st + cq * 256) = make_float2(s1.x + s1.y, s2.x + s2.y);
            asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
            const float2 q0 = *reinterpret_cast<const float2*>(st), q1 = *reinterpret_cast<const float2*>(st + 256);
            const float2 q2 = *reinterpret_cast<const float2*>(st + 512), q3 = *reinterpret_cast<const float2*>(st + 768);
            const float mu = ((q0.x + q1.x) + (q2.x + q3.x)) * (1.0f / 128.0f);
            const float rstd = rsqrtf(fmaxf(fmaf(-mu, mu, ((q0.y + q1.y) + (q2.y + q3.y)) * (1.0f / 128.0f)), 0.f) + 1e-5f);
            const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mu * rstd, -mu * rstd);
            const float* gam = sLn + mlp * 256 + cq * 32;
            const float* bet = gam + 128;
            uint32_t hi[16], lo[16];
            if (KF16 && mlp == 0) {
                // key MLP: one fp16 value per activation
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 b4 = ld4(bet + 2 * i);
                    float2 y0 = tc::fma2(x2[i], rs2, nm2), y1 = tc::fma2(x2[i + 1], rs2, nm2);
                    if (fold[0]) {
                        y0 = tc::add2(y0, make_float2(b4.x, b4.y)); y1 = tc::add2(y1, make_float2(b4.z, b4.w));
                    } else {
                        const float4 g4 = ld4(gam + 2 * i);
                        y0 = tc::fma2(y0, make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
                        y1 = tc::fma2(y1, make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
                    }
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(y0.y), "f"(y0.x));
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[i + 1]) : "f"(y1.y), "f"(y1.x));
                }
                tc::tmem_st16(hid + lane_base + cq * 16, hi);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars[B_HIDK]);
                return;
            }
            if (fold[mlp]) {
                // gamma > 0 everywhere: it lives in the columns of W2, only beta / gamma is added here (half the broadcast loads)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 b4 = ld4(bet + 2 * i);
                    tc::split_pair_relu(tc::add2(tc::fma2(x2[i], rs2, nm2), make_float2(b4.x, b4.y)), hi[i], lo[i]);
                    tc::split_pair_relu(tc::add2(tc::fma2(x2[i + 1], rs2, nm2), make_float2(b4.z, b4.w)), hi[i + 1], lo[i + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 g4 = ld4(gam + 2 * i), b4 = ld4(bet + 2 * i);
                    float2 y0 = tc::fma2(tc::fma2(x2[i], rs2, nm2), make_float2(g4.x, g4.y), make_float2(b4.x, b4.y));
                    float2 y1 = tc::fma2(tc::fma2(x2[i + 1], rs2, nm2), make_float2(g4.z, g4.w), make_float2(b4.z, b4.w));
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
        // alpha-weighted sum of the value rows over the segment (lanes), residual add into h_bond
        auto epilogue = [&](uint32_t outV, uint32_t parity) {
            tc::mbar_wait(&bars[B_OUTV], parity);
            tc::tc_fence_after();
            uint32_t vu[32];
            tc::tmem_ld32_nowait(outV + lane_base + cq * 32, vu);
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
                if (prev_valid) a.hb[(size_t)prev_eji * 128 + c] += o + sB2[128 + c];     // sum(alpha) = 1; residual (uni_denoiser.py:285)
            } else {
                // channel c belongs to head cq*4 + (lane >> 3): rescale the running output by that head's factor, add the
                // chunk, and divide by the running sum once the segment's last chunk is in
                const int hs = lane >> 3;
                const float sc = hs == 0 ? psc[0] : hs == 1 ? psc[1] : hs == 2 ? psc[2] : psc[3];
                oacc = prev_first ? o : fmaf(oacc, sc, o);
                if (prev_valid && prev_last) {
                    const float l = hs == 0 ? lrun[0] : hs == 1 ? lrun[1] : hs == 2 ? lrun[2] : lrun[3];
                    a.hb[(size_t)prev_eji * 128 + c] += __fdividef(oacc, l) + sB2[128 + c];
                }
            }
        };
        while (it.valid) {
            const uint32_t ph = tcount & 1;
            const uint32_t preK = tmem + (ph ? 256 : 0), hidK = tmem + (ph ? 0 : 256);
            const uint32_t preV = tmem + (ph ? 384 : 128), hidV = tmem + (ph ? 128 : 384);
            const Seg sg = seg_of(it, wq);
            const int srow = it.chunk * 32 + lane;                  // row inside the segment
            const bool rowvalid = sg.valid && srow < it.n - 2;
            const int trow = rowvalid ? lane + (srow >= sg.ti) : 0;   // staged row: unit row (srow + skip of k = i) - 32 * chunk
            const float* sQ = sQR + (tcount & 1) * (SM_QR / 8);
            const float* sR = sQ + 4 * 128;
            TRACE(role, 0);
            // ---- key MLP
            if (staged_k != it.stage) { tc::mbar_wait(&bars[B_PSK], psk); psk ^= 1; staged_k = it.stage; }   // P rows (key half) landed
            tc::mbar_wait(&bars[B_PREK], ph);
            tc::tc_fence_after();
            TRACE(role, 1);
            layer_norm(0, preK, hidK, trow, sR);
            TRACE(role, 2);
            // ---- value epilogue of the previous tile (its W2v MMA ran during that tile's logits and the LayerNorm above)
            if (tcount > 0) epilogue(ph ? 128 + tmem : 384 + tmem, ph ^ 1);
            TRACE(role, 4);
            // ---- value MLP
            if (staged_v != it.stage) { tc::mbar_wait(&bars[B_PSV], psv); psv ^= 1; staged_v = it.stage; }
            tc::mbar_wait(&bars[ph ? B_PREV1 : B_PREV0], (tcount >> 1) & 1);
            tc::tc_fence_after();
            TRACE(role, 5);
            layer_norm(1, preV, hidV, trow, sR);
            TRACE(role, 6);
            // ---- logits of this thread's 4 heads, segment softmax across the 32 lanes (rows) of the warp
            {
                tc::mbar_wait(&bars[B_OUTK], ph);
                tc::tc_fence_after();
                TRACE(role, 7);
                uint32_t vv[32];
                tc::tmem_ld32_nowait(preK + lane_base + cq * 32, vv);
                tc::tmem_ld_wait();
                // the key bias b2k shifts all logits of a (segment, head) by the same q . b: softmax-invariant, dropped
                const float* qrow = sQ + wq * 128 + cq * 32;
                constexpr float kScale = kInvSqrtD * 1.4426950408889634f;      // logits in log2 units -> ex2 directly
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int o = h * 8;
                    const float4 qa = ld4(qrow + h * 8), qb = ld4(qrow + h * 8 + 4);
                    float2 s0 = tc::mul2(make_float2(qa.x, qa.y), make_float2(__uint_as_float(vv[o]), __uint_as_float(vv[o + 1])));
                    float2 s1 = tc::mul2(make_float2(qb.x, qb.y), make_float2(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5])));
                    s0 = tc::fma2(make_float2(qa.z, qa.w), make_float2(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3])), s0);
                    s1 = tc::fma2(make_float2(qb.z, qb.w), make_float2(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7])), s1);
                    s0 = tc::add2(s0, s1);
                    al[h] = rowvalid ? (s0.x + s0.y) * kScale : -INFINITY;
                }
                if (MULTI) {
                    // on-line softmax across the chunks of the segment: weights stay un-normalised (relative to the running
                    // maximum), the epilogue rescales the running output and divides by the running sum at the end
                    const bool first = it.chunk == 0;
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const float mold = first ? -INFINITY : mrun[h];
                        const float mnew = fmaxf(mold, tc::warp_max_redux(al[h]));
                        al[h] = rowvalid ? tc::ex2_approx(al[h] - mnew) : 0.f;
                        const float sc = mold == -INFINITY ? 0.f : tc::ex2_approx(mold - mnew);
                        const float ls = tc::warp_sum01_redux(al[h]);
                        lrun[h] = first ? ls : fmaf(lrun[h], sc, ls);
                        mrun[h] = mnew; psc[h] = sc;
                    }
                } else if (!(a.flags & 1)) {
                    // segment softmax across the 32 lanes: max and sum on the REDUX unit
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const float mx = tc::warp_max_redux(al[h]);
                        al[h] = rowvalid ? tc::ex2_approx(al[h] - mx) : 0.f;
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] *= __frcp_rn(tc::warp_sum01_redux(al[h]));
                } else {
                    float mx[4], sm[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) mx[h] = al[h];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int h = 0; h < 4; h++) mx[h] = fmaxf(mx[h], __shfl_xor_sync(PG_FULL, mx[h], o));
#pragma unroll
                    for (int h = 0; h < 4; h++) { al[h] = rowvalid ? tc::ex2_approx(al[h] - mx[h]) : 0.f; sm[h] = al[h]; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int h = 0; h < 4; h++) sm[h] += __shfl_xor_sync(PG_FULL, sm[h], o);
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] *= __frcp_rn(sm[h]);
                }
            }
            TRACE(role, 8);
            prev_valid = sg.valid; prev_eji = sg.eji;
            if (MULTI) { prev_first = it.chunk == 0; prev_last = it.chunk == it.nchunk - 1; }
            iter_next<MULTI>(d, it);
            tcount++;
        }
        // drain: value epilogue of the last tile
        if (tcount > 0) epilogue(((tcount - 1) & 1) ? 384 + tmem : 128 + tmem, (tcount - 1) & 1);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}

// Per-edge partials of the triplet MLPs' first Linear, for every bond edge e = (src -> dst):
//   P[e] = h_bond[e] Wb (edge GEMM output T) + h_src Whk + h_dst Whj + b1 + smear(|x_dst - x_src|) Wrkj   (edge in the k->j role)
//   R[e] = smear(|x_dst - x_src|) Wrji                                                                    (edge in the j->i role)
// Both k|v halves (256 channels).  P rows of the edges into one atom are contiguous -> one bulk copy per unit.
constexpr int PR_EDGES = 4;      // edges per warp: every weight row read from L1 is used for 4 edges
__global__ void __launch_bounds__(256, 2) trip_pr_kernel(TripTcArgs a) {
    const PlanDev& d = a.d;
    const int lane = threadIdx.x & 31;
    const long long e0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * PR_EDGES;
    if (e0 >= d.Eb) return;
    float mine[PR_EDGES];
    unsigned long long acc[PR_EDGES][8];
    // the warp issues in order: every dependent load level (indices -> coordinates / partial rows) is requested for all
    // edges of the warp before the first use, so the warp pays one memory round trip per level instead of one per edge
    long long ee[PR_EDGES];
    int sn[PR_EDGES], tn[PR_EDGES];
#pragma unroll
    for (int k = 0; k < PR_EDGES; k++) {
        ee[k] = min(e0 + k, d.Eb - 1);
        sn[k] = d.esrc_node[ee[k]]; tn[k] = d.edst_node[ee[k]];
    }
    float4 tk[PR_EDGES], tv[PR_EDGES];
#pragma unroll
    for (int k = 0; k < PR_EDGES; k++) {
        tk[k] = ldg4(a.T + (size_t)ee[k] * a.ldt + a.t_k + lane * 4);
        tv[k] = ldg4(a.T + (size_t)ee[k] * a.ldt + a.t_v + lane * 4);
    }
    float xs[PR_EDGES][3], xt[PR_EDGES][3];
#pragma unroll
    for (int k = 0; k < PR_EDGES; k++) {
#pragma unroll
        for (int c = 0; c < 3; c++) { xs[k][c] = a.x[(size_t)sn[k] * 3 + c]; xt[k][c] = a.x[(size_t)tn[k] * 3 + c]; }
    }
#pragma unroll
    for (int k = 0; k < PR_EDGES; k++) {
        const float4 hs0 = ldg4(a.H + (size_t)sn[k] * a.ldh + a.hk_k + lane * 4), ht0 = ldg4(a.H + (size_t)tn[k] * a.ldh + a.hj_k + lane * 4);
        const float4 hs1 = ldg4(a.H + (size_t)sn[k] * a.ldh + a.hk_v + lane * 4), ht1 = ldg4(a.H + (size_t)tn[k] * a.ldh + a.hj_v + lane * 4);
        const float d0 = xt[k][0] - xs[k][0], d1 = xt[k][1] - xs[k][1], d2 = xt[k][2] - xs[k][2];
        const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        mine[k] = lane < 20 ? smear_val(dist, lane) : 0.f;
        const float4 p0 = f4add(f4add(tk[k], hs0), ht0);
        const float4 p1 = f4add(f4add(tv[k], hs1), ht1);
        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0ull;
        acc[k][4] = pk2(p0.x, p0.y); acc[k][5] = pk2(p0.z, p0.w); acc[k][6] = pk2(p1.x, p1.y); acc[k][7] = pk2(p1.z, p1.w);
    }
    // accumulators stay packed (two fp32 per 64-bit register pair) through the 20-term loop (FFMA2)
#pragma unroll 2
    for (int gg = 0; gg < 20; gg++) {
        const float4 w0 = ldg4(a.wrji + gg * 256 + lane * 4), w1 = ldg4(a.wrji + gg * 256 + 128 + lane * 4);
        const float4 w2 = ldg4(a.wrkj + gg * 256 + lane * 4), w3 = ldg4(a.wrkj + gg * 256 + 128 + lane * 4);
        const unsigned long long u0 = pk2(w0.x, w0.y), u1 = pk2(w0.z, w0.w), u2 = pk2(w1.x, w1.y), u3 = pk2(w1.z, w1.w);
        const unsigned long long u4 = pk2(w2.x, w2.y), u5 = pk2(w2.z, w2.w), u6 = pk2(w3.x, w3.y), u7 = pk2(w3.z, w3.w);
#pragma unroll
        for (int k = 0; k < PR_EDGES; k++) {
            const float sg = __shfl_sync(PG_FULL, mine[k], gg);
            const unsigned long long ss = pk2(sg, sg);
            acc[k][0] = fma2_raw(ss, u0, acc[k][0]); acc[k][1] = fma2_raw(ss, u1, acc[k][1]);
            acc[k][2] = fma2_raw(ss, u2, acc[k][2]); acc[k][3] = fma2_raw(ss, u3, acc[k][3]);
            acc[k][4] = fma2_raw(ss, u4, acc[k][4]); acc[k][5] = fma2_raw(ss, u5, acc[k][5]);
            acc[k][6] = fma2_raw(ss, u6, acc[k][6]); acc[k][7] = fma2_raw(ss, u7, acc[k][7]);
        }
    }
    auto f4 = [](unsigned long long lo, unsigned long long hi) { const float2 x = up2(lo), y = up2(hi); return make_float4(x.x, x.y, y.x, y.y); };
#pragma unroll
    for (int k = 0; k < PR_EDGES; k++) {
        const long long e = e0 + k;
        if (e >= d.Eb) break;
        st4(a.R + (size_t)e * 256 + lane * 4, f4(acc[k][0], acc[k][1]));
        st4(a.R + (size_t)e * 256 + 128 + lane * 4, f4(acc[k][2], acc[k][3]));
        // P: key / value halves in separate arrays of padded rows (the staging layout of trip_tc_kernel)
        st4(a.P + (size_t)e * PS_HALF + lane * 4, f4(acc[k][4], acc[k][5]));
        st4(a.P + (size_t)d.Eb * PS_HALF + (size_t)e * PS_HALF + lane * 4, f4(acc[k][6], acc[k][7]));
    }
}
}  // namespace

int pg_launch_trip_pr(const TripTcArgs& a, cudaStream_t s) {
    if (a.d.Eb <= 0) return PG_OK;
    trip_pr_kernel<<<(unsigned)((a.d.Eb + 8 * PR_EDGES - 1) / (8 * PR_EDGES)), 256, 0, s>>>(a);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

size_t pg_trip_tc_smem(int maxn) { return (size_t)SM_FIXED + ((size_t)2 * ps_rows(maxn) * PS_HALF + 2 * (size_t)maxn * 4) * sizeof(float) + 1024; }

int pg_launch_trip_tc(const TripTcArgs& a, int num_sms, cudaStream_t s) {
    if (a.d.Nl <= 0) return PG_OK;
    const size_t smem = pg_trip_tc_smem(a.maxn);
    if (smem > 227 * 1024) { pg_set_error("trip_tc: shared memory budget exceeded"); return PG_ELIMIT; }
    static size_t cur = 0;
    if (smem > cur) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    const unsigned grid = (unsigned)std::min<long long>(a.d.Nl, num_sms);
    // The kernel a molecule runs on depends on the molecule alone (n - 2 <= 32: single-chunk instantiation, else the chunked
    // one), never on the batch it is in: results are bit-identical however a job is batched or sharded.  A mixed batch
    // launches both; each walks the unit list and skips the other's molecules.
    const bool have_single = a.d.min_n <= PG_TRIP_TC_SINGLE_CHUNK_ATOMS, have_multi = a.maxn > PG_TRIP_TC_SINGLE_CHUNK_ATOMS;
    if (a.flags & 2) {
        if (have_single) trip_tc_kernel<false, false><<<grid, NTHREADS, smem, s>>>(a);
        if (have_multi) trip_tc_kernel<true, false><<<grid, NTHREADS, smem, s>>>(a);
    } else {
        if (have_single) trip_tc_kernel<false, true><<<grid, NTHREADS, smem, s>>>(a);
        if (have_multi) trip_tc_kernel<true, true><<<grid, NTHREADS, smem, s>>>(a);
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

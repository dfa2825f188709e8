// BondUpdateLayer (uni_denoiser.py:123-165) on the 5th-gen tensor cores.
//
// One persistent CTA per SM (640 threads) walks the ligand atoms j ("units").  For a unit the per-edge partials
// P[k->j] (n-1 rows x 256 channels, k|v; written by trip_pr_kernel below as bf16 hi/lo operand images) are staged once in
// shared memory by bulk copies; the n-1 segments (j->i) are processed four at a time as a 128-row tile: TMEM lane =
// triplet row, lane l of every segment <-> staged row l (the edge k->j with the l-th k); the lane with k = i is masked
// out of the softmax.  The whole first Linear is assembled by the tensor pipe: pre-activation = [angle features | segment
// indicator] x [Wa ; R rows of the tile] + one-hot(l) x P rows, so the row warps never touch P or R.
//
// Roles: 16 row warps, thread = (row, 32-channel quarter), warp w -> lane quarter w & 3, channel quarter w >> 2 (four row
//        warps per scheduler); warp 16 issues every MMA and does nothing else; warp 17 requests the next tile's query rows
//        (cp.async), R rows and P images (bulk copies); warps 18-19 compute the angular features of the tiles ahead.
//        setmaxnreg moves registers from the auxiliary warpgroup to the row warps (104 / 64).
//
//   1. angular encoding (11 distinct values per triplet; sin/cos of the angle itself occur twice in the reference's encoding
//      and their weight rows are summed at pack time) + a 1.0 in column 11 + segment slot -> bf16 hi/lo A tile in smem
//      (double buffered); per MLP three tcgen05.mma (M128 N128 K16, bf16x3) against [angle slice of the first Linear ;
//      R[j->i] rows of the tile's 4 segments] (B operand MN-major, rows 11..14 refilled per tile by bulk copies), then
//      two K16 slabs of a constant one-hot A operand against the staged P rows (hi and lo images: 4 MMAs)
//      -> pre-activation columns in TMEM.
//   2. every thread reads its 32-channel slice (tcgen05.ld), applies
//      LayerNorm + ReLU in registers with packed fp32x2 math (the four quarter statistics of a row meet in shared memory),
//      splits to bf16 hi/lo with the ReLU fused into the conversions and writes the A operand of the second Linear back to
//      TMEM (tcgen05.st).  Positive LayerNorm gains are folded into W2 at pack time (weights.py), so only beta/gamma is added.
//   3. second Linear of the key and value MLPs: A from TMEM, B = W2 (bf16 hi/lo, resident in smem, 128B swizzle),
//      24 tcgen05.mma (M128 N128 K16) each, fp32 accumulators in TMEM.
//   4. logits = q . k per head (thread-local dot; the key bias is softmax-invariant and dropped), segment softmax across
//      the 32 lanes on the REDUX unit, alpha-weighted sum of v over the lanes (through TMEM: the weighted rows return in
//      the 16x256b fragment shape, 7 shuffles; butterfly transpose-reduce in the chunked instantiation), residual add
//      into h_bond.
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
// Operand images of the first Linear's per-edge parts (written by trip_pr_kernel, bulk-copied verbatim into shared memory):
//   P[k->j], per MLP: for every unit a block [hi|lo][channel half][n-1 rows][64 bf16 = 128 B], the 16-byte chunks of a row
//   XOR-swizzled with (row & 7): the MN-major / 128-byte-swizzle canonical layout of a UMMA B operand (K = staged row,
//   N = channel), so rows 16s..16s+15 of the staged block are the B operand of K16 slab s.
//   R[j->i], per MLP: same block shape, indexed by the SOURCE atom's unit and the segment index of i, swizzled for rows
//   11 + (segment & 3) of the angle slab: the 4 segments of a tile are one contiguous 512-byte piece per (hi|lo, half).
constexpr int W_TILE = 32768;              // one [128 x 128] bf16 matrix in two 128B-swizzled K blocks
constexpr int SM_W = 4 * W_TILE;           // (k,v) x (hi,lo)
constexpr int SM_WA = 2 * 2 * 4096;        // angle slab B operand [mlp][hi|lo][half][16 rows][128 B], MN-major SW128: rows 0..10 weights, 11..14 R
constexpr int SM_P = 2 * 2 * 8192;         // staged P rows [mlp][hi|lo][half][32 rows][128 B], MN-major SW128
constexpr int SM_ONE = 2 * 4096;           // constant one-hot A operand: two K16 slabs [128 x 16] bf16, no swizzle
constexpr int SM_FEAT1 = 2 * 4096;         // (hi,lo) x [128 x 16] bf16, no swizzle
constexpr int SM_FEAT = 2 * SM_FEAT1;      // double-buffered: the feature warps run up to two tiles ahead
constexpr int SM_Q = 2 * 4 * 128 * 4;      // double-buffered query rows of a tile's 4 segments
constexpr int SM_STAT = 128 * 8 * 4;       // LayerNorm partial second moments [mlp][quarter][row]
constexpr int SM_FIXED = SM_W + SM_WA + SM_P + SM_ONE + SM_FEAT + SM_Q + SM_STAT + 6 * 128 * 4 /*ln + b2*/ + 256 /*barriers*/;
constexpr float kInvSqrtD = 0.35355339059327373f;
constexpr int ROW_WARPS = 16;               // 4 warps per TMEM lane quarter, each owning a 32-channel slice of the row
constexpr int MMA_WARP = ROW_WARPS;         // MMA issue; then the loader warp (q / R / P rows) and 2 warps that compute angular features
constexpr int LOAD_WARP = MMA_WARP + 1;
constexpr int NTHREADS = (ROW_WARPS + 4) * 32;
constexpr int ROW_THREADS = ROW_WARPS * 32;
constexpr int R_ROW0 = 11;                  // angle slab: K columns 0..10 features, 11..14 segment indicator, 15 zero

// mbarrier slots
// A parity wait is only safe while the barrier cannot complete a second time before the waiter looks at it.  B_PREV has
// two kinds of waiters (row warps and the feature warps, which run ahead on their own clock), so it alternates between
// two barriers by tile parity, like B_FEAT.
enum { B_FEAT0 = 0, B_FEAT1, B_PREK, B_PREV0, B_PREV1, B_HIDK, B_HIDV, B_OUTK, B_OUTV, B_PSK, B_PSV, B_RK, B_RV, B_COUNT };

// the sequence of (unit, tile) a CTA walks; every role steps through it redundantly
// MULTI: a unit with more than 32 rows k -> j is cut into chunks of 32 rows; a tile is then (group of 4 segments) x (chunk c),
// the chunks of a group on consecutive tiles (chunk-inner order), and the softmax runs on-line across them.
// `stage` counts the refills of the staged P rows: once per unit when the unit is a single chunk (all n-1 rows stay
// resident), once per tile otherwise (the rows of chunk c).
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
        if (n >= 3 && (MULTI ? n - 1 > 32 : n - 1 <= 32)) {       // each instantiation takes its own molecules (see the launcher)
            it.n = n; it.jl = it.u - d.lig_off[g]; it.ctx0 = d.ctx_off[g] + d.g_p[g]; it.eoff = d.eoff[g];
            it.nchunk = MULTI ? (n - 1 + 31) >> 5 : 1;
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
// `wq` = segment slot of the tile (0..3), `row` = row inside the chunk (0..31) = staged row of the edge k -> j
// K columns: 0 theta, 1..3 sin(theta * {1,2,3}), 4..5 sin(theta / {2,3}), 6..8 cos(theta * {1,2,3}), 9..10 cos(theta / {2,3})
// (sin theta and cos theta occur twice in the reference's 13 values: weights.py adds their weight rows), 11 + wq = 1
// (selects the tile's R[j->i] row in the B operand).
__device__ __forceinline__ void write_features(const float* xs, const TileIter& it, int wq, int row, uint8_t* sFeat) {
    const int lane = row;
    const Seg sg = seg_of(it, wq);
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = 0.f;
    const int srow = it.chunk * 32 + lane;          // row of the unit: k is the srow-th atom other than j
    if (sg.valid && srow < it.n - 1 && srow != sg.ti) {
        const float4 xj = ld4(xs + it.jl * 4), xi = ld4(xs + sg.il * 4), xk = ld4(xs + (srow + (srow >= it.jl)) * 4);
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
        f[0] = th; f[1] = s1; f[2] = s2; f[3] = s3; f[4] = sh; f[5] = st;
        f[6] = k1; f[7] = k2; f[8] = k3; f[9] = kh; f[10] = kt;
    }
    f[11] = wq == 0 ? 1.f : 0.f; f[12] = wq == 1 ? 1.f : 0.f; f[13] = wq == 2 ? 1.f : 0.f; f[14] = wq == 3 ? 1.f : 0.f;
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
//   KMODE 2 (PG_KEY=trip16x2) = the activations as an fp16 hi/lo pair (truncated hi, rounded remainder: ~2^-22) against the
//   single fp16 weight image: 16 MMAs.  The activation rounding was what cost KMODE 1 the internal-state bar; the weight
//   rounding is common to all rows of a segment's softmax.
template <bool MULTI, int KMODE>
__global__ void __launch_bounds__(NTHREADS, 1) trip_tc_kernel(TripTcArgs a) {
    constexpr bool KF16 = KMODE != 0, KX2 = KMODE == 2;
    extern __shared__ uint8_t smem_raw[];
    // offset arithmetic on the __shared__ array (not a uintptr_t round trip) so the compiler keeps emitting LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    uint8_t* sWa = sW + SM_W;                       // 1024-aligned (swizzled operand)
    uint8_t* sP = sWa + SM_WA;                      // 1024-aligned (swizzled operand)
    uint8_t* sOne = sP + SM_P;
    uint8_t* sFeat = sOne + SM_ONE;
    float* sQ0 = (float*)(sFeat + SM_FEAT);         // [2][4 x 128] query rows
    float* sStat = sQ0 + SM_Q / 4;                  // [2 mlp][4 quarters][128 rows] partial second moments of the LayerNorm (quarter-major: conflict-free)
    float* sLn = sStat + SM_STAT / 4;               // gk, bk, gv, bv
    float* sB2 = sLn + 4 * 128;                     // b2k, b2v
    uint64_t* bars = (uint64_t*)(sB2 + 2 * 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + B_COUNT);
    float* sX = (float*)(smem + SM_FIXED);          // [2][maxn][4] ligand coordinates of the current / next unit
    const PlanDev& d = a.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the role branches and the MMA issuer's descriptor arithmetic stay on the uniform datapath
    const int wq = warp & 3;

    if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        tc::mbar_init(&bars[B_FEAT0], 64 + 32); tc::mbar_init(&bars[B_FEAT1], 64 + 32);     // 64 feature threads + the loader warp (query rows)
        tc::mbar_init(&bars[B_PREK], 1); tc::mbar_init(&bars[B_PREV0], 1); tc::mbar_init(&bars[B_PREV1], 1);
        tc::mbar_init(&bars[B_HIDK], ROW_THREADS); tc::mbar_init(&bars[B_HIDV], ROW_THREADS);
        tc::mbar_init(&bars[B_OUTK], 1); tc::mbar_init(&bars[B_OUTV], 1);
        tc::mbar_init(&bars[B_PSK], 1); tc::mbar_init(&bars[B_PSV], 1);
        tc::mbar_init(&bars[B_RK], 1); tc::mbar_init(&bars[B_RV], 1);
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
    // angle slab image (weights.py: already in the shared-memory layout, R rows and row 15 zero)
    for (int idx = tid; idx < SM_WA / 16; idx += NTHREADS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(sWa) + idx * 16), "l"(a.wa_bf + idx * 8) : "memory");
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    // staged P rows start as zeros (rows beyond n-1 are selected by padding lanes: finite garbage is harmless, NaN is not)
    for (int idx = tid; idx < SM_P / 16; idx += NTHREADS) *reinterpret_cast<uint4*>(sP + idx * 16) = make_uint4(0u, 0u, 0u, 0u);
    // constant one-hot A operand: row r selects staged row (r & 31); slab s covers K columns 16 s .. 16 s + 15
    for (int idx = tid; idx < 2 * 128 * 2; idx += NTHREADS) {     // [slab][row][k chunk of 8]
        const int sl = idx >> 8, r = (idx >> 1) & 127, kc = idx & 1;
        const int one = (r & 31) - sl * 16 - kc * 8;                // position of the 1.0 inside this 8-element chunk, if any
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (one >= 0 && one < 8) w[one >> 1] = (one & 1) ? 0x3F800000u : 0x00003F80u;
        *reinterpret_cast<uint4*>(sOne + sl * 4096 + (r >> 3) * 256 + kc * 128 + (r & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
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
        // ================= MMA issue + loaders (query rows via cp.async, P / R rows via bulk copies) =================
        constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_bmn = idesc | (1u << 16);          // B operand MN-major (staged P rows, angle slab)
        const uint32_t sW_u32 = tc::smem_u32(sW), sWa_u32 = tc::smem_u32(sWa), sFeat_u32 = tc::smem_u32(sFeat);
        const uint32_t sP_u32 = tc::smem_u32(sP), sOne_u32 = tc::smem_u32(sOne);
        // first Linear of one MLP for tile `tl` (feature operand buffer tl & 1) -> pre-activation columns `dcol`:
        // [angle | indicator] x [Wa ; R] in bf16x3, then one-hot x staged P rows (hi and lo images), `nslab` K16 slabs
        // (warp-collective issue: all lanes run the descriptor arithmetic on the uniform datapath, one elected lane issues)
        auto feat_mma = [&](int tl, int mlp, int nslab, uint32_t dcol, uint64_t* bar) {
            const uint32_t fb = sFeat_u32 + (tl & 1) * SM_FEAT1;
            const uint64_t fh = tc::umma_desc_k16_noswizzle(fb), fl = tc::umma_desc_k16_noswizzle(fb + 4096);
            const uint64_t wh = tc::umma_desc_mn_sw128(sWa_u32 + (mlp * 2) * 4096, 2048), wl = tc::umma_desc_mn_sw128(sWa_u32 + (mlp * 2 + 1) * 4096, 2048);
            tc::umma_bf16_w(dcol, fh, wh, idesc_bmn, 0);
            tc::umma_bf16_w(dcol, fh, wl, idesc_bmn, 1);
            tc::umma_bf16_w(dcol, fl, wh, idesc_bmn, 1);
            for (int sl = 0; sl < nslab; sl++) {
                const uint64_t oh = tc::umma_desc_k16_noswizzle(sOne_u32 + sl * 4096);
                tc::umma_bf16_w(dcol, oh, tc::umma_desc_mn_sw128(sP_u32 + (mlp * 2) * 8192 + sl * 2048, 4096), idesc_bmn, 1);
                tc::umma_bf16_w(dcol, oh, tc::umma_desc_mn_sw128(sP_u32 + (mlp * 2 + 1) * 8192 + sl * 2048, 4096), idesc_bmn, 1);
            }
            tc::umma_commit_w(bar);
        };
        auto slabs_of = [](const TileIter& t) { return (min(32, t.n - 1 - t.chunk * 32) + 15) >> 4; };
        // second Linear of one MLP: A = bf16 hi/lo activations in TMEM columns `hid`, D = `dcol`
        auto w2_mma = [&](int mlp, uint32_t hid, uint32_t dcol, uint64_t* bar) {
            if (KF16 && mlp == 0) {
                constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A, B = F16
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint64_t bd = tc::umma_desc_sw128(sW_u32 + (ks >> 2) * 16384 + (ks & 3) * 32);
                    tc::umma_bf16_ts_w(dcol, hid + ks * 8, bd, idesc16, ks > 0);
                }
                if (KX2) {
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t bd = tc::umma_desc_sw128(sW_u32 + (ks >> 2) * 16384 + (ks & 3) * 32);
                        tc::umma_bf16_ts_w(dcol, hid + 64 + ks * 8, bd, idesc16, 1);
                    }
                }
                tc::umma_commit_w(bar);
                return;
            }
            uint32_t acc = 0;
#pragma unroll
            for (int combo = 0; combo < 3; combo++) {       // hi*hi, hi*lo, lo*hi
                const uint32_t abase = hid + (combo == 2 ? 64 : 0);
                const uint32_t bbase = sW_u32 + (mlp * 2 + (combo == 1 ? 1 : 0)) * W_TILE;
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint64_t bd = tc::umma_desc_sw128(bbase + (ks >> 2) * 16384 + (ks & 3) * 32);
                    tc::umma_bf16_ts_w(dcol, abase + ks * 8, bd, idesc, acc);
                    acc = 1;
                }
            }
            tc::umma_commit_w(bar);
        };
        uint32_t psk = 0, psv = 0, prk = 0, prv = 0;               // phases of the P / R transaction barriers
        tc::mbar_wait_wd(&bars[B_FEAT0], 0);
        tc::mbar_wait_wd(&bars[B_PSK], psk); psk ^= 1;
        tc::mbar_wait_wd(&bars[B_RK], prk); prk ^= 1;
        tc::tc_fence_after();
        feat_mma(0, 0, slabs_of(it), tmem + 0, &bars[B_PREK]);
        tc::mbar_wait_wd(&bars[B_PSV], psv); psv ^= 1;
        tc::mbar_wait_wd(&bars[B_RV], prv); prv ^= 1;
        tc::tc_fence_after();
        feat_mma(0, 1, slabs_of(it), tmem + 128, &bars[B_PREV0]);
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
            // key activations of tile t are in TMEM: its first-Linear MMAs have completed (the P_k image and the R_k rows
            // are free), and the logits of tile t-1 are done
            tc::mbar_wait_wd(&bars[B_HIDK], ph);
            TRACE(2, 1);
            tc::tc_fence_after();
            w2_mma(0, hidK, preK, &bars[B_OUTK]);
            TRACE(2, 2);
            if (nx.valid) {
                tc::mbar_wait_wd(&bars[B_OUTK], ph);      // hid_k(t) has been consumed: its columns take pre_k(t+1)
                TRACE(2, 9);
                tc::mbar_wait_wd(featbar, ((tcount + 1) >> 1) & 1);
                TRACE(2, 10);
                tc::mbar_wait_wd(&bars[B_RK], prk); prk ^= 1;
                TRACE(2, 11);
                if (newu) { tc::mbar_wait_wd(&bars[B_PSK], psk); psk ^= 1; }
                tc::tc_fence_after();
                TRACE(2, 3);
                feat_mma(tcount + 1, 0, slabs_of(nx), hidK, &bars[B_PREK]);
            }
            TRACE(2, 4);
            tc::mbar_wait_wd(&bars[B_HIDV], ph);
            TRACE(2, 5);
            tc::tc_fence_after();
            w2_mma(1, hidV, preV, &bars[B_OUTV]);
            TRACE(2, 6);
            if (nx.valid) {
                tc::mbar_wait_wd(&bars[B_OUTV], ph);
                TRACE(2, 12);
                tc::mbar_wait_wd(&bars[B_RV], prv); prv ^= 1;
                TRACE(2, 13);
                if (newu) { tc::mbar_wait_wd(&bars[B_PSV], psv); psv ^= 1; }
                tc::tc_fence_after();
                TRACE(2, 7);
                feat_mma(tcount + 1, 1, slabs_of(nx), hidV, &bars[(tcount + 1) & 1 ? B_PREV1 : B_PREV0]);
            }
            TRACE(2, 8);
            it = nx; tcount++;
        }
      } else if (warp == LOAD_WARP) {
        // ================= loader: query rows via cp.async, P / R rows via bulk copies =================
        // A warp of its own: on the MMA warp these requests (several hundred cycles of dependent address arithmetic per tile)
        // either delayed the MMA issue (requests first: tensor pipe idle meanwhile) or were delayed by it (MMAs first: the
        // issue blocks ~1500 cycles on the full MMA queue, 37.9 -> 40.4 ms/step).  The buffers of tile t+1 are free when the
        // row warps hand back tile t's activations (HIDK / HIDV: the first-Linear MMAs of tile t have completed); the next
        // completion of either barrier needs the rows requested here, so the parity waits cannot be overtaken.
        const size_t mlp_stride = (size_t)a.d.Eb * 128;           // floats per MLP in the P and R images (512 B per edge)
        auto load_q = [&](const TileIter& t, int buf) {
            const uint32_t q = tc::smem_u32(sQ0 + buf * (SM_Q / 8));
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const Seg sg = seg_of(t, s4);
                const float* qs = a.q + (size_t)sg.eji * 128 + lane * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(q + (s4 * 128 + lane * 4) * 4), "l"(qs) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // P rows of the edges k -> j of a unit (or of one 32-row chunk of it): four contiguous pieces per MLP, (hi|lo) x
        // (channel half), each nr x 128 B.  The key and value images travel separately (one mbarrier transaction each): an
        // image is dead as soon as the first-Linear MMAs of the unit's last tile have completed, which the issuer knows
        // when the row warps hand back that tile's activations (HIDK / HIDV).
        auto load_ps = [&](const TileIter& t, int mlp) {
            uint64_t* bar = &bars[mlp == 0 ? B_PSK : B_PSV];
            const int r0 = t.chunk * 32, nr = min(32, t.n - 1 - r0);
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(bar, (uint32_t)nr * 512u);
                const uint8_t* src = (const uint8_t*)(a.P + (size_t)mlp * mlp_stride) + (size_t)(t.eoff + (long long)t.jl * (t.n - 1)) * 512;
#pragma unroll
                for (int b = 0; b < 4; b++)
                    tc::bulk_copy_g2s(sP + (mlp * 4 + b) * 4096, src + ((size_t)b * (t.n - 1) + r0) * 128, (uint32_t)nr * 128u, bar);
            }
            __syncwarp();
        };
        // R rows of the tile's 4 segments (j -> i), contiguous in the source-major R image -> rows 11..14 of the angle slab
        auto load_r = [&](const TileIter& t, int mlp) {
            uint64_t* bar = &bars[mlp == 0 ? B_RK : B_RV];
            const int s0 = t.grp * 4, nr = min(4, t.n - 1 - s0);
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(bar, (uint32_t)nr * 512u);
                const uint8_t* src = (const uint8_t*)(a.R + (size_t)mlp * mlp_stride) + (size_t)(t.eoff + (long long)t.jl * (t.n - 1)) * 512;
#pragma unroll
                for (int b = 0; b < 4; b++)
                    tc::bulk_copy_g2s(sWa + (mlp * 4 + b) * 2048 + 1024 + (R_ROW0 - 8) * 128, src + ((size_t)b * (t.n - 1) + s0) * 128,
                                      (uint32_t)nr * 128u, bar);
            }
            __syncwarp();
        };
        load_ps(it, 0);
        load_ps(it, 1);
        load_r(it, 0);
        load_r(it, 1);
        load_q(it, 0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        tc::mbar_arrive(&bars[B_FEAT0]);
        int tcount = 0;
        while (it.valid) {
            const uint32_t ph = tcount & 1;
            TileIter nx = it;
            iter_next<MULTI>(d, nx);
            const bool newu = nx.valid && nx.stage != it.stage;     // the staged P rows change with the next tile
            tc::mbar_wait_wd_sleep(&bars[B_HIDK], ph);              // the P_k image, the R_k rows and query buffer (t+1)&1 are free
            if (nx.valid) { load_q(nx, (tcount + 1) & 1); load_r(nx, 0); }
            if (newu) load_ps(nx, 0);
            if (nx.valid) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                tc::mbar_arrive(&bars[(tcount + 1) & 1 ? B_FEAT1 : B_FEAT0]);
            }
            tc::mbar_wait_wd_sleep(&bars[B_HIDV], ph);
            if (nx.valid) load_r(nx, 1);
            if (newu) load_ps(nx, 1);
            it = nx; tcount++;
        }
      } else {
        // ================= feature warps: angular features, up to two tiles ahead (double-buffered operand) =================
        // 64 threads cover the 128 rows of a tile in two passes; they also stage the next unit's coordinates.
        const int ft = tid - (LOAD_WARP + 1) * 32;       // 0..63
        auto stage_x = [&](const TileIter& t, int b) {
            for (int i = ft; i < t.n; i += 64) {
                const float* src = a.x + (size_t)(t.ctx0 + i) * 3;
                st4(sX + ((size_t)b * a.maxn + i) * 4, make_float4(src[0], src[1], src[2], 0.f));
            }
            asm volatile("bar.sync 2, 64;" ::: "memory");
        };
        auto features = [&](const TileIter& t, int b, int tl) {
            for (int r = ft; r < 128; r += 64) write_features(sX + (size_t)b * a.maxn * 4, t, r >> 5, r & 31, sFeat + (tl & 1) * SM_FEAT1);
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
            if (s >= 1) tc::mbar_wait_wd_sleep(&bars[(s - 1) & 1 ? B_PREV1 : B_PREV0], ((s - 1) >> 1) & 1);
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
        // (the tensor pipe has already summed the angle, P[k->j] and R[j->i] parts of the first Linear)
        auto layer_norm = [&](int mlp, uint32_t pre, uint32_t hid) {
            float2 x2[16];
            {
                uint32_t xu[32];
                tc::tmem_ld32_nowait(pre + lane_base + cq * 32, xu);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) x2[i] = make_float2(__uint_as_float(xu[2 * i]), __uint_as_float(xu[2 * i + 1]));
            }
            // The pre-activation arrives MEAN-FREE: every operand of the first Linear (Wa, P, R) has its channel mean removed
            // (weights._center_first_linears), so the LayerNorm only needs the second moment: var = sum x^2 / 128.
            float2 s2 = make_float2(0.f, 0.f), s2b = s2;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                s2 = tc::fma2(x2[i], x2[i], s2); s2b = tc::fma2(x2[i + 1], x2[i + 1], s2b);
            }
            s2 = tc::add2(s2, s2b);
            // combine with the other three channel quarters of the same row (warps w +- 4k, same lane).  The barrier also
            // orders this lane quarter's reads of the previous tile's accumulators before the hid columns overwrite them.
            // quarter-major layout [mlp][quarter][row]: consecutive lanes touch consecutive words (no bank conflicts)
            float* st = sStat + (size_t)(mlp * 4) * 128 + wq * 32 + lane;
            st[cq * 128] = s2.x + s2.y;
            asm volatile("bar.sync %0, 128;" ::"r"(3 + wq) : "memory");
            const float rstd = rsqrtf(((st[0] + st[128]) + (st[256] + st[384])) * (1.0f / 128.0f) + 1e-5f);
            const float2 rs2 = make_float2(rstd, rstd);
            const float* gam = sLn + mlp * 256 + cq * 32;
            const float* bet = gam + 128;
            uint32_t hi[16], lo[16];
            if (KF16 && mlp == 0) {
                // key MLP: one fp16 value per activation
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
                    if (KX2) {
                        tc::split_pair_relu_f16(y0, hi[i], lo[i]);
                        tc::split_pair_relu_f16(y1, hi[i + 1], lo[i + 1]);
                    } else {
                        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(y0.y), "f"(y0.x));
                        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[i + 1]) : "f"(y1.y), "f"(y1.x));
                    }
                }
                tc::tmem_st16(hid + lane_base + cq * 16, hi);
                if (KX2) tc::tmem_st16(hid + lane_base + 64 + cq * 16, lo);
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
        // alpha-weighted sum of the value rows over the segment (lanes), residual add into h_bond
        auto epilogue = [&](uint32_t outV, uint32_t parity) {
            tc::mbar_wait(&bars[B_OUTV], parity);      // (also orders W2v(t-1)'s writes before this thread's next hid_v rows)
            tc::tc_fence_after();
            if (!prev_valid) return;                   // no segment in this lane quarter of the previous tile
            uint32_t vu[32];
            tc::tmem_ld32_nowait(outV + lane_base + cq * 32, vu);
            tc::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {                                           // alpha is 0 on padded rows
                const float2 pr = tc::mul2(make_float2(al[i >> 3], al[i >> 3]), make_float2(__uint_as_float(vu[i]), __uint_as_float(vu[i + 1])));
                v[i] = pr.x; v[i + 1] = pr.y;
            }
#if defined(PG_TRIP_EXP) && (PG_TRIP_EXP & 1)
            float o = 0.f;                                   // knock-out experiment: no cross-lane reduction (wrong results)
#pragma unroll
            for (int i = 0; i < 32; i++) o += v[i];
            const int cl = lane;
#elif defined(PG_TRIP_SHFL_EPILOGUE)
            const float o = transpose_reduce32(v, lane);
            const int cl = lane;
#else
            // the weighted rows go back through the accumulator columns they came from (this lane quarter's own 32 columns:
            // free until the next value LayerNorm's barrier) and return in the MMA-fragment shape: 7 shuffles instead of 31
            // (default single-chunk instantiation only: the chunked one carries the on-line softmax state and would spill)
            constexpr bool SHFL = MULTI || KF16;
            const float o = SHFL ? transpose_reduce32(v, lane) : tc::tmem_transpose_reduce32(outV + lane_base + cq * 32, v, lane);
            const int cl = SHFL ? lane : tc::tmem_reduce_col(lane);        // cl >> 3 == lane >> 3: the lane's head does not change
#endif
            const int c = cq * 32 + cl;
            if (!MULTI) {
#if defined(PG_TRIP_EXP) && (PG_TRIP_EXP & 4)
                if (prev_valid && o == 123.456f) a.hb[(size_t)prev_eji * 128 + c] += o + sB2[128 + c];   // knock-out: no h_bond update
#else
                if (prev_valid) a.hb[(size_t)prev_eji * 128 + c] += o + sB2[128 + c];     // sum(alpha) = 1; residual (uni_denoiser.py:285)
#endif
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
            const int srow = it.chunk * 32 + lane;                  // row of the unit (edge k -> j); the lane with k = i is masked
            const bool rowvalid = sg.valid && srow < it.n - 1 && srow != sg.ti;
            const float* sQ = sQ0 + (tcount & 1) * (SM_Q / 8);
            TRACE(role, 0);
            // ---- key MLP
            tc::mbar_wait(&bars[B_PREK], ph);
            tc::tc_fence_after();
            TRACE(role, 1);
            // A lane quarter without a segment (the 3 spare slots of a unit's last tile: 9 % of all slots at n = 30) only keeps
            // the hand-off protocol going: it reads nothing, so it may arrive at once; its stale hid rows feed MMAs whose
            // output rows nobody reads.  The skipped arithmetic is power the capped GPU spends on clock instead.
            if (sg.valid) layer_norm(0, preK, hidK);
            else { tc::tc_fence_before(); tc::mbar_arrive(&bars[B_HIDK]); }
            TRACE(role, 2);
            // ---- value epilogue of the previous tile (its W2v MMA ran during that tile's logits and the LayerNorm above)
            if (tcount > 0) epilogue(ph ? 128 + tmem : 384 + tmem, ph ^ 1);
            TRACE(role, 4);
            // ---- value MLP
            tc::mbar_wait(&bars[ph ? B_PREV1 : B_PREV0], (tcount >> 1) & 1);
            tc::tc_fence_after();
            TRACE(role, 5);
            if (sg.valid) layer_norm(1, preV, hidV);
            else { tc::tc_fence_before(); tc::mbar_arrive(&bars[B_HIDV]); }
            TRACE(role, 6);
            // ---- logits of this thread's 4 heads, segment softmax across the 32 lanes (rows) of the warp
            {
                tc::mbar_wait(&bars[B_OUTK], ph);
                tc::tc_fence_after();
                TRACE(role, 7);
                if (sg.valid) {
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
#if defined(PG_TRIP_EXP) && (PG_TRIP_EXP & 2)
                } else if (true) {                            // knock-out experiment: no reductions (wrong results)
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] = rowvalid ? tc::ex2_approx(al[h] * 1e-3f) * 0.03f : 0.f;
#endif
                } else if (!(a.flags & 1)) {
                    // segment softmax across the 32 lanes: max and sum on the REDUX unit
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const float mx = tc::warp_max_redux(al[h]);
                        al[h] = rowvalid ? tc::ex2_approx(al[h] - mx) : 0.f;
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) al[h] *= tc::rcp_approx(tc::warp_sum01_redux(al[h]));
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
                    for (int h = 0; h < 4; h++) al[h] *= tc::rcp_approx(sm[h]);
                }
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
// Both k|v halves (256 channels), written as the bf16 hi/lo operand images trip_tc_kernel bulk-copies (layout at the top of
// this file): P in the block of the DESTINATION atom's unit at row k (its position among the edges into dst), R in the block
// of the SOURCE atom's unit at the segment index of dst, pre-swizzled for rows 11..14 of the angle slab.
// The smearing products depend on the distance only, which both directions of a bond edge share: a CTA takes one molecule
// and walks its UNDIRECTED pairs (s < t), computing smear(d_st) Wrkj and smear(d_st) Wrji once (half the FFMA work of an
// edge-wise kernel) and emitting the P and R rows of both directed edges s->t and t->s.
constexpr int PR_PAIRS = 2;      // pairs per warp and iteration: every weight row read from L1 serves 2 pairs = 4 edges
constexpr int PR_WARPS = 8;
__global__ void __launch_bounds__(PR_WARPS * 32, 2) trip_pr_kernel(TripTcArgs a) {
    const PlanDev& d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int n = d.g_n[g];
    if (n < 2) return;
    const int c0 = d.ctx_off[g] + d.g_p[g];                // context node of the molecule's first ligand atom
    const long long eo = d.eoff[g];
    const int npairs = n * (n - 1) / 2;
    const int half = lane >> 4, c8 = (lane & 15) >> 1, sub = (lane & 1) * 8;
    const size_t mlp_bytes = (size_t)d.Eb * 512;
    // this lane's 4 channels of one MLP -> bf16 hi and lo (8 bytes each) at chunk ((lane & 15) >> 1) ^ swizzle of the row
    auto put = [&](uint8_t* img, long long row0, int nr, int swz, float4 v) {
        uint32_t h0, l0, h1, l1;
        tc::split_pair_trunc(v.x, v.y, h0, l0);
        tc::split_pair_trunc(v.z, v.w, h1, l1);
        uint8_t* dst = img + (size_t)(row0 + (long long)half * nr) * 128 + ((c8 ^ swz) << 4) + sub;
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);                                    // hi image: blocks 0, 1
        *reinterpret_cast<uint2*>(dst + (size_t)2 * nr * 128) = make_uint2(l0, l1);             // lo image: blocks 2, 3
    };
    auto f4 = [](unsigned long long lo, unsigned long long hi) { const float2 x = up2(lo), y = up2(hi); return make_float4(x.x, x.y, y.x, y.y); };
    for (int p0 = (blockIdx.y * PR_WARPS + warp) * PR_PAIRS; p0 < npairs; p0 += gridDim.y * PR_WARPS * PR_PAIRS) {
        int sl[PR_PAIRS], tl[PR_PAIRS];
        float mine[PR_PAIRS];
        unsigned long long acc[PR_PAIRS][8];
        float4 pst[PR_PAIRS][2], pts[PR_PAIRS][2];         // partial P rows (k, v) of s->t and t->s before the smearing term
#pragma unroll
        for (int k = 0; k < PR_PAIRS; k++) {
            const int p = min(p0 + k, npairs - 1);
            // pair index p = t (t - 1) / 2 + s with s < t
            int t = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
            while (t * (t - 1) / 2 > p) t--;
            while ((t + 1) * t / 2 <= p) t++;
            tl[k] = t; sl[k] = p - t * (t - 1) / 2;
        }
        // every dependent load level is requested for both pairs before its first use
        float xs[PR_PAIRS][3], xt[PR_PAIRS][3];
#pragma unroll
        for (int k = 0; k < PR_PAIRS; k++) {
#pragma unroll
            for (int c = 0; c < 3; c++) { xs[k][c] = a.x[(size_t)(c0 + sl[k]) * 3 + c]; xt[k][c] = a.x[(size_t)(c0 + tl[k]) * 3 + c]; }
        }
#pragma unroll
        for (int k = 0; k < PR_PAIRS; k++) {
            const long long est = eo + (long long)tl[k] * (n - 1) + sl[k];            // s -> t: row s of unit t (s < t)
            const long long ets = eo + (long long)sl[k] * (n - 1) + (tl[k] - 1);      // t -> s: row t - 1 of unit s
            const float* hs = a.H + (size_t)(c0 + sl[k]) * a.ldh + lane * 4;
            const float* ht = a.H + (size_t)(c0 + tl[k]) * a.ldh + lane * 4;
            const float4 t0 = ldg4(a.T + (size_t)est * a.ldt + a.t_k + lane * 4), t1 = ldg4(a.T + (size_t)est * a.ldt + a.t_v + lane * 4);
            const float4 u0 = ldg4(a.T + (size_t)ets * a.ldt + a.t_k + lane * 4), u1 = ldg4(a.T + (size_t)ets * a.ldt + a.t_v + lane * 4);
            // (edge GEMM slice + h_src Whk) + h_dst Whj, as in the edge-wise formulation
            pst[k][0] = f4add(f4add(t0, ldg4(hs + a.hk_k)), ldg4(ht + a.hj_k));
            pst[k][1] = f4add(f4add(t1, ldg4(hs + a.hk_v)), ldg4(ht + a.hj_v));
            pts[k][0] = f4add(f4add(u0, ldg4(ht + a.hk_k)), ldg4(hs + a.hj_k));
            pts[k][1] = f4add(f4add(u1, ldg4(ht + a.hk_v)), ldg4(hs + a.hj_v));
            const float d0 = xt[k][0] - xs[k][0], d1 = xt[k][1] - xs[k][1], d2 = xt[k][2] - xs[k][2];
            const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            mine[k] = lane < 20 ? smear_val(dist, lane) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) acc[k][i] = 0ull;
        }
        // accumulators stay packed (two fp32 per 64-bit register pair) through the 20-term loop (FFMA2)
#pragma unroll 2
        for (int gg = 0; gg < 20; gg++) {
            const float4 w0 = ldg4(a.wrji + gg * 256 + lane * 4), w1 = ldg4(a.wrji + gg * 256 + 128 + lane * 4);
            const float4 w2 = ldg4(a.wrkj + gg * 256 + lane * 4), w3 = ldg4(a.wrkj + gg * 256 + 128 + lane * 4);
            const unsigned long long u0 = pk2(w0.x, w0.y), u1 = pk2(w0.z, w0.w), u2 = pk2(w1.x, w1.y), u3 = pk2(w1.z, w1.w);
            const unsigned long long u4 = pk2(w2.x, w2.y), u5 = pk2(w2.z, w2.w), u6 = pk2(w3.x, w3.y), u7 = pk2(w3.z, w3.w);
#pragma unroll
            for (int k = 0; k < PR_PAIRS; k++) {
                const float sg = __shfl_sync(PG_FULL, mine[k], gg);
                const unsigned long long ss = pk2(sg, sg);
                acc[k][0] = fma2_raw(ss, u0, acc[k][0]); acc[k][1] = fma2_raw(ss, u1, acc[k][1]);
                acc[k][2] = fma2_raw(ss, u2, acc[k][2]); acc[k][3] = fma2_raw(ss, u3, acc[k][3]);
                acc[k][4] = fma2_raw(ss, u4, acc[k][4]); acc[k][5] = fma2_raw(ss, u5, acc[k][5]);
                acc[k][6] = fma2_raw(ss, u6, acc[k][6]); acc[k][7] = fma2_raw(ss, u7, acc[k][7]);
            }
        }
#pragma unroll
        for (int k = 0; k < PR_PAIRS; k++) {
            if (p0 + k >= npairs) break;
            const int nr = n - 1;
            const float4 rk = f4(acc[k][0], acc[k][1]), rv = f4(acc[k][2], acc[k][3]);       // smear Wrji (key | value)
            const float4 sk = f4(acc[k][4], acc[k][5]), sv = f4(acc[k][6], acc[k][7]);       // smear Wrkj
            const long long ut = (eo + (long long)tl[k] * nr) * 4, us = (eo + (long long)sl[k] * nr) * 4;   // units of t / of s, in 128-byte rows
            // P rows: s -> t is row s of unit t, t -> s is row t - 1 of unit s
            put((uint8_t*)a.P, ut + sl[k], nr, sl[k] & 7, f4add(pst[k][0], sk));
            put((uint8_t*)a.P + mlp_bytes, ut + sl[k], nr, sl[k] & 7, f4add(pst[k][1], sv));
            put((uint8_t*)a.P, us + tl[k] - 1, nr, (tl[k] - 1) & 7, f4add(pts[k][0], sk));
            put((uint8_t*)a.P + mlp_bytes, us + tl[k] - 1, nr, (tl[k] - 1) & 7, f4add(pts[k][1], sv));
            // R rows live in the unit of the SOURCE atom at the segment index of the destination
            const int sidx_st = tl[k] - 1, sidx_ts = sl[k];                                   // s -> t in unit s; t -> s in unit t
            put((uint8_t*)a.R, us + sidx_st, nr, (R_ROW0 + (sidx_st & 3)) & 7, rk);
            put((uint8_t*)a.R + mlp_bytes, us + sidx_st, nr, (R_ROW0 + (sidx_st & 3)) & 7, rv);
            put((uint8_t*)a.R, ut + sidx_ts, nr, (R_ROW0 + (sidx_ts & 3)) & 7, rk);
            put((uint8_t*)a.R + mlp_bytes, ut + sidx_ts, nr, (R_ROW0 + (sidx_ts & 3)) & 7, rv);
        }
    }
}
}  // namespace

int pg_launch_trip_pr(const TripTcArgs& a, cudaStream_t s) {
    if (a.d.Eb <= 0) return PG_OK;
    // gridDim.x = molecules; gridDim.y splits a molecule's pairs so that the grid has at least ~8 waves of CTAs (tail effect)
    const int split = std::max(1, std::min(8, (8 * 2 * 148 + a.d.G - 1) / a.d.G));
    trip_pr_kernel<<<dim3((unsigned)a.d.G, (unsigned)split), PR_WARPS * 32, 0, s>>>(a);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

size_t pg_trip_tc_smem(int maxn) { return (size_t)SM_FIXED + (2 * (size_t)maxn * 4) * sizeof(float) + 1024; }

int pg_launch_trip_tc(const TripTcArgs& a, int num_sms, cudaStream_t s) {
    if (a.d.Nl <= 0) return PG_OK;
    const size_t smem = pg_trip_tc_smem(a.maxn);
    if (smem > 227 * 1024) { pg_set_error("trip_tc: shared memory budget exceeded"); return PG_ELIMIT; }
    static size_t cur = 0;
    if (smem > cur) {
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PG_CUDA_CHECK(cudaFuncSetAttribute(trip_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    const unsigned grid = (unsigned)std::min<long long>(a.d.Nl, num_sms);
    // The kernel a molecule runs on depends on the molecule alone (n - 2 <= 32: single-chunk instantiation, else the chunked
    // one), never on the batch it is in: results are bit-identical however a job is batched or sharded.  A mixed batch
    // launches both; each walks the unit list and skips the other's molecules.
    const bool have_single = a.d.min_n <= PG_TRIP_TC_SINGLE_CHUNK_ATOMS, have_multi = a.maxn > PG_TRIP_TC_SINGLE_CHUNK_ATOMS;
    if (a.flags & 2) {
        if (have_single) trip_tc_kernel<false, 0><<<grid, NTHREADS, smem, s>>>(a);
        if (have_multi) trip_tc_kernel<true, 0><<<grid, NTHREADS, smem, s>>>(a);
    } else if (a.flags & 4) {
        if (have_single) trip_tc_kernel<false, 2><<<grid, NTHREADS, smem, s>>>(a);
        if (have_multi) trip_tc_kernel<true, 2><<<grid, NTHREADS, smem, s>>>(a);
    } else {
        if (have_single) trip_tc_kernel<false, 1><<<grid, NTHREADS, smem, s>>>(a);
        if (have_multi) trip_tc_kernel<true, 1><<<grid, NTHREADS, smem, s>>>(a);
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

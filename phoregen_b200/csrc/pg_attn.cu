// Fused gather -> smear/edge-feature -> edge-MLP -> segmented softmax -> segmented reduce kernels
// (SURVEY.md §8(a) rows N1, P1, B2, E2, K3).  fp32 throughout.  One warp owns one destination segment
// (a dst node, or a dst bond edge for the triplet layer); a row of 128 channels is spread 4 per lane.
//
// Exact algebra used (changes fp summation order only; DESIGN.md "Factorisation"):
//  (1) the first Linear of every k/v MLP distributes over the concatenated input, so its per-node and
//      per-edge parts are computed once by the GEMM kernel and summed per row here;
//  (2) one-hot edge type (x) smearing = a 20x128 slice of the first Linear selected by the type;
//  (3) logits_h = q_h . (W2k_h hid + b2k_h) = hid . (W2k_h^T q_h) + q_h . b2k_h   (key second Linear folded
//      into the per-segment query);
//  (4) sum_e alpha_eh (W2v_h hid_e + b2v_h) = W2v_h (sum_e alpha_eh hid_e) + b2v_h sum_e alpha_eh
//      (value second Linear applied once per segment, after the segmented reduce).
// The softmax is a deterministic two-pass segment softmax (max, exp, sum) in registers/shared memory;
// no atomics anywhere.
#include "pg_attn.h"

namespace {
constexpr float kInvSqrtD = 0.35355339059327373f;   // 1/sqrt(8)  (uni_denoiser.py:62)

// fold the key MLP's second Linear into the query: qt[h][c] = sum_d W2k[h*8+d][c] q[h*8+d]
__device__ __forceinline__ void fold_query(const float* __restrict__ qrow, const float* __restrict__ w2k,
                                           const float* __restrict__ b2k, float* qs, int lane, float4 (&qt)[16],
                                           float& lb) {
#pragma unroll
    for (int i = 0; i < 4; i++) qs[lane + 32 * i] = qrow[lane + 32 * i];
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 16; h++) {
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int dd = 0; dd < 8; dd++) {
            const int o = h * 8 + dd;
            acc = f4fma(qs[o], ldg4(w2k + o * 128 + lane * 4), acc);
        }
        qt[h] = acc;
    }
    const int h = (lane >> 1) & 15;
    float s = 0.f;
#pragma unroll
    for (int dd = 0; dd < 8; dd++) s = fmaf(qs[h * 8 + dd], __ldg(b2k + h * 8 + dd), s);
    lb = s;
}

__device__ __forceinline__ void store_logits(const float4 hid, const float4 (&qt)[16], float lb, float* lg_row, int lane) {
    float p[16];
#pragma unroll
    for (int h = 0; h < 16; h++) p[h] = f4dot(hid, qt[h]);
    const float v = transpose_reduce16(p, lane);
    if (!(lane & 1)) lg_row[(lane >> 1) & 15] = (v + lb) * kInvSqrtD;
}

// Two-pass softmax over the R rows of lg[R][16] per head; result (optionally times ew[r]) written back.
// sw[h] = sum_r alpha[r][h] * ew[r]  (only when ew != nullptr).
__device__ __forceinline__ void segment_softmax(float* lg, int R, const float* ew, float* sw, int lane) {
    const int h = lane & 15, half = lane >> 4;
    float m = -INFINITY;
    for (int r = half; r < R; r += 2) m = fmaxf(m, lg[r * 16 + h]);
    m = fmaxf(m, __shfl_xor_sync(PG_FULL, m, 16));
    float s = 0.f;
    for (int r = half; r < R; r += 2) {
        const float e = expf(lg[r * 16 + h] - m);
        lg[r * 16 + h] = e;
        s += e;
    }
    s += __shfl_xor_sync(PG_FULL, s, 16);
    const float inv = 1.0f / s;
    float acc = 0.f;
    for (int r = half; r < R; r += 2) {
        float a = lg[r * 16 + h] * inv;
        if (ew) { a *= ew[r]; acc += a; }
        lg[r * 16 + h] = a;
    }
    if (ew) {
        acc += __shfl_xor_sync(PG_FULL, acc, 16);
        if (half == 0) sw[h] = acc;
    }
}

// out[o] = sum_c W2v[o][c] s[o/8][c] + b2v[o] * sw[o/8]; lane L gets outputs chunk*32 + L.
template <typename StoreFn>
__device__ __forceinline__ void value_out_transform(const float4 (&s)[16], const float* __restrict__ w2v,
                                                    const float* __restrict__ b2v, const float* sw, int lane,
                                                    StoreFn store) {
#pragma unroll
    for (int chunk = 0; chunk < 4; chunk++) {
        float p[32];
#pragma unroll
        for (int oo = 0; oo < 32; oo++) {
            const int o = chunk * 32 + oo;
            p[oo] = f4dot(ldg4(w2v + o * 128 + lane * 4), s[o >> 3]);
        }
        const float v = transpose_reduce32(p, lane);
        const int o = chunk * 32 + lane;
        const float swh = sw ? sw[o >> 3] : 1.0f;
        store(o, fmaf(__ldg(b2v + o), swh, v));
    }
}

// first-Linear contribution of the 24 edge features of two rows; when both rows have the same edge type the four
// 20x128 slices are read once for the pair (halves the L1 traffic of this table, the kernel's busiest data path)
__device__ __forceinline__ void edge_feature_pair(const float* __restrict__ tab, int t0, int t1, const float* f0, const float* f1,
                                                  int lane, float4& p0, float4& p1) {
    const float* tb0 = tab + (size_t)t0 * (24 * 128) + lane * 4;
    if (t0 == t1) {
#pragma unroll
        for (int f4 = 0; f4 < 6; f4++) {
            const float4 a = ld4(f0 + f4 * 4), b = ld4(f1 + f4 * 4);
            float4 w = ldg4(tb0 + (f4 * 4 + 0) * 128); p0 = f4fma(a.x, w, p0); p1 = f4fma(b.x, w, p1);
            w = ldg4(tb0 + (f4 * 4 + 1) * 128); p0 = f4fma(a.y, w, p0); p1 = f4fma(b.y, w, p1);
            w = ldg4(tb0 + (f4 * 4 + 2) * 128); p0 = f4fma(a.z, w, p0); p1 = f4fma(b.z, w, p1);
            w = ldg4(tb0 + (f4 * 4 + 3) * 128); p0 = f4fma(a.w, w, p0); p1 = f4fma(b.w, w, p1);
        }
    } else {
        const float* tb1 = tab + (size_t)t1 * (24 * 128) + lane * 4;
#pragma unroll
        for (int f4 = 0; f4 < 6; f4++) {
            const float4 a = ld4(f0 + f4 * 4), b = ld4(f1 + f4 * 4);
            p0 = f4fma(a.x, ldg4(tb0 + (f4 * 4 + 0) * 128), p0); p1 = f4fma(b.x, ldg4(tb1 + (f4 * 4 + 0) * 128), p1);
            p0 = f4fma(a.y, ldg4(tb0 + (f4 * 4 + 1) * 128), p0); p1 = f4fma(b.y, ldg4(tb1 + (f4 * 4 + 1) * 128), p1);
            p0 = f4fma(a.z, ldg4(tb0 + (f4 * 4 + 2) * 128), p0); p1 = f4fma(b.z, ldg4(tb1 + (f4 * 4 + 2) * 128), p1);
            p0 = f4fma(a.w, ldg4(tb0 + (f4 * 4 + 3) * 128), p0); p1 = f4fma(b.w, ldg4(tb1 + (f4 * 4 + 3) * 128), p1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ kNN / phore
// FEAT 0: joint ligand+pharmacophore kNN graph (NodeUpdateLayer / PosUpdateLayer with edge features,
//         uni_denoiser.py:264-281,291).  FEAT 1: pharmacophore encoder (p^2 edges incl. self loops, feature =
//         distance; diffusion.py:186-191).
template <int FEAT, int POS>
__global__ void __launch_bounds__(128) knn_attn_kernel(KnnAttnArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x * 4 + warp;
    const PlanDev& d = a.d;
    const int nrows = FEAT == 0 ? d.N : d.P;
    if (v >= nrows) return;
    constexpr int FW = FEAT == 0 ? 24 : 4;
    const int maxr = a.maxr;
    const int per_warp = 128 + maxr * 16 + maxr * FW + maxr * 4 + maxr + maxr + maxr + maxr + 16;
    float* qs = sm + (size_t)warp * per_warp;
    float* lg = qs + 128;
    float* feat = lg + maxr * 16;
    float* rel = feat + maxr * FW;
    float* ewr = rel + maxr * 4;
    int* srcs = (int*)(ewr + maxr);
    int* typ = srcs + maxr;
    int* ord = typ + maxr;           // rows regrouped by edge type so that two rows can share one pass over the feature table
    float* swv = (float*)(ord + maxr);

    int R, src0 = 0;
    long long e0 = 0;
    const int g = FEAT == 0 ? d.node_graph[v] : d.ph_graph[v];
    if (FEAT == 0) {
        const int ng = d.g_n[g] + d.g_p[g];
        R = min(PG_KNN, ng - 1);
        e0 = d.koff[g] + (long long)(v - d.ctx_off[g]) * R;
    } else {
        R = d.g_p[g];
        src0 = d.ph_off[g];
    }
    const float xv0 = a.x[(size_t)v * 3], xv1 = a.x[(size_t)v * 3 + 1], xv2 = a.x[(size_t)v * 3 + 2];
    // ---- per-row geometry and edge features (lane-parallel)
    for (int r = lane; r < R; r += 32) {
        const int s = FEAT == 0 ? a.knn_src[e0 + r] : src0 + r;
        srcs[r] = s;
        const float r0 = xv0 - a.x[(size_t)s * 3], r1 = xv1 - a.x[(size_t)s * 3 + 1], r2 = xv2 - a.x[(size_t)s * 3 + 2];
        rel[r * 4] = r0; rel[r * 4 + 1] = r1; rel[r * 4 + 2] = r2;            // rel_x = x[dst] - x[src]
        const float dist = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
        if (FEAT == 0) {
            const bool sl = (s - d.ctx_off[g]) >= d.g_p[g], dl = (v - d.ctx_off[g]) >= d.g_p[g];
            typ[r] = sl ? (dl ? 0 : 1) : (dl ? 2 : 3);                         // uni_denoiser.py:373-378
#pragma unroll
            for (int gg = 0; gg < 20; gg++) feat[r * FW + gg] = smear_val(dist, gg);
            feat[r * FW + 20] = 1.0f;
            const float* c1 = a.comb + (size_t)s * 3;                          // vec_1 = comb[src]
            const float* c2 = a.comb + (size_t)v * 3;                          // vec_2 = comb[dst]
            const float v30 = -r0, v31 = -r1, v32 = -r2;                       // vec_3 = x[src] - x[dst]
            feat[r * FW + 21] = c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2];
            feat[r * FW + 22] = c1[0] * v30 + c1[1] * v31 + c1[2] * v32;
            feat[r * FW + 23] = c2[0] * v30 + c2[1] * v31 + c2[2] * v32;
            ewr[r] = a.ew[e0 + r];
        } else {
            feat[r * FW] = dist;
        }
    }
    if (FEAT == 0) {      // R <= 32: lane = row.  Stable partition of the rows by edge type (4 ballots).
        const int myt = lane < R ? typ[lane] : 4;
        int pos = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const unsigned m = __ballot_sync(PG_FULL, myt == t);
            if (myt == t) pos += __popc(m & ((1u << lane) - 1));
            else if (myt > t) pos += __popc(m);
        }
        if (lane < R) ord[pos] = lane;
    }
    __syncwarp();

    // ---- pass 1: keys -> logits
    {
        float4 qt[16];
        float lb;
        fold_query(a.q + (size_t)v * 128, a.w.w2k, a.w.b2k, qs, lane, qt, lb);
        const float4 base = ldg4(a.nc.A + (size_t)v * a.nc.lda + a.nc.dst_k + lane * 4);
        const float4 gk = ldg4(a.w.lnk_g + lane * 4), bk = ldg4(a.w.lnk_b + lane * 4);
        float4 wd = make_float4(0, 0, 0, 0);
        if (FEAT == 1) wd = ldg4(a.w.tab_k + lane * 4);
        if (FEAT == 0) {
            for (int i = 0; i < R; i += 2) {
                const int r0 = ord[i], r1 = ord[min(i + 1, R - 1)];
                float4 p0 = f4add(base, ldg4(a.nc.A + (size_t)srcs[r0] * a.nc.lda + a.nc.src_k + lane * 4));
                float4 p1 = f4add(base, ldg4(a.nc.A + (size_t)srcs[r1] * a.nc.lda + a.nc.src_k + lane * 4));
                edge_feature_pair(a.w.tab_k, typ[r0], typ[r1], feat + r0 * FW, feat + r1 * FW, lane, p0, p1);
                store_logits(ln_relu_row(p0, gk, bk), qt, lb, lg + r0 * 16, lane);
                if (i + 1 < R) store_logits(ln_relu_row(p1, gk, bk), qt, lb, lg + r1 * 16, lane);
            }
        } else {
            for (int r = 0; r < R; r++) {
                float4 pre = f4add(base, ldg4(a.nc.A + (size_t)srcs[r] * a.nc.lda + a.nc.src_k + lane * 4));
                pre = f4fma(feat[r * FW], wd, pre);
                store_logits(ln_relu_row(pre, gk, bk), qt, lb, lg + r * 16, lane);
            }
        }
    }
    __syncwarp();
    segment_softmax(lg, R, FEAT == 0 ? ewr : nullptr, swv, lane);
    __syncwarp();

    // ---- pass 2: values
    const float4 base = ldg4(a.nc.A + (size_t)v * a.nc.lda + a.nc.dst_v + lane * 4);
    const float4 gv = ldg4(a.w.lnv_g + lane * 4), bv = ldg4(a.w.lnv_b + lane * 4);
    float4 wd = make_float4(0, 0, 0, 0);
    if (FEAT == 1) wd = ldg4(a.w.tab_v + lane * 4);
    // value-MLP hidden rows of a pair of rows (same pass over the feature table when both have the same edge type)
    auto value_hidden_pair = [&](int r0, int r1, float4& h0, float4& h1) {
        float4 p0 = f4add(base, ldg4(a.nc.A + (size_t)srcs[r0] * a.nc.lda + a.nc.src_v + lane * 4));
        float4 p1 = f4add(base, ldg4(a.nc.A + (size_t)srcs[r1] * a.nc.lda + a.nc.src_v + lane * 4));
        if (FEAT == 0) {
            edge_feature_pair(a.w.tab_v, typ[r0], typ[r1], feat + r0 * FW, feat + r1 * FW, lane, p0, p1);
        } else {
            p0 = f4fma(feat[r0 * FW], wd, p0);
            p1 = f4fma(feat[r1 * FW], wd, p1);
        }
        h0 = ln_relu_row(p0, gv, bv);
        h1 = ln_relu_row(p1, gv, bv);
    };
    auto row_at = [&](int i) { return FEAT == 0 ? ord[i] : i; };
    if (POS == 0) {
        float4 s[16];
#pragma unroll
        for (int h = 0; h < 16; h++) s[h] = make_float4(0, 0, 0, 0);
        for (int i = 0; i < R; i += 2) {
            const int r0 = row_at(i), r1 = row_at(min(i + 1, R - 1));
            float4 hid0, hid1;
            value_hidden_pair(r0, r1, hid0, hid1);
            const float w1 = (i + 1 < R) ? 1.f : 0.f;
#pragma unroll
            for (int h4 = 0; h4 < 4; h4++) {
                const float4 al = ld4(lg + r0 * 16 + h4 * 4);
                float4 bl = ld4(lg + r1 * 16 + h4 * 4);
                bl = make_float4(bl.x * w1, bl.y * w1, bl.z * w1, bl.w * w1);
                s[h4 * 4 + 0] = f4fma(bl.x, hid1, f4fma(al.x, hid0, s[h4 * 4 + 0]));
                s[h4 * 4 + 1] = f4fma(bl.y, hid1, f4fma(al.y, hid0, s[h4 * 4 + 1]));
                s[h4 * 4 + 2] = f4fma(bl.z, hid1, f4fma(al.z, hid0, s[h4 * 4 + 2]));
                s[h4 * 4 + 3] = f4fma(bl.w, hid1, f4fma(al.w, hid0, s[h4 * 4 + 3]));
            }
        }
        float* orow = a.out + (size_t)v * 128;
        value_out_transform(s, a.w.w2v, a.w.b2v, FEAT == 0 ? swv : nullptr, lane, [&](int o, float val) { orow[o] = val; });
    } else {
        float4 xv[16];
#pragma unroll
        for (int h = 0; h < 16; h++) xv[h] = ldg4(a.w.w2v + h * 128 + lane * 4);
        const int hh = (lane >> 1) & 15;
        const float b2 = __ldg(a.w.b2v + hh);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int i = 0; i < R; i += 2) {
            const int r0 = row_at(i), r1 = row_at(min(i + 1, R - 1));
            float4 hid0, hid1;
            value_hidden_pair(r0, r1, hid0, hid1);
            float p[16], pb[16];
#pragma unroll
            for (int h = 0; h < 16; h++) { p[h] = f4dot(hid0, xv[h]); pb[h] = f4dot(hid1, xv[h]); }
            float c0 = (transpose_reduce16(p, lane) + b2) * lg[r0 * 16 + hh];   // alpha (already times e_w) * v_h
            float c1 = (transpose_reduce16(pb, lane) + b2) * lg[r1 * 16 + hh];
            c0 = (lane & 1) ? 0.f : c0;
            c1 = ((lane & 1) || i + 1 >= R) ? 0.f : c1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { c0 += __shfl_xor_sync(PG_FULL, c0, o); c1 += __shfl_xor_sync(PG_FULL, c1, o); }
            a0 = fmaf(c1, rel[r1 * 4], fmaf(c0, rel[r0 * 4], a0));
            a1 = fmaf(c1, rel[r1 * 4 + 1], fmaf(c0, rel[r0 * 4 + 1], a1));
            a2 = fmaf(c1, rel[r1 * 4 + 2], fmaf(c0, rel[r0 * 4 + 2], a2));
        }
        if (lane == 0) {
            a.out[(size_t)v * 3] = a0 * (1.0f / 16.0f);
            a.out[(size_t)v * 3 + 1] = a1 * (1.0f / 16.0f);
            a.out[(size_t)v * 3 + 2] = a2 * (1.0f / 16.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ bond graph
// NodeUpdateLayer / PosUpdateLayer over the complete ligand bond graph (uni_denoiser.py:284,294): the
// segment of ligand atom i is the contiguous block of its n-1 incoming edges in the internal order.
template <int POS>
__global__ void __launch_bounds__(128, 4) bond_attn_kernel(BondAttnArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * 4 + warp;
    const PlanDev& d = a.d;
    if (u >= d.Nl) return;
    const int per_warp = 128 + a.maxr * 16;
    float* qs = sm + (size_t)warp * per_warp;
    float* lg = qs + 128;
    const int g = d.lig_graph[u];
    const int n = d.g_n[g], il = u - d.lig_off[g];
    const int ctx0 = d.ctx_off[g] + d.g_p[g];
    const int v = ctx0 + il;
    const int R = n - 1;
    if (a.tc_max_rows > 0 && R >= 1 && R <= a.tc_max_rows) return;
    const long long e0 = d.eoff[g] + (long long)il * (n - 1);
    {
        float4 qt[16];
        float lb;
        fold_query(a.q + (size_t)v * 128, a.w.w2k, a.w.b2k, qs, lane, qt, lb);
        const float4 base = ldg4(a.nc.A + (size_t)v * a.nc.lda + a.nc.dst_k + lane * 4);
        const float4 gk = ldg4(a.w.lnk_g + lane * 4), bk = ldg4(a.w.lnk_b + lane * 4);
        for (int r = 0; r < R; r++) {
            const int sj = ctx0 + r + (r >= il);
            float4 pre = f4add(base, ldg4(a.nc.A + (size_t)sj * a.nc.lda + a.nc.src_k + lane * 4));
            pre = f4add(pre, ldg4(a.B + (size_t)(e0 + r) * a.ldb + a.b_k + lane * 4));
            store_logits(ln_relu_row(pre, gk, bk), qt, lb, lg + r * 16, lane);
        }
    }
    __syncwarp();
    segment_softmax(lg, R, nullptr, nullptr, lane);
    __syncwarp();
    const float4 base = ldg4(a.nc.A + (size_t)v * a.nc.lda + a.nc.dst_v + lane * 4);
    const float4 gv = ldg4(a.w.lnv_g + lane * 4), bv = ldg4(a.w.lnv_b + lane * 4);
    auto value_hidden = [&](int r) {
        const int sj = ctx0 + r + (r >= il);
        float4 pre = f4add(base, ldg4(a.nc.A + (size_t)sj * a.nc.lda + a.nc.src_v + lane * 4));
        pre = f4add(pre, ldg4(a.B + (size_t)(e0 + r) * a.ldb + a.b_v + lane * 4));
        return ln_relu_row(pre, gv, bv);
    };
    if (POS == 0) {
        float4 s[16];
#pragma unroll
        for (int h = 0; h < 16; h++) s[h] = make_float4(0, 0, 0, 0);
        for (int r = 0; r < R; r++) {
            const float4 hid = value_hidden(r);
#pragma unroll
            for (int h4 = 0; h4 < 4; h4++) {
                const float4 al = ld4(lg + r * 16 + h4 * 4);
                s[h4 * 4 + 0] = f4fma(al.x, hid, s[h4 * 4 + 0]);
                s[h4 * 4 + 1] = f4fma(al.y, hid, s[h4 * 4 + 1]);
                s[h4 * 4 + 2] = f4fma(al.z, hid, s[h4 * 4 + 2]);
                s[h4 * 4 + 3] = f4fma(al.w, hid, s[h4 * 4 + 3]);
            }
        }
        float* orow = a.out + (size_t)v * 128;
        value_out_transform(s, a.w.w2v, a.w.b2v, nullptr, lane, [&](int o, float val) { orow[o] = val; });
    } else {
        float4 xv[16];
#pragma unroll
        for (int h = 0; h < 16; h++) xv[h] = ldg4(a.w.w2v + h * 128 + lane * 4);
        const int hh = (lane >> 1) & 15;
        const float b2 = __ldg(a.w.b2v + hh);
        const float xi0 = a.x[(size_t)v * 3], xi1 = a.x[(size_t)v * 3 + 1], xi2 = a.x[(size_t)v * 3 + 2];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int r = 0; r < R; r++) {
            const float4 hid = value_hidden(r);
            float p[16];
#pragma unroll
            for (int h = 0; h < 16; h++) p[h] = f4dot(hid, xv[h]);
            float c = (transpose_reduce16(p, lane) + b2) * lg[r * 16 + hh];
            c = warp_sum((lane & 1) ? 0.f : c);
            const int sj = ctx0 + r + (r >= il);
            a0 = fmaf(c, xi0 - a.x[(size_t)sj * 3], a0);                 // rel_x = x[dst] - x[src]
            a1 = fmaf(c, xi1 - a.x[(size_t)sj * 3 + 1], a1);
            a2 = fmaf(c, xi2 - a.x[(size_t)sj * 3 + 2], a2);
        }
        if (lane == 0) {
            a.out[(size_t)v * 3] = a0 * (1.0f / 16.0f);
            a.out[(size_t)v * 3 + 1] = a1 * (1.0f / 16.0f);
            a.out[(size_t)v * 3 + 2] = a2 * (1.0f / 16.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ bond triplets
// BondUpdateLayer (uni_denoiser.py:123-165).  One CTA per ligand atom j: the per-edge partial
// P[k->j] = h_bond[kj] Wb + r_kj Wrkj + h_k Whk + h_j Whj + b1 of its n-1 incoming edges is staged once in
// shared memory and shared by the n-1 segments (j->i); one warp per segment.  Triplet indices are arithmetic
// (complete graph): rows k ascending over all atoms except i and j.
constexpr int TRIP_WARPS = 8;
__global__ void __launch_bounds__(TRIP_WARPS * 32) trip_kernel(TripArgs a) {
    extern __shared__ __align__(16) float sm[];
    const PlanDev& d = a.d;
    const int u = blockIdx.x;                      // ligand atom j (global ligand index)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = d.lig_graph[u];
    const int n = d.g_n[g], jl = u - d.lig_off[g];
    if (n < 3 || n < a.min_atoms) return;
    const int ctx0 = d.ctx_off[g] + d.g_p[g];
    const int cj = ctx0 + jl;
    const long long eoff = d.eoff[g];
    float* Ps = sm;                                        // [(maxn-1)][256]
    float* smr = Ps + (size_t)(a.maxn - 1) * 256;          // [(maxn-1)][20]
    const int per_warp = 128 + a.maxr * 16 + a.maxr * 16;
    float* wbase = smr + (size_t)(a.maxn - 1) * 20 + (size_t)warp * per_warp;
    float* qs = wbase;
    float* lg = qs + 128;
    float* feat = lg + a.maxr * 16;
    const float xj0 = a.x[(size_t)cj * 3], xj1 = a.x[(size_t)cj * 3 + 1], xj2 = a.x[(size_t)cj * 3 + 2];

    // ---- stage P rows of the edges k -> j
    for (int t = tid; t < n - 1; t += blockDim.x) {
        const int ck = ctx0 + t + (t >= jl);
        const float d0 = xj0 - a.x[(size_t)ck * 3], d1 = xj1 - a.x[(size_t)ck * 3 + 1], d2 = xj2 - a.x[(size_t)ck * 3 + 2];
        const float dist = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
#pragma unroll
        for (int gg = 0; gg < 20; gg++) smr[t * 20 + gg] = smear_val(dist, gg);
    }
    __syncthreads();
    {
        const int c = tid;                                  // 256 channels: k | v
        const int cc = c & 127;
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
            for (int gg = 0; gg < 20; gg++) val = fmaf(smr[t * 20 + gg], wr[gg], val);
            Ps[t * 256 + c] = val;
        }
    }
    __syncthreads();

    const int R = n - 2;
    const float4 gk = ldg4(a.w.lnk_g + lane * 4), bk = ldg4(a.w.lnk_b + lane * 4);
    const float4 gv = ldg4(a.w.lnv_g + lane * 4), bv = ldg4(a.w.lnv_b + lane * 4);
    for (int il = warp; il < n; il += TRIP_WARPS) {
        if (il == jl) continue;
        const int ci = ctx0 + il;
        const long long eji = eoff + (long long)il * (n - 1) + (jl - (jl > il));
        const float xi0 = a.x[(size_t)ci * 3], xi1 = a.x[(size_t)ci * 3 + 1], xi2 = a.x[(size_t)ci * 3 + 2];
        const float pj0 = xj0 - xi0, pj1 = xj1 - xi1, pj2 = xj2 - xi2;            // pos_ji = pos[j] - pos[i]
        const int ti = il - (il > jl);                                            // P row of k == i (skipped)
        // r_ji part of the first Linear (same for all rows of the segment)
        float4 basek = make_float4(0, 0, 0, 0), basev = basek;
        {
            const float dist = sqrtf(pj0 * pj0 + pj1 * pj1 + pj2 * pj2);
            const float mine = lane < 20 ? smear_val(dist, lane) : 0.f;
#pragma unroll
            for (int gg = 0; gg < 20; gg++) {
                const float sg = __shfl_sync(PG_FULL, mine, gg);
                basek = f4fma(sg, ldg4(a.wrji + gg * 256 + lane * 4), basek);
                basev = f4fma(sg, ldg4(a.wrji + gg * 256 + 128 + lane * 4), basev);
            }
        }
        // angular encoding of every triplet of the segment (common.py:67-87; uni_denoiser.py:131-135)
        for (int r = lane; r < R; r += 32) {
            const int t = r + (r >= ti);
            const int ck = ctx0 + t + (t >= jl);
            const float pk0 = a.x[(size_t)ck * 3] - xi0, pk1 = a.x[(size_t)ck * 3 + 1] - xi1, pk2 = a.x[(size_t)ck * 3 + 2] - xi2;
            const float dotv = pj0 * pk0 + pj1 * pk1 + pj2 * pk2;
            const float c0 = pj1 * pk2 - pj2 * pk1, c1 = pj2 * pk0 - pj0 * pk2, c2 = pj0 * pk1 - pj1 * pk0;
            const float th = atan2f(sqrtf(c0 * c0 + c1 * c1 + c2 * c2), dotv);
            float* f = feat + r * 16;
            float s1, k1, s2, k2, s3, k3, sh, kh, st, kt;
            sincosf(th, &s1, &k1); sincosf(th * 2.0f, &s2, &k2); sincosf(th * 3.0f, &s3, &k3);
            sincosf(th * 0.5f, &sh, &kh); sincosf(th * (1.0f / 3.0f), &st, &kt);
            f[0] = th;
            f[1] = s1; f[2] = s2; f[3] = s3; f[4] = s1; f[5] = sh; f[6] = st;
            f[7] = k1; f[8] = k2; f[9] = k3; f[10] = k1; f[11] = kh; f[12] = kt;
        }
        {   // pass 1
            float4 qt[16];
            float lb;
            fold_query(a.q + (size_t)eji * 128, a.w.w2k, a.w.b2k, qs, lane, qt, lb);   // includes __syncwarp
            for (int r = 0; r < R; r++) {
                const int t = r + (r >= ti);
                float4 pre = f4add(ld4(Ps + t * 256 + lane * 4), basek);
                const float* f = feat + r * 16;
#pragma unroll
                for (int ff = 0; ff < 13; ff++) pre = f4fma(f[ff], ldg4(a.wa + ff * 256 + lane * 4), pre);
                store_logits(ln_relu_row(pre, gk, bk), qt, lb, lg + r * 16, lane);
            }
        }
        __syncwarp();
        segment_softmax(lg, R, nullptr, nullptr, lane);
        __syncwarp();
        float4 s[16];
#pragma unroll
        for (int h = 0; h < 16; h++) s[h] = make_float4(0, 0, 0, 0);
        for (int r = 0; r < R; r++) {
            const int t = r + (r >= ti);
            float4 pre = f4add(ld4(Ps + t * 256 + 128 + lane * 4), basev);
            const float* f = feat + r * 16;
#pragma unroll
            for (int ff = 0; ff < 13; ff++) pre = f4fma(f[ff], ldg4(a.wa + ff * 256 + 128 + lane * 4), pre);
            const float4 hid = ln_relu_row(pre, gv, bv);
#pragma unroll
            for (int h4 = 0; h4 < 4; h4++) {
                const float4 al = ld4(lg + r * 16 + h4 * 4);
                s[h4 * 4 + 0] = f4fma(al.x, hid, s[h4 * 4 + 0]);
                s[h4 * 4 + 1] = f4fma(al.y, hid, s[h4 * 4 + 1]);
                s[h4 * 4 + 2] = f4fma(al.z, hid, s[h4 * 4 + 2]);
                s[h4 * 4 + 3] = f4fma(al.w, hid, s[h4 * 4 + 3]);
            }
        }
        float* orow = a.hb + (size_t)eji * 128;
        value_out_transform(s, a.w.w2v, a.w.b2v, nullptr, lane, [&](int o, float val) { orow[o] += val; });   // residual (uni_denoiser.py:285)
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ global edge weight
// e_w = sigmoid(MLP(20->128->LN->ReLU->1)(smear(|x_dst - x_src|)))  (uni_denoiser.py:410-415).
// Warp per destination node: its <= 32 kNN edges are contiguous, so the neighbour indices and distances are fetched by
// the lanes in one round trip, and the first-Linear rows stay in registers across the edges.
__global__ void __launch_bounds__(256) edge_weight_kernel(PlanDev d, const float* __restrict__ x,
                                                          const int* __restrict__ knn_src, const float* __restrict__ w1t,
                                                          const float* __restrict__ b1, const float* __restrict__ g,
                                                          const float* __restrict__ b, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, float* __restrict__ ew) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (v >= d.N) return;
    const int gph = d.node_graph[v];
    const int ng = d.g_n[gph] + d.g_p[gph];
    const int R = min(PG_KNN, ng - 1);
    if (R <= 0) return;
    const long long e0 = d.koff[gph] + (long long)(v - d.ctx_off[gph]) * R;
    float dist = 0.f;
    if (lane < R) {
        const int s = knn_src[e0 + lane];
        const float r0 = x[(size_t)v * 3] - x[(size_t)s * 3], r1 = x[(size_t)v * 3 + 1] - x[(size_t)s * 3 + 1], r2 = x[(size_t)v * 3 + 2] - x[(size_t)s * 3 + 2];
        dist = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
    }
    float4 w[20];
#pragma unroll
    for (int gg = 0; gg < 20; gg++) w[gg] = ldg4(w1t + gg * 128 + lane * 4);
    const float4 bias = ldg4(b1 + lane * 4), g4 = ldg4(g + lane * 4), b4 = ldg4(b + lane * 4), w24 = ldg4(w2 + lane * 4);
    const float bb2 = __ldg(b2);
    float mine_out = 0.f;
    for (int r = 0; r < R; r++) {
        const float dr = __shfl_sync(PG_FULL, dist, r);
        const float mine = lane < 20 ? smear_val(dr, lane) : 0.f;
        float4 pre = bias;
#pragma unroll
        for (int gg = 0; gg < 20; gg++) pre = f4fma(__shfl_sync(PG_FULL, mine, gg), w[gg], pre);
        const float4 hid = ln_relu_row(pre, g4, b4);
        const float logit = warp_sum(f4dot(hid, w24)) + bb2;
        if (lane == r) mine_out = 1.0f / (1.0f + expf(-logit));
    }
    if (lane < R) ew[e0 + lane] = mine_out;
}
}  // namespace

int pg_launch_knn_attn(const KnnAttnArgs& a, int feat, int pos, cudaStream_t s) {
    const int rows = feat == 0 ? a.d.N : a.d.P;
    if (rows <= 0) return PG_OK;
    const int FW = feat == 0 ? 24 : 4;
    const size_t smem = (size_t)4 * (128 + a.maxr * 16 + a.maxr * FW + a.maxr * 4 + a.maxr * 4 + 16) * sizeof(float);
    const unsigned grid = (unsigned)((rows + 3) / 4);
    if (smem > 200 * 1024) { pg_set_error("knn_attn: segment too long (%d rows)", a.maxr); return PG_ELIMIT; }
#define PG_KA(F, P)                                                                                                   \
    do {                                                                                                              \
        static size_t cur = 0;                                                                                        \
        if (smem > cur) {                                                                                             \
            PG_CUDA_CHECK(cudaFuncSetAttribute(knn_attn_kernel<F, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cur = smem;                                                                                               \
        }                                                                                                             \
        knn_attn_kernel<F, P><<<grid, 128, smem, s>>>(a);                                                             \
    } while (0)
    if (feat == 0 && pos == 0) PG_KA(0, 0);
    else if (feat == 0 && pos == 1) PG_KA(0, 1);
    else if (feat == 1 && pos == 0) PG_KA(1, 0);
    else { pg_set_error("knn_attn: unsupported variant"); return PG_EINVAL; }
#undef PG_KA
    PG_LAUNCH_CHECK();
    return PG_OK;
}

int pg_launch_bond_attn(const BondAttnArgs& a, int pos, cudaStream_t s) {
    if (a.d.Nl <= 0) return PG_OK;
    const size_t smem = (size_t)4 * (128 + a.maxr * 16) * sizeof(float);
    const unsigned grid = (unsigned)((a.d.Nl + 3) / 4);
    if (pos == 0) {
        static size_t cur = 0;
        if (smem > cur) { PG_CUDA_CHECK(cudaFuncSetAttribute(bond_attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
        bond_attn_kernel<0><<<grid, 128, smem, s>>>(a);
    } else {
        static size_t cur = 0;
        if (smem > cur) { PG_CUDA_CHECK(cudaFuncSetAttribute(bond_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
        bond_attn_kernel<1><<<grid, 128, smem, s>>>(a);
    }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

int pg_launch_trip(const TripArgs& a, cudaStream_t s) {
    if (a.d.Nl <= 0 || a.maxn < 3) return PG_OK;
    const size_t smem = ((size_t)(a.maxn - 1) * 276 + (size_t)TRIP_WARPS * (128 + a.maxr * 32)) * sizeof(float);
    if (smem > 220 * 1024) { pg_set_error("trip: molecule too large for shared memory (%d atoms)", a.maxn); return PG_ELIMIT; }
    static size_t cur = 0;
    if (smem > cur) { PG_CUDA_CHECK(cudaFuncSetAttribute(trip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
    trip_kernel<<<(unsigned)a.d.Nl, TRIP_WARPS * 32, smem, s>>>(a);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

int pg_launch_edge_weight(const PlanDev& d, const float* x, const int* knn_src, const float* w1t, const float* b1,
                          const float* g, const float* b, const float* w2, const float* b2, float* ew, cudaStream_t s) {
    if (d.Ek <= 0) return PG_OK;
    edge_weight_kernel<<<(unsigned)((d.N + 7) / 8), 256, 0, s>>>(d, x, knn_src, w1t, b1, g, b, w2, b2, ew);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

// Packed-weight slot table, embedding / head kernels and the forward orchestration
// (call stack §3.2).
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include "pg_attn.h"
#include "pg_gemm.h"
#include "pg_plan.h"

int pg_launch_knn(PgPlan* p, const float* x, const float* phore_norm, int mode, int* knn_src, float* comb,
                  int64_t* ei, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ slots
namespace {
const char* kSub[5] = {"nk", "nb", "tr", "pk", "pb"};   // node-kNN, node-bond, triplet, pos-kNN, pos-bond

std::vector<PgSlotDesc> build_slots() {
    std::vector<PgSlotDesc> v;
    auto add = [&](const std::string& n, long long numel) {
        PgSlotDesc s;
        memset(&s, 0, sizeof(s));
        strncpy(s.name, n.c_str(), sizeof(s.name) - 1);
        s.numel = numel;
        v.push_back(s);
    };
    add("G.node_emb_t", 12 * 118); add("G.edge_emb_t", 6 * 118);
    add("G.time_coeff", 10); add("G.time_offset", 10);
    add("G.ph_emb_wt", 18 * 128); add("G.ph_emb_b", 128);
    add("PE.wcat_t", 128 * 640); add("PE.wcat_t.bf", 128 * 640); add("PE.bcat", 640);
    add("PE.wd_k", 128); add("PE.wd_v", 128);
    add("PE.lnk_g", 128); add("PE.lnk_b", 128); add("PE.lnv_g", 128); add("PE.lnv_b", 128);
    add("PE.lnq_g", 128); add("PE.lnq_b", 128); add("PE.w2q_t", 128 * 128); add("PE.w2q_t.bf", 128 * 128); add("PE.b2q", 128);
    add("PE.w2k", 128 * 128); add("PE.b2k", 128); add("PE.w2v", 128 * 128); add("PE.b2v", 128);
    add("G.ew.w1t", 20 * 128); add("G.ew.b1", 128); add("G.ew.ln_g", 128); add("G.ew.ln_b", 128);
    add("G.ew.w2", 128); add("G.ew.b2", 4);
    add("G.vinf.w1t", 128 * 128); add("G.vinf.w1t.bf", 128 * 128); add("G.vinf.b1", 128); add("G.vinf.w2", 12 * 128); add("G.vinf.b2", 12);
    for (int c = 0; c < 2; c++) {     // atom-count heads atom_mlp / atom_mlp_1 (diffusion.py:77-88)
        const std::string C = "G.cnt" + std::to_string(c) + ".";
        add(C + "w1t", 128 * 256); add(C + "b1", 256); add(C + "w2", 256); add(C + "b2", 4);
    }
    add("G.binf.w1t", 128 * 128); add("G.binf.w1t.bf", 128 * 128); add("G.binf.b1", 128); add("G.binf.w2", 6 * 128); add("G.binf.b2", 8);
    for (int l = 0; l < PG_NUM_LAYERS; l++) {
        std::string L = "L" + std::to_string(l) + ".";
        add(L + "n1.wt", 128 * 1920); add(L + "n1.wt.bf", 128 * 1920); add(L + "n1.b", 1920);
        add(L + "e1.wt", 128 * 640); add(L + "e1.wt.bf", 128 * 640); add(L + "e1.b", 640);
        add(L + "n2.wt", 128 * 1280); add(L + "n2.wt.bf", 128 * 1280); add(L + "n2.b", 1280);
        add(L + "e2.wt", 128 * 256); add(L + "e2.wt.bf", 128 * 256); add(L + "e2.b", 256);
        add(L + "lin.wt", 128 * 128); add(L + "lin.wt.bf", 128 * 128); add(L + "lin.b", 128);
        // e2 of this layer and e1 of the next read the same h_bond: one GEMM with 896 output columns (tcgen05 path)
        if (l + 1 < PG_NUM_LAYERS) { add(L + "e2e1.wt", 128 * 896); add(L + "e2e1.wt.bf", 128 * 896); }
        for (int s = 0; s < 5; s++) {
            std::string S = L + kSub[s] + ".";
            add(S + "lnq_g", 128); add(S + "lnq_b", 128); add(S + "w2q_t", 128 * 128); add(S + "w2q_t.bf", 128 * 128); add(S + "b2q", 128);
            add(S + "lnk_g", 128); add(S + "lnk_b", 128); add(S + "lnv_g", 128); add(S + "lnv_b", 128);
            add(S + "lnk_bf", 128); add(S + "lnv_bf", 128); add(S + "fold", 4);   // tensor-core kernels: beta (/ gamma when folded into W2), flags
            add(S + "w2k", 128 * 128); add(S + "b2k", 128);
            const bool pos = s >= 3;
            add(S + "w2v", (pos ? 16 : 128) * 128); add(S + "b2v", pos ? 16 : 128);
            if (s == 0 || s == 3) { add(S + "tab_k", 4 * 24 * 128); add(S + "tab_v", 4 * 24 * 128); }
            if (s != 2) { add(S + "w2k.bf", 128 * 128); add(S + "w2v.bf", (pos ? 16 : 128) * 128); }   // bond_tc / knn_tc
            add(S + "w2k.h", 128 * 64);                                                                 // key second Linear as fp16
            if (s == 0 || s == 3) { add(S + "tab_k.bf", 96 * 128); add(S + "tab_v.bf", 96 * 128); }
            if (s == 2) {
                add(S + "wrkj", 20 * 256); add(S + "wrji", 20 * 256); add(S + "wa", 13 * 256);
                add(S + "w2k.bf", 128 * 128); add(S + "w2v.bf", 128 * 128); add(S + "wa.bf", 2 * 256 * 16 / 2);
            }
        }
    }
    return v;
}
}  // namespace

const std::vector<PgSlotDesc>& pg_slot_table() {
    static std::vector<PgSlotDesc> t = build_slots();
    return t;
}
int pg_slot_index(const char* name) {
    const auto& t = pg_slot_table();
    for (size_t i = 0; i < t.size(); i++)
        if (!strcmp(t[i].name, name)) return (int)i;
    return -1;
}
extern "C" int pg_weight_slot_count(void) { return (int)pg_slot_table().size(); }
extern "C" const char* pg_weight_slot_name(int s) {
    const auto& t = pg_slot_table();
    return (s < 0 || s >= (int)t.size()) ? nullptr : t[s].name;
}
extern "C" int64_t pg_weight_slot_numel(int s) {
    const auto& t = pg_slot_table();
    return (s < 0 || s >= (int)t.size()) ? -1 : t[s].numel;
}
extern "C" int pg_model_create(PgModel** out, const float* d_blob, const int64_t* h_offsets, int n_slots) {
    if (!out || !d_blob || !h_offsets || n_slots != pg_weight_slot_count()) { pg_set_error("pg_model_create: expected %d slots", pg_weight_slot_count()); return PG_EINVAL; }
    for (int i = 0; i < n_slots; i++)
        if (h_offsets[i] < 0 || (h_offsets[i] & 3)) { pg_set_error("slot %d offset must be a non-negative multiple of 4 floats", i); return PG_EINVAL; }
    if (((uintptr_t)d_blob & 15) != 0) { pg_set_error("weight blob must be 16-byte aligned"); return PG_EINVAL; }
    PgModel* m = new PgModel();
    m->blob = d_blob;
    m->off.assign(h_offsets, h_offsets + n_slots);
    *out = m;
    return PG_OK;
}
extern "C" void pg_model_destroy(PgModel* m) { delete m; }

namespace {
struct W {   // named slot lookup, cached per model instance lifetime is unnecessary: lookups are cheap string compares
    const PgModel* m;
    const float* operator()(const std::string& name) const {
        int i = pg_slot_index(name.c_str());
        return i < 0 ? nullptr : m->w(i);
    }
};

// ------------------------------------------------------------------------------------------------ small kernels
// time embedding (common.py:34-55, type_='linear'): exp(coeff_g * (clamp(t) - offset_g)^2)
__device__ __forceinline__ float time_feat(float t, const float* coeff, const float* offset, int gq) {
    t = fminf(fmaxf(t, 0.f), offset[9]);
    const float dlt = t - offset[gq];
    return expf(coeff[gq] * dlt * dlt);
}

// context rows (diffusion.py:180-181,193-200): ligand row = [node_embedder(h_node) (118) | time (10)], phore row = encoder output
__global__ void __launch_bounds__(256) embed_nodes_kernel(PlanDev d, const float* __restrict__ h_node, const float* __restrict__ pos,
                                                          const int64_t* __restrict__ tstep, const float* __restrict__ h_ph,
                                                          const float* __restrict__ pos_ph, const float* __restrict__ wn_t,
                                                          const float* __restrict__ tcoef, const float* __restrict__ toff,
                                                          float* __restrict__ h, float* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (v >= d.N) return;
    const int g = d.node_graph[v];
    const int loc = v - d.ctx_off[g], p = d.g_p[g];
    if (loc < p) {
        const int pi = d.ph_off[g] + loc;
        st4(h + (size_t)v * 128 + lane * 4, ldg4(h_ph + (size_t)pi * 128 + lane * 4));
        if (lane < 3) x[(size_t)v * 3 + lane] = pos_ph[(size_t)pi * 3 + lane];
    } else {
        const int ai = d.lig_off[g] + loc - p;
        const float t = (float)tstep[g];
        float hv[12];
#pragma unroll
        for (int c = 0; c < 12; c++) hv[c] = h_node[(size_t)ai * 12 + c];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int o = lane * 4 + i;
            float acc;
            if (o < 118) {
                acc = 0.f;
#pragma unroll
                for (int c = 0; c < 12; c++) acc = fmaf(hv[c], __ldg(wn_t + c * 118 + o), acc);
            } else {
                acc = time_feat(t, tcoef, toff, o - 118);
            }
            h[(size_t)v * 128 + o] = acc;
        }
        if (lane < 3) x[(size_t)v * 3 + lane] = pos[(size_t)ai * 3 + lane];
    }
}

// bond rows (diffusion.py:183,205): [edge_embedder(h_edge) (118) | time (10)], scattered to the internal edge order
__global__ void __launch_bounds__(256) embed_edges_kernel(PlanDev d, const float* __restrict__ h_edge, const int64_t* __restrict__ tstep,
                                                          const float* __restrict__ we_t, const float* __restrict__ tcoef,
                                                          const float* __restrict__ toff, float* __restrict__ hb) {
    const int lane = threadIdx.x & 31;
    // 4 reference-order edges per warp; every dependent load level is requested for all four before its first use
    const long long e0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;
    if (e0 >= d.Eb) return;
    long long ee[4], slot[4];
    int gr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { ee[k] = min(e0 + k, d.Eb - 1); gr[k] = d.edge_graph[ee[k]]; slot[k] = d.perm[ee[k]]; }
    float hv[4][6], t[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int c = 0; c < 6; c++) hv[k][c] = h_edge[(size_t)ee[k] * 6 + c];
        t[k] = (float)tstep[gr[k]];
    }
    float wv[4][6];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int o = lane * 4 + i;
#pragma unroll
        for (int c = 0; c < 6; c++) wv[i][c] = o < 118 ? __ldg(we_t + c * 118 + o) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (e0 + k >= d.Eb) break;
        float4 o4;
        float* op = &o4.x;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int o = lane * 4 + i;
            float acc;
            if (o < 118) {
                acc = 0.f;
#pragma unroll
                for (int c = 0; c < 6; c++) acc = fmaf(hv[k][c], wv[i][c], acc);
            } else {
                acc = time_feat(t[k], tcoef, toff, o - 118);
            }
            op[i] = acc;
        }
        st4(hb + (size_t)slot[k] * 128 + lane * 4, o4);
    }
}

// phore_embedding Linear(18 -> 128) (diffusion.py:186)
__global__ void __launch_bounds__(256) phore_embed_kernel(int P, const float* __restrict__ xph, const float* __restrict__ wt,
                                                          const float* __restrict__ b, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (v >= P) return;
    float4 acc = ldg4(b + lane * 4);
#pragma unroll
    for (int c = 0; c < PG_PHORE_FEAT; c++) acc = f4fma(xph[(size_t)v * PG_PHORE_FEAT + c], ldg4(wt + c * 128 + lane * 4), acc);
    st4(out + (size_t)v * 128 + lane * 4, acc);
}

// rows of a [*,128] matrix permuted between reference and internal edge order
__global__ void __launch_bounds__(256) permute_rows_kernel(long long rows, const int* __restrict__ perm, const float* __restrict__ src,
                                                           float* __restrict__ dst, int scatter) {
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= rows) return;
    const long long s = perm[e];
    if (scatter) st4(dst + (size_t)s * 128 + lane * 4, ld4(src + (size_t)e * 128 + lane * 4));   // dst[perm[e]] = src[e]
    else st4(dst + (size_t)e * 128 + lane * 4, ld4(src + (size_t)s * 128 + lane * 4));           // dst[e] = src[perm[e]]
}

// x += (dx1 + dx2) * mask_ligand  (uni_denoiser.py:295-296)
__global__ void pos_update_kernel(PlanDev d, float* __restrict__ x, const float* __restrict__ dx1, const float* __restrict__ dx2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N * 3) return;
    const int v = i / 3;
    const int g = d.node_graph[v];
    if (v - d.ctx_off[g] >= d.g_p[g]) x[i] += dx1[i] + dx2[i];
}

// second half of the output heads (diffusion.py:55-59,71-75): out = W2 . (softplus(hid) - ln 2) + b2 ; warp per row.
// ROWSEL 0: ligand atoms (row a -> context node), also emits the atom's final position; 1: reference-order edges.
template <int K, int ROWSEL>
__global__ void __launch_bounds__(256) head_out_kernel(PlanDev d, const float* __restrict__ hid, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ out,
                                                       const float* __restrict__ x, float* __restrict__ x_out) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long rows = ROWSEL == 0 ? d.Nl : d.Eb;
    if (r >= rows) return;
    long long srow;
    if (ROWSEL == 0) {
        const int g = d.lig_graph[r];
        srow = d.ctx_off[g] + d.g_p[g] + (r - d.lig_off[g]);
        if (x_out && lane < 3) x_out[r * 3 + lane] = x[srow * 3 + lane];
    } else {
        srow = d.perm[r];
    }
    float4 v = ld4(hid + (size_t)srow * 128 + lane * 4);
    // shifted softplus (common.py:26-33); fast exp/log: absolute error < 1e-6, far inside the 1e-4 absolute parity bar
    auto ssp = [](float t) { return (t > 20.f ? t : __logf(1.0f + __expf(t))) - 0.69314718055994531f; };
    v = make_float4(ssp(v.x), ssp(v.y), ssp(v.z), ssp(v.w));
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float s = warp_sum(f4dot(v, ldg4(w2 + k * 128 + lane * 4)));
        if (lane == 0) out[r * K + k] = s + __ldg(b2 + k);
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------------ orchestration
namespace {
#define PG_TRY(expr)              \
    do {                          \
        int _r = (expr);          \
        if (_r != PG_OK) return _r; \
    } while (0)

inline int round4(int v) { return (v + 3) & ~3; }

int num_sms() {
    static int v = 0;
    if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); if (v <= 0) v = 148; }
    return v;
}
// tcgen05 triplet kernel unless PG_TRIP=fp32
bool use_tc_trip(const PlanDev& d) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PG_TRIP"); v = (e && !strcmp(e, "fp32")) ? 0 : 1; }
    return v == 1 && d.max_n >= 3;
}

// Precision of the key MLPs' second Linear (its output only feeds softmax logits).  Default bf16x3 everywhere.  Opt-in
// single-pass fp16 (measured, configs[1]): PG_KEY=trip16 (triplet kernel only) 71.8 -> 67.1 ms/step, model outputs at
// 0.50 x tolerance but the internal h_bond at 1.07 x; PG_KEY=fp16 (all three attention kernels) 65.4 ms/step, outputs at
// 0.72 x, internal h / h_bond at 1.36 x / 1.20 x.  PG_KEY=trip16x2 (triplet kernel: fp16 hi/lo ACTIVATIONS against the single
// fp16 weight image, 16 MMAs instead of 24): 65.2 -> 61.8 ms/step in same-box A/B, outputs at <= 0.42 x, internal h_bond
// at 1.11 x -- the fp16 rounding of the weights alone costs as much as that of the activations.  None is the default
// because of the internal-state bar.
int key_mode() {      // 0 bf16x3 everywhere, 2 fp16 everywhere, 3 fp16 in the triplet kernel only, 4 fp16 hi/lo activations x fp16 weights in the triplet kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("PG_KEY"); v = !e ? 0 : !strcmp(e, "fp16") ? 2 : !strcmp(e, "trip16") ? 3 : !strcmp(e, "trip16x2") ? 4 : 0; }
    return v;
}

// tcgen05 bond-graph attention unless PG_BOND=fp32 (the FFMA reference kernel, A/B validation only)
bool use_tc_bond() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PG_BOND"); v = (e && !strcmp(e, "fp32")) ? 0 : 1; }
    return v == 1;
}
template <class W>
int launch_bond(PgPlan* p, const BondAttnArgs& a0, const W& w, const std::string& S, int pos, cudaStream_t s) {
    BondAttnArgs a = a0;
    a.tc_max_rows = 0;
    PgTimed timed(p, KC_BOND_ATTN, s);
    if (use_tc_bond()) {
        BondTcArgs t;
        t.d = a.d; t.x = a.x; t.nc = a.nc; t.B = a.B; t.ldb = a.ldb; t.b_k = a.b_k; t.b_v = a.b_v; t.q = a.q; t.w = a.w;
        t.w2k_bf = (const uint16_t*)w(S + "w2k.bf"); t.w2v_bf = (const uint16_t*)w(S + "w2v.bf"); t.out = a.out;
        t.w2k_h = (const uint16_t*)w(S + "w2k.h"); t.key_bf16x3 = key_mode() != 2;
        PG_TRY(pg_launch_bond_tc(t, pos, num_sms(), s));
        p->launches++;
        return PG_OK;       // every atom has 1 <= n-1 <= 127 incoming edges (pg_plan_create): all segments ran there
    }
    PG_TRY(pg_launch_bond_attn(a, pos, s));
    p->launches++;
    return PG_OK;
}

// tcgen05 kNN-graph attention unless PG_KNN_ATTN=fp32
bool use_tc_knn() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PG_KNN_ATTN"); v = (e && !strcmp(e, "fp32")) ? 0 : 1; }
    return v == 1;
}
template <class W>
int launch_knn_attn(PgPlan* p, const KnnAttnArgs& a, const W& w, const std::string& S, int pos, cudaStream_t s) {
    PgTimed timed(p, KC_KNN_ATTN, s);
    if (use_tc_knn()) {
        KnnTcArgs t;
        t.d = a.d; t.x = a.x; t.comb = a.comb; t.knn_src = a.knn_src; t.ew = a.ew; t.nc = a.nc; t.q = a.q; t.w = a.w;
        t.w2k_bf = (const uint16_t*)w(S + "w2k.bf"); t.w2v_bf = (const uint16_t*)w(S + "w2v.bf");
        t.tabk_bf = (const uint16_t*)w(S + "tab_k.bf"); t.tabv_bf = (const uint16_t*)w(S + "tab_v.bf");
        t.w2k_h = (const uint16_t*)w(S + "w2k.h"); t.key_bf16x3 = key_mode() != 2;
        t.alpha = p->abuf; t.alpha_sum = p->abuf + (size_t)a.d.Ek * 16; t.out = a.out;
        PG_TRY(pg_launch_knn_tc(t, pos, num_sms(), s));
        p->launches += 2;
        return PG_OK;
    }
    PG_TRY(pg_launch_knn_attn(a, 0, pos, s));
    p->launches++;
    return PG_OK;
}

bool use_simt_gemm() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PG_GEMM"); v = (e && !strcmp(e, "simt")) ? 1 : 0; }
    return v == 1;
}

// `wname` is the slot of the k-major fp32 weight; "<wname>.bf" holds its bf16 hi/lo split for the tcgen05 kernel
int gemm(PgPlan* p, cudaStream_t s, int pro, long long M, const float* A, long long lda, const W& w, const std::string& wname,
         long long ldw, const float* bias, float* C, long long ldc, int ntiles, const float* A2 = nullptr, long long lda2 = 0,
         const int* gidx = nullptr, const float* lng = nullptr, const float* lnb = nullptr, const float* resid = nullptr,
         long long ldr = 0, float* C2 = nullptr, long long ldc2 = 0, int csplit = 0) {
    GemmArgs a;
    a.C2 = C2; a.ldc2 = ldc2; a.csplit = csplit;
    a.M = M; a.A = A; a.lda = lda; a.A2 = A2; a.lda2 = lda2; a.gidx = gidx; a.ln_g = lng; a.ln_b = lnb;
    a.Wt = w(wname); a.Wbf = w(wname + ".bf"); a.ldw = ldw; a.bias = bias; a.C = C; a.ldc = ldc; a.ntiles = ntiles;
    a.resid = resid; a.ldr = ldr; a.relu = 0;
    p->launches++;
    PgTimed timed(p, KC_GEMM, s);
    static const char* only = getenv("PG_GEMM_SIMT_ONLY");      // debug: fp32 kernel for weights whose slot name contains this substring
    const bool simt = use_simt_gemm() || (only && strstr(wname.c_str(), only));
    return simt ? pg_launch_gemm(a, pro, s) : pg_launch_gemm_tc(a, pro, s);
}

AttnW attn_w(const W& w, const std::string& S, bool tabs) {
    AttnW a;
    a.tab_k = tabs ? w(S + "tab_k") : nullptr; a.tab_v = tabs ? w(S + "tab_v") : nullptr;
    a.lnk_g = w(S + "lnk_g"); a.lnk_b = w(S + "lnk_b"); a.lnv_g = w(S + "lnv_g"); a.lnv_b = w(S + "lnv_b");
    a.lnk_bf = w(S + "lnk_bf"); a.lnv_bf = w(S + "lnv_bf"); a.fold = w(S + "fold");
    a.w2k = w(S + "w2k"); a.b2k = w(S + "b2k"); a.w2v = w(S + "w2v"); a.b2v = w(S + "b2v");
    return a;
}

// column blocks of the packed node GEMMs (must match phoregen_b200/weights.py)
enum { N1_NK_DK = 0, N1_NK_SK = 128, N1_NK_DV = 256, N1_NK_SV = 384, N1_NK_Q = 512, N1_NB_DK = 640, N1_NB_SK = 768,
       N1_NB_DV = 896, N1_NB_SV = 1024, N1_NB_Q = 1152, N1_TR_HK_K = 1280, N1_TR_HJ_K = 1408, N1_TR_HK_V = 1536,
       N1_TR_HJ_V = 1664, N1_TR_Q = 1792, N1_COLS = 1920 };
enum { E1_NB_K = 0, E1_NB_V = 128, E1_TR_K = 256, E1_TR_V = 384, E1_TR_Q = 512, E1_COLS = 640 };
enum { N2_PK_DK = 0, N2_PK_SK = 128, N2_PK_DV = 256, N2_PK_SV = 384, N2_PK_Q = 512, N2_PB_DK = 640, N2_PB_SK = 768,
       N2_PB_DV = 896, N2_PB_SV = 1024, N2_PB_Q = 1152, N2_COLS = 1280 };

// the 6-layer denoiser on the plan's internal buffers h / x / hb (uni_denoiser.py:396-430)
int run_denoiser(const PgModel* m, PgPlan* p, const float* phore_norm, cudaStream_t s) {
    const PlanDev& d = p->d;
    const W w{m};
    const long long N = d.N, Eb = d.Eb;
    PG_CUDA_CHECK(cudaMemsetAsync(p->o2, 0, (size_t)N * 128 * sizeof(float), s));
    PG_CUDA_CHECK(cudaMemsetAsync(p->dx2, 0, (size_t)N * 3 * sizeof(float), s));
    PG_TRY(pg_launch_knn(p, p->x, nullptr, 0, p->knn_src, nullptr, nullptr, s));
    PG_TRY(pg_launch_edge_weight(d, p->x, p->knn_src, w("G.ew.w1t"), w("G.ew.b1"), w("G.ew.ln_g"), w("G.ew.ln_b"),
                                 w("G.ew.w2"), w("G.ew.b2"), p->ew, s));
    p->launches++;
    const int maxr_knn = round4(std::min(PG_KNN, d.max_ng - 1));
    const int maxr_bond = round4(d.max_n - 1);
    const int maxr_trip = round4(std::max(d.max_n - 2, 1));
    for (int l = 0; l < PG_NUM_LAYERS; l++) {
        const std::string L = "L" + std::to_string(l) + ".";
        // direction vectors (k=3 ligand kNN) of this layer's coordinates
        PG_TRY(pg_launch_knn(p, p->x, phore_norm, 1, nullptr, p->comb, nullptr, s));
        // first Linear of every MLP, node and bond parts
        PG_TRY(gemm(p, s, PRO_PLAIN, N, p->h, 128, w, L + "n1.wt", N1_COLS, w(L + "n1.b"), p->nbuf, N1_COLS, N1_COLS / 128));
        // (from the second layer on the previous layer's merged e2 | e1 GEMM has already produced this layer's edge partials)
        static const bool no_merge = getenv("PG_NO_MERGE_E") != nullptr;      // A/B switch: separate e2 / e1 GEMMs as with PG_GEMM=simt
        const bool merged_e = !use_simt_gemm() && !no_merge;
        if (l == 0 || !merged_e)
            PG_TRY(gemm(p, s, PRO_PLAIN, Eb, p->hb, 128, w, L + "e1.wt", E1_COLS, w(L + "e1.b"), p->ebuf, E1_COLS, E1_COLS / 128));
        // queries: LN -> ReLU -> second Linear
        PG_TRY(gemm(p, s, PRO_LNRELU_MF, N, p->nbuf + N1_NK_Q, N1_COLS, w, L + "nk.w2q_t", 128, w(L + "nk.b2q"), p->qn1, 128, 1,
                    nullptr, 0, nullptr, w(L + "nk.lnq_g"), w(L + "nk.lnq_b")));
        PG_TRY(gemm(p, s, PRO_LNRELU_MF, N, p->nbuf + N1_NB_Q, N1_COLS, w, L + "nb.w2q_t", 128, w(L + "nb.b2q"), p->qn2, 128, 1,
                    nullptr, 0, nullptr, w(L + "nb.lnq_g"), w(L + "nb.lnq_b")));
        PG_TRY(gemm(p, s, PRO_LNRELU_MF, Eb, p->ebuf + E1_TR_Q, E1_COLS, w, L + "tr.w2q_t", 128, w(L + "tr.b2q"), p->qt, 128, 1,
                    p->nbuf + N1_TR_Q, N1_COLS, d.edst_node, w(L + "tr.lnq_g"), w(L + "tr.lnq_b")));
        {   // node update over the kNN graph
            KnnAttnArgs a;
            a.d = d; a.x = p->x; a.comb = p->comb; a.knn_src = p->knn_src; a.ew = p->ew;
            a.nc = NodeCols{p->nbuf, N1_COLS, N1_NK_DK, N1_NK_SK, N1_NK_DV, N1_NK_SV};
            a.q = p->qn1; a.w = attn_w(w, L + "nk.", true); a.out = p->o1; a.maxr = maxr_knn;
            PG_TRY(launch_knn_attn(p, a, w, L + "nk.", 0, s));
        }
        {   // node update over the bond graph
            BondAttnArgs a;
            a.d = d; a.x = p->x;
            a.nc = NodeCols{p->nbuf, N1_COLS, N1_NB_DK, N1_NB_SK, N1_NB_DV, N1_NB_SV};
            a.B = p->ebuf; a.ldb = E1_COLS; a.b_k = E1_NB_K; a.b_v = E1_NB_V;
            a.q = p->qn2; a.w = attn_w(w, L + "nb.", false); a.out = p->o2; a.maxr = maxr_bond;
            PG_TRY(launch_bond(p, a, w, L + "nb.", 0, s));
        }
        {   // bond update over triplets (uses the old h, x); h_bond updated in place
            TripArgs a;
            a.d = d; a.x = p->x; a.T = p->ebuf; a.ldt = E1_COLS; a.t_k = E1_TR_K; a.t_v = E1_TR_V;
            a.H = p->nbuf; a.ldh = N1_COLS; a.hk_k = N1_TR_HK_K; a.hj_k = N1_TR_HJ_K; a.hk_v = N1_TR_HK_V; a.hj_v = N1_TR_HJ_V;
            a.q = p->qt; a.wrkj = w(L + "tr.wrkj"); a.wrji = w(L + "tr.wrji"); a.wa = w(L + "tr.wa");
            a.w = attn_w(w, L + "tr.", false); a.hb = p->hb; a.maxr = maxr_trip; a.maxn = d.max_n;
            a.min_atoms = 0;
            if (use_tc_trip(d)) {
                TripTcArgs t;
                t.d = d; t.x = p->x; t.T = a.T; t.ldt = a.ldt; t.t_k = a.t_k; t.t_v = a.t_v; t.H = a.H; t.ldh = a.ldh;
                t.hk_k = a.hk_k; t.hj_k = a.hj_k; t.hk_v = a.hk_v; t.hj_v = a.hj_v; t.q = p->qt; t.R = p->rbuf; t.P = p->pbuf2;
                t.wrkj = a.wrkj; t.wrji = a.wrji;
                t.w2k_bf = (const uint16_t*)w(L + "tr.w2k.bf"); t.w2v_bf = (const uint16_t*)w(L + "tr.w2v.bf");
                t.w2k_h = (const uint16_t*)w(L + "tr.w2k.h");
                t.wa_bf = (const uint16_t*)w(L + "tr.wa.bf");
                { static int fl = -1; if (fl < 0) { const char* e = getenv("PG_TRIP_FLAGS"); fl = e ? atoi(e) : 0; } t.flags = fl | (key_mode() >= 2 ? 0 : 2) | (key_mode() == 4 ? 4 : 0); }       // bit 1 set = bf16x3 key MLP
                t.lnk_g = a.w.lnk_g; t.lnk_b = a.w.lnk_b; t.lnv_g = a.w.lnv_g; t.lnv_b = a.w.lnv_b; t.b2k = a.w.b2k; t.b2v = a.w.b2v;
                t.lnk_bf = a.w.lnk_bf; t.lnv_bf = a.w.lnv_bf; t.fold = a.w.fold;
                t.hb = p->hb; t.maxn = d.max_n;
                { PgTimed timed(p, KC_OTHER, s); PG_TRY(pg_launch_trip_pr(t, s)); }
                { PgTimed timed(p, KC_TRIP, s); PG_TRY(pg_launch_trip_tc(t, num_sms(), s)); } p->launches += 2;
            } else {   // PG_TRIP=fp32: the FFMA reference kernel (A/B validation only)
                PgTimed timed(p, KC_TRIP, s); PG_TRY(pg_launch_trip(a, s)); p->launches++;
            }
        }
        // h <- h + lin_node(o1 + o2)
        PG_TRY(gemm(p, s, PRO_SUM2, N, p->o1, 128, w, L + "lin.wt", 128, w(L + "lin.b"), p->h, 128, 1, p->o2, 128, nullptr,
                    nullptr, nullptr, p->h, 128));
        // position update with the new h / h_bond and the old coordinates
        PG_TRY(gemm(p, s, PRO_PLAIN, N, p->h, 128, w, L + "n2.wt", N2_COLS, w(L + "n2.b"), p->nbuf, N2_COLS, N2_COLS / 128));
        // bond partials of the position layer.  h_bond does not change between here and the next layer's first Linears, so
        // with the tcgen05 GEMM one launch computes [e2 of this layer | e1 of the next] (896 columns, both biases are zero):
        // the e2 part goes to the triplet work space (the P images are dead until the next trip_pr), the e1 part to ebuf.
        float* e2buf = p->ebuf;
        if (merged_e && l + 1 < PG_NUM_LAYERS) {
            e2buf = p->pbuf2;
            PG_TRY(gemm(p, s, PRO_PLAIN, Eb, p->hb, 128, w, L + "e2e1.wt", 896, nullptr, e2buf, 256, 7, nullptr, 0, nullptr, nullptr, nullptr,
                        nullptr, 0, p->ebuf, E1_COLS, 256));
        } else {
            PG_TRY(gemm(p, s, PRO_PLAIN, Eb, p->hb, 128, w, L + "e2.wt", 256, w(L + "e2.b"), p->ebuf, 256, 2));
        }
        PG_TRY(gemm(p, s, PRO_LNRELU_MF, N, p->nbuf + N2_PK_Q, N2_COLS, w, L + "pk.w2q_t", 128, w(L + "pk.b2q"), p->qn1, 128, 1,
                    nullptr, 0, nullptr, w(L + "pk.lnq_g"), w(L + "pk.lnq_b")));
        PG_TRY(gemm(p, s, PRO_LNRELU_MF, N, p->nbuf + N2_PB_Q, N2_COLS, w, L + "pb.w2q_t", 128, w(L + "pb.b2q"), p->qn2, 128, 1,
                    nullptr, 0, nullptr, w(L + "pb.lnq_g"), w(L + "pb.lnq_b")));
        {
            KnnAttnArgs a;
            a.d = d; a.x = p->x; a.comb = p->comb; a.knn_src = p->knn_src; a.ew = p->ew;
            a.nc = NodeCols{p->nbuf, N2_COLS, N2_PK_DK, N2_PK_SK, N2_PK_DV, N2_PK_SV};
            a.q = p->qn1; a.w = attn_w(w, L + "pk.", true); a.out = p->dx1; a.maxr = maxr_knn;
            PG_TRY(launch_knn_attn(p, a, w, L + "pk.", 1, s));
        }
        {
            BondAttnArgs a;
            a.d = d; a.x = p->x;
            a.nc = NodeCols{p->nbuf, N2_COLS, N2_PB_DK, N2_PB_SK, N2_PB_DV, N2_PB_SV};
            a.B = e2buf; a.ldb = 256; a.b_k = 0; a.b_v = 128;
            a.q = p->qn2; a.w = attn_w(w, L + "pb.", false); a.out = p->dx2; a.maxr = maxr_bond;
            PG_TRY(launch_bond(p, a, w, L + "pb.", 1, s));
        }
        pos_update_kernel<<<(unsigned)((N * 3 + 255) / 256), 256, 0, s>>>(d, p->x, p->dx1, p->dx2);
        PG_LAUNCH_CHECK(); p->launches++;
    }
    return PG_OK;
}
}  // namespace

extern "C" int pg_phore_encode(const PgModel* m, PgPlan* p, const float* d_h_phore, const float* d_pos_phore, float* d_out,
                               void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const PlanDev& d = p->d;
    const W w{m};
    if (d.max_p > 256) { pg_set_error("phore encoder: more than 256 pharmacophore nodes per graph"); return PG_ELIMIT; }
    phore_embed_kernel<<<(unsigned)((d.P + 7) / 8), 256, 0, s>>>(d.P, d_h_phore, w("G.ph_emb_wt"), w("G.ph_emb_b"), p->pemb);
    PG_LAUNCH_CHECK(); p->launches++;
    PG_TRY(gemm(p, s, PRO_PLAIN, d.P, p->pemb, 128, w, "PE.wcat_t", 640, w("PE.bcat"), p->pbuf, 640, 5));
    PG_TRY(gemm(p, s, PRO_LNRELU, d.P, p->pbuf + 512, 640, w, "PE.w2q_t", 128, w("PE.b2q"), p->pq, 128, 1, nullptr, 0, nullptr,
                w("PE.lnq_g"), w("PE.lnq_b")));
    KnnAttnArgs a;
    a.d = d; a.x = d_pos_phore; a.comb = nullptr; a.knn_src = nullptr; a.ew = nullptr;
    a.nc = NodeCols{p->pbuf, 640, 0, 128, 256, 384};
    a.q = p->pq;
    a.w.tab_k = w("PE.wd_k"); a.w.tab_v = w("PE.wd_v");
    a.w.lnk_g = w("PE.lnk_g"); a.w.lnk_b = w("PE.lnk_b"); a.w.lnv_g = w("PE.lnv_g"); a.w.lnv_b = w("PE.lnv_b");
    a.w.w2k = w("PE.w2k"); a.w.b2k = w("PE.b2k"); a.w.w2v = w("PE.w2v"); a.w.b2v = w("PE.b2v");
    a.out = d_out; a.maxr = round4(d.max_p);
    PG_TRY(pg_launch_knn_attn(a, 1, 0, s)); p->launches++;
    return PG_OK;
}

extern "C" int pg_denoiser_forward(const PgModel* m, PgPlan* p, const float* d_h, const float* d_x, const float* d_h_bond,
                                   const float* d_phore_norm, float* d_h_out, float* d_x_out, float* d_h_bond_out,
                                   void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const PlanDev& d = p->d;
    PG_CUDA_CHECK(cudaMemcpyAsync(p->h, d_h, (size_t)d.N * 128 * 4, cudaMemcpyDeviceToDevice, s));
    PG_CUDA_CHECK(cudaMemcpyAsync(p->x, d_x, (size_t)d.N * 3 * 4, cudaMemcpyDeviceToDevice, s));
    permute_rows_kernel<<<(unsigned)((d.Eb + 7) / 8), 256, 0, s>>>(d.Eb, d.perm, d_h_bond, p->hb, 1);
    PG_LAUNCH_CHECK(); p->launches++;
    PG_TRY(run_denoiser(m, p, d_phore_norm, s));
    PG_CUDA_CHECK(cudaMemcpyAsync(d_h_out, p->h, (size_t)d.N * 128 * 4, cudaMemcpyDeviceToDevice, s));
    PG_CUDA_CHECK(cudaMemcpyAsync(d_x_out, p->x, (size_t)d.N * 3 * 4, cudaMemcpyDeviceToDevice, s));
    permute_rows_kernel<<<(unsigned)((d.Eb + 7) / 8), 256, 0, s>>>(d.Eb, d.perm, p->hb, d_h_bond_out, 0);
    PG_LAUNCH_CHECK(); p->launches++;
    return PG_OK;
}

extern "C" int pg_phorediff_forward(const PgModel* m, PgPlan* p, const float* d_h_node, const float* d_pos,
                                    const float* d_h_edge, const int64_t* d_time_step, const float* d_h_phore_emb,
                                    const float* d_pos_phore, const float* d_phore_norm, float* d_logits_node,
                                    float* d_pos_out, float* d_logits_edge, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const PlanDev& d = p->d;
    const W w{m};
    embed_nodes_kernel<<<(unsigned)((d.N + 7) / 8), 256, 0, s>>>(d, d_h_node, d_pos, d_time_step, d_h_phore_emb, d_pos_phore,
                                                                  w("G.node_emb_t"), w("G.time_coeff"), w("G.time_offset"), p->h, p->x);
    PG_LAUNCH_CHECK(); p->launches++;
    embed_edges_kernel<<<(unsigned)((d.Eb + 31) / 32), 256, 0, s>>>(d, d_h_edge, d_time_step, w("G.edge_emb_t"), w("G.time_coeff"),
                                                                   w("G.time_offset"), p->hb);
    PG_LAUNCH_CHECK(); p->launches++;
    PG_TRY(run_denoiser(m, p, d_phore_norm, s));
    // output heads
    PG_TRY(gemm(p, s, PRO_PLAIN, d.N, p->h, 128, w, "G.vinf.w1t", 128, w("G.vinf.b1"), p->qn1, 128, 1));
    head_out_kernel<PG_NODE_CLASSES, 0><<<(unsigned)((d.Nl + 7) / 8), 256, 0, s>>>(d, p->qn1, w("G.vinf.w2"), w("G.vinf.b2"),
                                                                                   d_logits_node, p->x, d_pos_out);
    PG_LAUNCH_CHECK(); p->launches++;
    PG_TRY(gemm(p, s, PRO_PLAIN, d.Eb, p->hb, 128, w, "G.binf.w1t", 128, w("G.binf.b1"), p->qt, 128, 1));
    head_out_kernel<PG_EDGE_CLASSES, 1><<<(unsigned)((d.Eb + 7) / 8), 256, 0, s>>>(d, p->qt, w("G.binf.w2"), w("G.binf.b2"),
                                                                                   d_logits_edge, nullptr, nullptr);
    PG_LAUNCH_CHECK(); p->launches++;
    return PG_OK;
}

// ------------------------------------------------------------------------------------------------ O2 / D2: atom-count heads
// predict_atom_count (diffusion.py:148-163): count = mean_g sigmoid(atom_mlp(h_p)); count_l = mean over the non-EX nodes of
// sigmoid(atom_mlp_1(h_p)); count_u = count_l + relu(count - count_l).  One CTA per graph, thread = hidden unit of the
// 128 -> 256 -> 1 heads, nodes of the graph in order (deterministic sums).  Also emits the integer interval of
// sample_nodes (diffusion.py:379-380): round-half-even(c * (max_atom - min_atom) + min_atom).
namespace {
__global__ void __launch_bounds__(256) atom_count_kernel(PlanDev d, const float* __restrict__ hp, const float* __restrict__ xph,
                                                         int ex_col, const float* __restrict__ w1t0, const float* __restrict__ b10,
                                                         const float* __restrict__ w20, const float* __restrict__ b20,
                                                         const float* __restrict__ w1t1, const float* __restrict__ b11,
                                                         const float* __restrict__ w21, const float* __restrict__ b21,
                                                         float min_atom, float max_atom, float* __restrict__ cl, float* __restrict__ cu,
                                                         int* __restrict__ lo, int* __restrict__ hi) {
    const int g = blockIdx.x, j = threadIdx.x;
    __shared__ float row[128];
    __shared__ float red[2][8];
    const int p0 = d.ph_off[g], p = d.g_p[g];
    float s0 = 0.f, s1 = 0.f;
    int n1 = 0;
    for (int v = 0; v < p; v++) {
        if (j < 128) row[j] = hp[(size_t)(p0 + v) * 128 + j];
        __syncthreads();
        const bool nonex = xph[(size_t)(p0 + v) * PG_PHORE_FEAT + ex_col] != 1.0f;
        float a0 = b10[j], a1 = b11[j];
#pragma unroll 8
        for (int k = 0; k < 128; k++) { a0 = fmaf(row[k], w1t0[k * 256 + j], a0); a1 = fmaf(row[k], w1t1[k * 256 + j], a1); }
        float t0 = fmaxf(a0, 0.f) * w20[j], t1 = fmaxf(a1, 0.f) * w21[j];
        t0 = warp_sum(t0); t1 = warp_sum(t1);
        if ((j & 31) == 0) { red[0][j >> 5] = t0; red[1][j >> 5] = t1; }
        __syncthreads();
        if (j == 0) {
            float u0 = b20[0], u1 = b21[0];
            for (int w = 0; w < 8; w++) { u0 += red[0][w]; u1 += red[1][w]; }
            s0 += 1.0f / (1.0f + expf(-u0));
            if (nonex) { s1 += 1.0f / (1.0f + expf(-u1)); n1++; }
        }
        __syncthreads();
    }
    if (j == 0) {
        const float c = s0 / (float)max(p, 1);
        const float l = s1 / (float)max(n1, 1);           // scatter-mean of an empty set is 0 (torch_scatter clamps the count)
        const float u = l + fmaxf(c - l, 0.f);
        cl[g] = l; cu[g] = u;
        if (lo) { lo[g] = (int)rintf(l * (max_atom - min_atom) + min_atom); hi[g] = (int)rintf(u * (max_atom - min_atom) + min_atom); }
    }
}
}  // namespace

extern "C" int pg_atom_count(const PgModel* m, PgPlan* p, const float* d_h_phore_emb, const float* d_h_phore, int ex_col,
                             float min_atom, float max_atom, float* d_count_l, float* d_count_u, int32_t* d_lo, int32_t* d_hi,
                             void* stream) {
    if (!d_h_phore_emb || !d_h_phore || !d_count_l || !d_count_u || ex_col < 0 || ex_col >= PG_PHORE_FEAT || (!d_lo) != (!d_hi)) {
        pg_set_error("pg_atom_count: bad argument"); return PG_EINVAL;
    }
    const W w{m};
    atom_count_kernel<<<(unsigned)p->d.G, 256, 0, (cudaStream_t)stream>>>(
        p->d, d_h_phore_emb, d_h_phore, ex_col, w("G.cnt0.w1t"), w("G.cnt0.b1"), w("G.cnt0.w2"), w("G.cnt0.b2"), w("G.cnt1.w1t"),
        w("G.cnt1.b1"), w("G.cnt1.w2"), w("G.cnt1.b2"), min_atom, max_atom, d_count_l, d_count_u, d_lo, d_hi);
    PG_LAUNCH_CHECK();
    p->launches++;
    return PG_OK;
}

// ------------------------------------------------------------------------------------------------ stand-alone contraction
// C[M, 128*ntiles128] = pro(A)[M,128] @ W + bias (+ resid), the building block of every first / second Linear here.
// impl 0: tcgen05 bf16x3 kernel (needs d_w_bf16_tiles = weights.bf16_tiles64 image), impl 1: fp32 FFMA kernel (needs d_wt).
extern "C" int pg_gemm_k128(int impl, int prologue, int64_t M, const float* d_a, int64_t lda, const float* d_a2, int64_t lda2,
                            const int32_t* d_gather, const float* d_ln_g, const float* d_ln_b, const float* d_wt,
                            const float* d_w_bf16_tiles, const float* d_bias, const float* d_resid, int64_t ldr, float* d_c,
                            int64_t ldc, int ntiles128, void* stream) {
    if (prologue < 0 || prologue > 2 || ntiles128 <= 0) { pg_set_error("pg_gemm_k128: bad argument"); return PG_EINVAL; }
    GemmArgs a;
    a.M = M; a.A = d_a; a.lda = lda; a.A2 = d_a2; a.lda2 = lda2; a.gidx = d_gather; a.ln_g = d_ln_g; a.ln_b = d_ln_b;
    a.Wt = d_wt; a.Wbf = d_w_bf16_tiles; a.ldw = 128LL * ntiles128; a.bias = d_bias; a.C = d_c; a.ldc = ldc; a.ntiles = ntiles128;
    a.resid = d_resid; a.ldr = ldr; a.relu = 0; a.C2 = nullptr; a.ldc2 = 0; a.csplit = 0;
    return impl == 1 ? pg_launch_gemm(a, prologue, (cudaStream_t)stream) : pg_launch_gemm_tc(a, prologue, (cudaStream_t)stream);
}

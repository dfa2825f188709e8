// Graph construction kernels + batch plan.
// All integer artefacts are bit-exact against oracle/phoregen_oracle.py.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include "pg_plan.h"

// ---------------------------------------------------------------- error text
static thread_local char g_err[512] = "";
void pg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* pg_last_error(void) { return g_err; }
extern "C" int pg_version(void) { return 100; }

// ---------------------------------------------------------------- work space carving
namespace {
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* b) : base((char*)b), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Sizes {
    long long N, Nl, P, Eb, Ek, E3;
    int max_n, max_p, max_ng, min_n;
};

Sizes plan_sizes(int G, const int32_t* na, const int32_t* np) {
    Sizes s{0, 0, 0, 0, 0, 0, 0, 0, 0, 1 << 30};
    for (int g = 0; g < G; g++) {
        long long n = na[g], p = np[g], ng = n + p;
        s.Nl += n; s.P += p; s.N += ng;
        s.Eb += n * (n - 1);
        s.Ek += ng * std::min<long long>(PG_KNN, ng - 1);
        s.E3 += n * (n - 1) * std::max<long long>(n - 2, 0);
        s.max_n = std::max<int>(s.max_n, (int)n);
        s.min_n = std::min<int>(s.min_n, (int)n);
        s.max_p = std::max<int>(s.max_p, (int)p);
        s.max_ng = std::max<int>(s.max_ng, (int)ng);
    }
    return s;
}

// carve every buffer; with base == nullptr only the total size is computed
void carve(Carver& c, const Sizes& s, int G, PgPlan* pl) {
    auto I = [&](size_t n) { return c.take<int>(n); };
    auto L = [&](size_t n) { return c.take<long long>(n); };
    auto F = [&](size_t n) { return c.take<float>(n); };
    int* ctx_off = I(G + 1); int* lig_off = I(G + 1); int* ph_off = I(G + 1);
    long long* eoff = L(G + 1); long long* koff = L(G + 1); long long* t3off = L(G + 1);
    int* g_n = I(G); int* g_p = I(G);
    int* node_graph = I(s.N); int* lig_graph = I(s.Nl); int* ph_graph = I(s.P);
    int* perm = I(s.Eb); int* inv_perm = I(s.Eb); int* edge_graph = I(s.Eb);
    int* esrc = I(s.Eb); int* edst = I(s.Eb);
    int* flag = I(4);
    int* btile_off = I(G + 1); int* btile_graph = I(s.Nl);
    float* h = F(s.N * 128); float* x = F(s.N * 3 + 4); float* hb = F(s.Eb * 128);
    float* nbuf = F(s.N * 1920);
    float* qn1 = F(s.N * 128); float* qn2 = F(s.N * 128);
    float* o1 = F(s.N * 128); float* o2 = F(s.N * 128);
    float* dx1 = F(s.N * 3 + 4); float* dx2 = F(s.N * 3 + 4);
    float* ebuf = F(s.Eb * 640);
    float* qt = F(s.Eb * 128);
    float* rbuf = F(s.Eb * 256); float* pbuf2 = F(s.Eb * 256 + 64); float* abuf = F(s.Ek * 16 + s.N * 16);
    float* ew = F(s.Ek); float* comb = F(s.N * 3 + 4);
    int* knn_src = I(s.Ek);
    float* pbuf = F(s.P * 640); float* pq = F(s.P * 128); float* pemb = F(s.P * 128);
    float* gcnt = F(G * 4 + 4);
    if (pl) {
        PlanDev& d = pl->d;
        d.ctx_off = ctx_off; d.lig_off = lig_off; d.ph_off = ph_off; d.eoff = eoff; d.koff = koff; d.t3off = t3off;
        d.g_n = g_n; d.g_p = g_p; d.node_graph = node_graph; d.lig_graph = lig_graph; d.ph_graph = ph_graph;
        d.perm = perm; d.edge_graph = edge_graph; d.esrc_node = esrc; d.edst_node = edst;
        d.btile_off = btile_off; d.btile_graph = btile_graph;
        pl->inv_perm = inv_perm; pl->flag = flag;
        pl->h = h; pl->x = x; pl->hb = hb; pl->nbuf = nbuf; pl->qn1 = qn1; pl->qn2 = qn2; pl->o1 = o1; pl->o2 = o2;
        pl->dx1 = dx1; pl->dx2 = dx2; pl->ebuf = ebuf; pl->qt = qt; pl->rbuf = rbuf; pl->pbuf2 = pbuf2; pl->abuf = abuf; pl->ew = ew;
        pl->comb = comb; pl->knn_src = knn_src; pl->pbuf = pbuf; pl->pq = pq; pl->pemb = pemb; pl->gcnt = gcnt;
    }
}
}  // namespace

extern "C" int64_t pg_plan_workspace_bytes(int G, const int32_t* na, const int32_t* np) {
    if (G <= 0 || !na || !np) return PG_EINVAL;
    Sizes s = plan_sizes(G, na, np);
    Carver c(nullptr);
    carve(c, s, G, nullptr);
    return (int64_t)c.off + 256;
}

// ---------------------------------------------------------------- perm from a caller-supplied edge list
// ref edge e = (src, dst) in ligand numbering -> internal slot; flag[0] counts invalid edges, and every
// internal slot must be hit exactly once (inv_perm initialised to -1, atomicCAS).
__global__ void perm_from_edges_kernel(PlanDev d, const int64_t* __restrict__ ei, int* perm, int* inv_perm,
                                       int* edge_graph, int* flag) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= d.Eb) return;
    long long s = ei[e], t = ei[d.Eb + e];
    if (s < 0 || t < 0 || s >= d.Nl || t >= d.Nl || s == t) { atomicAdd(flag, 1); perm[e] = 0; return; }
    int g = d.lig_graph[t];
    if (d.lig_graph[s] != g) { atomicAdd(flag, 1); perm[e] = 0; return; }
    int n = d.g_n[g];
    int i = (int)t - d.lig_off[g], j = (int)s - d.lig_off[g];
    long long slot = d.eoff[g] + (long long)i * (n - 1) + (j - (j > i));
    perm[e] = (int)slot;
    edge_graph[e] = g;
    if (atomicCAS(&inv_perm[slot], -1, (int)e) != -1) atomicAdd(flag, 1);
}

extern "C" int pg_plan_create(PgPlan** out, int G, const int32_t* na, const int32_t* np, int edge_order,
                              const int64_t* d_ref_edge_index, void* d_workspace, int64_t workspace_bytes,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out || G <= 0 || !na || !np || !d_workspace) { pg_set_error("pg_plan_create: bad argument"); return PG_EINVAL; }
    if (edge_order < 0 || edge_order > 2 || (edge_order == 2 && !d_ref_edge_index)) {
        pg_set_error("pg_plan_create: bad edge_order"); return PG_EINVAL;
    }
    for (int g = 0; g < G; g++) {
        if (na[g] < 2 || na[g] > PG_MAX_ATOMS) { pg_set_error("graph %d: %d ligand atoms outside [2,%d]", g, na[g], PG_MAX_ATOMS); return PG_ELIMIT; }
        if (np[g] < 1 || na[g] + np[g] > PG_MAX_CTX_NODES) { pg_set_error("graph %d: %d phore nodes unsupported (limit %d nodes per graph)", g, np[g], PG_MAX_CTX_NODES); return PG_ELIMIT; }
    }
    int64_t need = pg_plan_workspace_bytes(G, na, np);
    if (workspace_bytes < need) { pg_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need); return PG_EWORKSPACE; }
    if (((uintptr_t)d_workspace & 255) != 0) { pg_set_error("workspace must be 256-byte aligned"); return PG_EINVAL; }
    Sizes s = plan_sizes(G, na, np);
    if (s.Eb * 256 >= (1LL << 40) || s.Eb >= (1LL << 31) || s.Ek >= (1LL << 31)) { pg_set_error("batch too large for 32-bit edge ids"); return PG_ELIMIT; }

    PgPlan* pl = new PgPlan();
    pl->edge_order = edge_order;
    pl->launches = 0;
    pl->timing = 0;
    Carver c(d_workspace);
    carve(c, s, G, pl);
    PlanDev& d = pl->d;
    d.G = G; d.N = (int)s.N; d.Nl = (int)s.Nl; d.P = (int)s.P; d.Eb = s.Eb; d.Ek = s.Ek; d.E3 = s.E3;
    d.max_n = s.max_n; d.max_p = s.max_p; d.max_ng = s.max_ng; d.min_n = G > 0 ? s.min_n : 0;

    pl->n.assign(na, na + G); pl->p.assign(np, np + G);
    pl->ctx_off.assign(G + 1, 0); pl->lig_off.assign(G + 1, 0); pl->ph_off.assign(G + 1, 0);
    pl->eoff.assign(G + 1, 0); pl->koff.assign(G + 1, 0); pl->t3off.assign(G + 1, 0);
    for (int g = 0; g < G; g++) {
        long long n = na[g], p = np[g], ng = n + p;
        pl->ctx_off[g + 1] = pl->ctx_off[g] + (int)ng;
        pl->lig_off[g + 1] = pl->lig_off[g] + (int)n;
        pl->ph_off[g + 1] = pl->ph_off[g] + (int)p;
        pl->eoff[g + 1] = pl->eoff[g] + n * (n - 1);
        pl->koff[g + 1] = pl->koff[g] + ng * std::min<long long>(PG_KNN, ng - 1);
        pl->t3off[g + 1] = pl->t3off[g] + n * (n - 1) * std::max<long long>(n - 2, 0);
    }
    std::vector<int> node_graph(s.N), lig_graph(s.Nl), ph_graph(s.P), perm(s.Eb), inv_perm(s.Eb), edge_graph(s.Eb),
        esrc(s.Eb), edst(s.Eb);
    std::vector<int> btile_off(G + 1, 0), btile_graph;
    for (int g = 0; g < G; g++) {
        const int apt = std::max(pg_bond_atoms_per_tile(na[g]), 1);
        const int nt = na[g] - 1 > 32 ? (na[g] + apt - 1) / apt : 0;     // single-quarter molecules take the packed kernel
        btile_off[g + 1] = btile_off[g] + nt;
        btile_graph.insert(btile_graph.end(), nt, g);
    }
    d.nbt = btile_off[G];
    for (int g = 0; g < G; g++) {
        int n = na[g], p = np[g];
        for (int v = pl->ctx_off[g]; v < pl->ctx_off[g + 1]; v++) node_graph[v] = g;
        for (int v = pl->lig_off[g]; v < pl->lig_off[g + 1]; v++) lig_graph[v] = g;
        for (int v = pl->ph_off[g]; v < pl->ph_off[g + 1]; v++) ph_graph[v] = g;
        int ctx0 = pl->ctx_off[g] + p;
        long long e0 = pl->eoff[g];
        for (int i = 0; i < n; i++)
            for (int t = 0; t < n - 1; t++) {
                int j = t + (t >= i);
                long long e = e0 + (long long)i * (n - 1) + t;
                esrc[e] = ctx0 + j; edst[e] = ctx0 + i;
            }
        if (edge_order == 1) {
            for (long long r = 0; r < (long long)n * (n - 1); r++) { perm[e0 + r] = (int)(e0 + r); inv_perm[e0 + r] = (int)(e0 + r); edge_graph[e0 + r] = g; }
        } else if (edge_order == 0) {   // utils/sample_utils.py:47-48: triu pairs (a<b) as (src=a,dst=b), then flipped
            long long H = (long long)n * (n - 1) / 2, hcnt = 0;
            for (int a = 0; a < n; a++)
                for (int b = a + 1; b < n; b++, hcnt++) {
                    long long up = e0 + (long long)b * (n - 1) + a;            // src=a,dst=b: a<b -> rank a
                    long long dn = e0 + (long long)a * (n - 1) + (b - 1);      // src=b,dst=a: b>a -> rank b-1
                    perm[e0 + hcnt] = (int)up; inv_perm[up] = (int)(e0 + hcnt);
                    perm[e0 + H + hcnt] = (int)dn; inv_perm[dn] = (int)(e0 + H + hcnt);
                    edge_graph[e0 + hcnt] = g; edge_graph[e0 + H + hcnt] = g;
                }
        }
    }
    auto up = [&](const void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync((void*)dst, src, bytes, cudaMemcpyHostToDevice, stream); };
    cudaError_t e = cudaSuccess;
    e = e ? e : up(d.ctx_off, pl->ctx_off.data(), (G + 1) * 4);
    e = e ? e : up(d.lig_off, pl->lig_off.data(), (G + 1) * 4);
    e = e ? e : up(d.ph_off, pl->ph_off.data(), (G + 1) * 4);
    e = e ? e : up(d.eoff, pl->eoff.data(), (G + 1) * 8);
    e = e ? e : up(d.koff, pl->koff.data(), (G + 1) * 8);
    e = e ? e : up(d.t3off, pl->t3off.data(), (G + 1) * 8);
    e = e ? e : up(d.g_n, pl->n.data(), G * 4);
    e = e ? e : up(d.g_p, pl->p.data(), G * 4);
    e = e ? e : up(d.node_graph, node_graph.data(), s.N * 4);
    e = e ? e : up(d.lig_graph, lig_graph.data(), s.Nl * 4);
    e = e ? e : up(d.ph_graph, ph_graph.data(), s.P * 4);
    e = e ? e : up(d.esrc_node, esrc.data(), s.Eb * 4);
    e = e ? e : up(d.edst_node, edst.data(), s.Eb * 4);
    e = e ? e : up(d.btile_off, btile_off.data(), (G + 1) * 4);
    e = e ? e : up(d.btile_graph, btile_graph.data(), (size_t)d.nbt * 4);
    e = e ? e : cudaMemsetAsync(pl->flag, 0, 16, stream);
    if (edge_order != 2) {
        e = e ? e : up(d.perm, perm.data(), s.Eb * 4);
        e = e ? e : up(pl->inv_perm, inv_perm.data(), s.Eb * 4);
        e = e ? e : up(d.edge_graph, edge_graph.data(), s.Eb * 4);
    } else {
        e = e ? e : cudaMemsetAsync((void*)pl->inv_perm, 0xff, s.Eb * 4, stream);
        if (!e && s.Eb > 0) {
            perm_from_edges_kernel<<<(unsigned)((s.Eb + 255) / 256), 256, 0, stream>>>(
                d, d_ref_edge_index, (int*)d.perm, (int*)pl->inv_perm, (int*)d.edge_graph, pl->flag);
            e = cudaGetLastError();
        }
    }
    int hflag = 0;
    e = e ? e : cudaMemcpyAsync(&hflag, pl->flag, 4, cudaMemcpyDeviceToHost, stream);
    e = e ? e : cudaStreamSynchronize(stream);   // host vectors above must outlive the copies
    if (e != cudaSuccess) { pg_set_error("pg_plan_create: %s", cudaGetErrorString(e)); delete pl; return PG_ECUDA; }
    if (hflag != 0) {
        pg_set_error("bond_index is not the complete directed ligand graph of every molecule (%d offending edges)", hflag);
        delete pl; return PG_EINVAL;
    }
    *out = pl;
    return PG_OK;
}

extern "C" void pg_plan_destroy(PgPlan* p) { delete p; }
extern "C" int64_t pg_plan_num_ligand_atoms(const PgPlan* p) { return p->d.Nl; }
extern "C" int64_t pg_plan_num_phore_nodes(const PgPlan* p) { return p->d.P; }
extern "C" int64_t pg_plan_num_bond_edges(const PgPlan* p) { return p->d.Eb; }
extern "C" int64_t pg_plan_num_knn_edges(const PgPlan* p) { return p->d.Ek; }
extern "C" int64_t pg_plan_num_triplets(const PgPlan* p) { return p->d.E3; }
extern "C" const int32_t* pg_plan_ligand_graph(const PgPlan* p) { return p->d.lig_graph; }
extern "C" const int32_t* pg_plan_edge_graph(const PgPlan* p) { return p->d.edge_graph; }
extern "C" int64_t pg_plan_kernel_launches(const PgPlan* p) { return p->launches; }
extern "C" int pg_plan_timing_enable(PgPlan* p, int on) {
    for (int c = 0; c < KC_COUNT; c++) { for (auto e : p->tev[c]) cudaEventDestroy(e); p->tev[c].clear(); }
    p->timing = on;
    return PG_OK;
}
extern "C" int pg_plan_timing_read(PgPlan* p, int kclass, double* ms_total, int64_t* launches) {
    if (kclass < 0 || kclass >= KC_COUNT || !ms_total || !launches) { pg_set_error("pg_plan_timing_read: bad argument"); return PG_EINVAL; }
    double tot = 0; auto& v = p->tev[kclass];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        PG_CUDA_CHECK(cudaEventSynchronize(v[i + 1]));
        float ms = 0; PG_CUDA_CHECK(cudaEventElapsedTime(&ms, v[i], v[i + 1])); tot += ms;
    }
    *ms_total = tot; *launches = (int64_t)(v.size() / 2);
    return PG_OK;
}

// ---------------------------------------------------------------- G1 export
__global__ void export_edges_kernel(PlanDev d, const int* __restrict__ inv_perm, int64_t* ei, int64_t* eb) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // internal slot
    if (e >= d.Eb) return;
    int r = inv_perm[e];
    int g = d.node_graph[d.edst_node[e]];
    int ctx0 = d.ctx_off[g] + d.g_p[g];
    ei[r] = d.esrc_node[e] - ctx0 + d.lig_off[g];
    ei[d.Eb + r] = d.edst_node[e] - ctx0 + d.lig_off[g];
    eb[r] = g;
}
extern "C" int pg_plan_export_bond_edges(const PgPlan* p, int64_t* ei, int64_t* eb, void* stream) {
    if (p->d.Eb == 0) return PG_OK;
    export_edges_kernel<<<(unsigned)((p->d.Eb + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p->d, p->inv_perm, ei, eb);
    PG_LAUNCH_CHECK();
    const_cast<PgPlan*>(p)->launches++;
    return PG_OK;
}

// ---------------------------------------------------------------- B1 export (uni_denoiser.py:101-121)
// One thread per (reference edge, r): triplet r of edge e=(j->i) is the r-th k ascending with k != i, j.
__global__ void export_triplets_kernel(PlanDev d, const int* __restrict__ inv_perm, int64_t* oi, int64_t* oj,
                                       int64_t* ok, int64_t* okj, int64_t* oji) {
    long long e = blockIdx.x;                 // reference-order edge id
    if (e >= d.Eb) return;
    int slot = d.perm[e];
    int g = d.node_graph[d.edst_node[slot]];
    int n = d.g_n[g];
    if (n < 3) return;
    int ctx0 = d.ctx_off[g] + d.g_p[g];
    int i = d.edst_node[slot] - ctx0, j = d.esrc_node[slot] - ctx0;
    long long base = d.t3off[g] + (e - d.eoff[g]) * (long long)(n - 2);
    int lo = min(i, j), hi = max(i, j);
    for (int r = threadIdx.x; r < n - 2; r += blockDim.x) {
        int k = r; if (k >= lo) k++; if (k >= hi) k++;
        long long kj_slot = d.eoff[g] + (long long)j * (n - 1) + (k - (k > j));   // edge k -> j
        long long o = base + r;
        oi[o] = ctx0 + i; oj[o] = ctx0 + j; ok[o] = ctx0 + k;
        okj[o] = inv_perm[kj_slot]; oji[o] = e;
    }
}
extern "C" int pg_plan_export_triplets(const PgPlan* p, int64_t* oi, int64_t* oj, int64_t* ok, int64_t* okj,
                                       int64_t* oji, void* stream) {
    if (p->d.Eb == 0) return PG_OK;
    export_triplets_kernel<<<(unsigned)p->d.Eb, 64, 0, (cudaStream_t)stream>>>(p->d, p->inv_perm, oi, oj, ok, okj, oji);
    PG_LAUNCH_CHECK();
    const_cast<PgPlan*>(p)->launches++;
    return PG_OK;
}

// ---------------------------------------------------------------- K1 / S3: brute-force kNN, one warp per query
// Squared distance in fp32 as ((dx*dx + dy*dy) + dz*dz) without FMA contraction; candidates ranked by
// (distance, index); the k+1 best including the query itself are taken and the query is dropped
// (torch_cluster knn + PyG knn_graph(loop=False) semantics).
// MODE 0: all context nodes, k = 32, writes knn_src (int32, internal CSR) and/or int64 edge_index.
// MODE 1: ligand atoms only, k = 3, writes comb_norm (common.py:300-314) and/or int64 edge_index.
template <int MODE>
__global__ void knn_kernel(PlanDev d, const float* __restrict__ x, const float* __restrict__ phore_norm,
                           int* __restrict__ knn_src, float* __restrict__ comb, int64_t* __restrict__ ei,
                           long long E_total) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * wpb + warp;
    if (q >= d.N) return;
    float* dist = sm + (size_t)warp * (d.max_ng + 8);
    int* nb = (int*)(dist + d.max_ng);   // [<=4] selected neighbours (MODE 1)
    const int g = d.node_graph[q];
    const int p = d.g_p[g], n = d.g_n[g];
    const int c0 = d.ctx_off[g] + (MODE == 1 ? p : 0);     // first candidate
    const int nc = MODE == 1 ? n : n + p;                   // number of candidates
    const int K = MODE == 1 ? 3 : PG_KNN;
    const int kk = min(K, nc - 1);
    if (MODE == 1 && q < c0) {                              // pharmacophore row: comb_norm = phore_norm
        if (comb && lane < 3) comb[(size_t)q * 3 + lane] = phore_norm[(size_t)(d.ph_off[g] + q - d.ctx_off[g]) * 3 + lane];
        return;
    }
    const float qx = x[(size_t)q * 3], qy = x[(size_t)q * 3 + 1], qz = x[(size_t)q * 3 + 2];
    for (int m = lane; m < nc; m += 32) {
        const float* c = x + (size_t)(c0 + m) * 3;
        float dx = qx - c[0], dy = qy - c[1], dz = qz - c[2];
        dist[m] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    if (lane < 4) nb[lane] = -1;
    __syncwarp();
    const int qs = q - c0;
    int rs = 0;                                             // rank of the query itself (distance 0)
    for (int m = 0; m < nc; m++) rs += (dist[m] < dist[qs]) || (dist[m] == dist[qs] && m < qs);
    long long ebase;
    if (MODE == 0) ebase = d.koff[g] + (long long)(q - d.ctx_off[g]) * kk;
    else {   // k=3 edge list: graphs before g contribute n_h * min(3, n_h - 1)
        ebase = 0;   // computed below only when exporting
    }
    for (int j = lane; j < nc; j += 32) {
        if (j == qs) continue;
        const float dj = dist[j];
        int rank = 0;
        for (int m = 0; m < nc; m++) { float dm = dist[m]; rank += (dm < dj) || (dm == dj && m < j); }
        const int pos = rank - (rank > rs);
        if (pos < kk) {
            if (MODE == 0) {
                if (knn_src) knn_src[ebase + pos] = c0 + j;
                if (ei) { ei[ebase + pos] = c0 + j; ei[E_total + ebase + pos] = q; }
            } else {
                nb[pos] = j;
            }
        }
    }
    if (MODE == 1) {
        __syncwarp();
        if (comb && lane < 3) {       // scatter-mean of neighbour positions (edge order = ascending distance) minus x
            float s = 0.f;
            for (int t = 0; t < kk; t++) s += x[(size_t)(c0 + nb[t]) * 3 + lane];
            comb[(size_t)q * 3 + lane] = s / (float)max(kk, 1) - x[(size_t)q * 3 + lane];
        }
        if (ei && lane == 0) {
            long long eb = 0;
            for (int h = 0; h < g; h++) { int nh = d.g_n[h]; eb += (long long)nh * min(3, nh - 1); }
            eb += (long long)qs * kk;
            // ligand numbering for the exported k=3 graph (the reference calls knn_graph on x[mask_ligand])
            for (int t = 0; t < kk; t++) { ei[eb + t] = d.lig_off[g] + nb[t]; ei[E_total + eb + t] = d.lig_off[g] + qs; }
        }
    }
}

int pg_launch_knn(PgPlan* p, const float* x, const float* phore_norm, int mode, int* knn_src, float* comb,
                  int64_t* ei, cudaStream_t stream) {
    const PlanDev& d = p->d;
    const int wpb = 4;
    size_t smem = (size_t)wpb * (d.max_ng + 8) * sizeof(float);
    unsigned grid = (unsigned)((d.N + wpb - 1) / wpb);
    PgTimed timed(p, KC_GRAPH, stream);
    if (mode == 0) {
        knn_kernel<0><<<grid, wpb * 32, smem, stream>>>(d, x, phore_norm, knn_src, comb, ei, d.Ek);
    } else {
        long long E = 0;
        for (int g = 0; g < d.G; g++) E += (long long)p->n[g] * std::min(3, p->n[g] - 1);
        knn_kernel<1><<<grid, wpb * 32, smem, stream>>>(d, x, phore_norm, knn_src, comb, ei, E);
    }
    PG_LAUNCH_CHECK();
    p->launches++;
    return PG_OK;
}

extern "C" int pg_knn_graph(const PgPlan* p, const float* d_x, int mode, int64_t* d_edge_index, void* stream) {
    if (mode != 0 && mode != 1) { pg_set_error("pg_knn_graph: mode must be 0 or 1"); return PG_EINVAL; }
    return pg_launch_knn(const_cast<PgPlan*>(p), d_x, nullptr, mode, nullptr, nullptr, d_edge_index, (cudaStream_t)stream);
}

// Host-side plan / model structs shared by the translation units of libphoregen_b200.
#pragma once
#include <vector>
#include "pg_common.cuh"

struct PgPlan {
    PlanDev d;
    std::vector<int> n, p, ctx_off, lig_off, ph_off;
    std::vector<long long> eoff, koff, t3off;
    int edge_order;
    const int* inv_perm;  // device [Eb] internal -> reference order
    // activation work space (device)
    float *h, *x, *hb;            // [N,128] [N,3] [Eb,128] state inside one forward (internal edge order)
    float* nbuf;                  // [N,1920] node GEMM outputs
    float *qn1, *qn2;             // [N,128] per-node queries (kNN / bond sub-layers)
    float *o1, *o2;               // [N,128] attention outputs
    float *dx1, *dx2;             // [N,3]
    float* ebuf;                  // [Eb,640] edge GEMM outputs
    float* qt;                    // [Eb,128] per-edge triplet queries / head hidden
    float* abuf;                  // [Ek,16] + [N,16] attention weights of the tcgen05 kNN attention (key pass -> value pass)
    float* rbuf;                  // [2][Eb,128] r_ji slice of the triplet MLPs as bf16 hi/lo operand images (tcgen05 triplet kernel)
    float* pbuf2;                 // [2][Eb,128] per-edge first-Linear partials P[k->j] as bf16 hi/lo operand images (tcgen05 triplet kernel)
    float *ew, *comb;             // [Ek], [N,3]
    int* knn_src;                 // [Ek]
    float *pbuf, *pq, *pemb;      // phore encoder: [P,640], [P,128], [P,128]
    float* gcnt;                  // [G] scratch (guidance)
    int* flag;                    // device error flag
    long long launches;
    // optional per-kernel-class device timing (eager mode only; bench.py's roofline leg)
    int timing;
    std::vector<cudaEvent_t> tev[8];   // pairs (start, stop) per class
};

enum PgKernelClass { KC_GEMM = 0, KC_KNN_ATTN, KC_BOND_ATTN, KC_TRIP, KC_GRAPH, KC_OTHER, KC_COUNT };

struct PgTimed {   // RAII: CUDA events around one launch on the launching stream
    PgPlan* p; int c; cudaStream_t s;
    PgTimed(PgPlan* p_, int c_, cudaStream_t s_) : p(p_), c(c_), s(s_) {
        if (p->timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); p->tev[c].push_back(e); }
    }
    ~PgTimed() {
        if (p->timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); p->tev[c].push_back(e); }
    }
};

// weight slots -------------------------------------------------------------------------------------
enum PgAttnKind { AK_NODE_KNN = 0, AK_NODE_BOND, AK_TRIP, AK_POS_KNN, AK_POS_BOND, AK_COUNT };

struct PgSlotDesc {
    char name[48];
    long long numel;
};

struct PgModel {
    const float* blob;
    std::vector<long long> off;
    const float* w(int slot) const { return blob + off[slot]; }
};

int pg_slot_index(const char* name);            // -1 if absent
const std::vector<PgSlotDesc>& pg_slot_table();

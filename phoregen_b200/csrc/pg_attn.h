#pragma once
#include "pg_common.cuh"

// Weights of one attention sub-layer (pointers into the packed blob).
struct AttnW {
    const float *tab_k, *tab_v;                   // kNN: [4][24][128]; phore encoder: [128] distance column
    const float *lnk_g, *lnk_b, *lnv_g, *lnv_b;   // LayerNorm of the key / value MLPs
    const float *lnk_bf, *lnv_bf, *fold;          // tensor-core kernels: beta, or beta / gamma where gamma > 0 was folded into
                                                  //   the bf16 W2 images; fold = [fold_k, fold_v, -, -] (1.0 = folded)
    const float *w2k, *b2k;                       // key MLP second Linear   [128 out][128 in], [128]
    const float *w2v, *b2v;                       // value MLP second Linear [128|16 out][128 in], [128|16]
};

struct NodeCols {            // column offsets inside the node-GEMM output (row stride lda)
    const float* A; long long lda;
    int dst_k, src_k, dst_v, src_v;
};

struct KnnAttnArgs {
    PlanDev d;
    const float* x;          // [N,3] context coordinates (FEAT 0) / pharmacophore positions [P,3] (FEAT 1)
    const float* comb;       // [N,3] direction vectors (FEAT 0)
    const int* knn_src;      // [Ek]
    const float* ew;         // [Ek]
    NodeCols nc;
    const float* q;          // [rows,128]
    AttnW w;
    float* out;              // node mode [rows,128]; pos mode [rows,3]
    int maxr;
};

struct BondAttnArgs {
    PlanDev d;
    const float* x;          // [N,3]
    NodeCols nc;
    const float* B; long long ldb; int b_k, b_v;   // per-edge partial pre-activations (edge GEMM output)
    const float* q;          // [N,128]
    AttnW w;
    float* out;              // node mode [N,128]; pos mode [N,3]
    int maxr;
    int tc_max_rows;         // > 0: atoms with 1 <= n-1 <= tc_max_rows are skipped (they ran on the tcgen05 kernel)
};

// tcgen05 version of the bond-graph attention (pg_bond_tc.cu): same inputs plus the bf16 hi/lo second-Linear weights
struct BondTcArgs {
    PlanDev d;
    const float* x;
    NodeCols nc;
    const float* B; long long ldb; int b_k, b_v;
    const float* q;
    AttnW w;
    const uint16_t *w2k_bf, *w2v_bf;       // [hi|lo][128][128] and [hi|lo][128 | 16][128] bf16, K-major
    const uint16_t* w2k_h;                 // [128][128] fp16, K-major (single-pass key MLP)
    int key_bf16x3;                        // 1: key MLP in bf16x3 (PG_KEY=bf16x3, A/B only)
    float* out;
};
// tcgen05 version of the kNN-graph attention (pg_knn_tc.cu): key pass + value pass
struct KnnTcArgs {
    PlanDev d;
    const float* x;          // [N,3]
    const float* comb;       // [N,3]
    const int* knn_src;      // [Ek]
    const float* ew;         // [Ek]
    NodeCols nc;
    const float* q;          // [N,128]
    AttnW w;
    const uint16_t *w2k_bf, *w2v_bf;       // [hi|lo][128][128] and [hi|lo][128 | 16][128] bf16, K-major
    const uint16_t* w2k_h;                 // [128][128] fp16, K-major (single-pass key MLP)
    int key_bf16x3;                        // 1: key MLP in bf16x3 (PG_KEY=bf16x3, A/B only)
    const uint16_t *tabk_bf, *tabv_bf;     // [hi|lo][128][96] bf16: first-Linear slices of the 4 edge types x 24 features
    float* alpha;            // scratch [Ek,16]: softmax weights times e_w
    float* alpha_sum;        // scratch [N,16]: their per-head sums
    float* out;              // node mode [N,128]; pos mode [N,3]
};
int pg_launch_knn_tc(const KnnTcArgs& a, int pos, int num_sms, cudaStream_t s);
constexpr int PG_BOND_TC_SINGLE_CHUNK_ROWS = 32;    // n-1 <= 32: one TMEM lane quarter per segment, else 32-row chunks
int pg_launch_bond_tc(const BondTcArgs& a, int pos, int num_sms, cudaStream_t s);

struct TripArgs {
    PlanDev d;
    const float* x;
    const float* T; long long ldt; int t_k, t_v;   // h_bond @ Wb (edge GEMM output)
    const float* H; long long ldh; int hk_k, hj_k, hk_v, hj_v;   // node GEMM output (h @ Whk, h @ Whj + b1)
    const float* q;          // [Eb,128] per-edge query
    const float *wrkj, *wrji, *wa;   // [20][256], [20][256], [13][256]  (k | v along the 256 axis)
    AttnW w;
    float* hb;               // [Eb,128] updated in place: hb += attention output
    int maxr, maxn;
    int min_atoms;           // units of molecules with fewer atoms are skipped (they ran on the tcgen05 kernel)
};

// tcgen05 version of the triplet layer (csrc/pg_trip_tc.cu); segments of up to 32 rows are one TMEM lane quarter, longer
// ones are cut into 32-row chunks with an on-line softmax across the chunks (any n up to PG_MAX_ATOMS)
struct TripTcArgs {
    PlanDev d;
    const float* x;
    const float* T; long long ldt; int t_k, t_v;
    const float* H; long long ldh; int hk_k, hj_k, hk_v, hj_v;
    const float* q;                        // [Eb,128]
    float* R;                              // [2][Eb,128] work space: smear(d_e) @ Wrji as bf16 hi/lo operand images (key | value), indexed by the source atom's unit
    float* P;                              // [2][Eb,128] work space: per-edge partial of the first Linear (k->j role) as bf16 hi/lo operand images (key | value)
    const float *wrkj, *wrji;              // [20][256] fp32
    const uint16_t *w2k_bf, *w2v_bf;       // [hi|lo][128][128] bf16, K-major
    const uint16_t* w2k_h;                 // [128][128] fp16, K-major: the key MLP's second Linear at single precision-16
    const uint16_t* wa_bf;                 // angle slab image [mlp][hi|lo][half][16][64] bf16, MN-major SW128 (weights.angle_slab_image)
    int flags;                             // switches (PG_TRIP_FLAGS): bit 0 = shuffle-butterfly softmax instead of REDUX; bit 1 = key MLP in bf16x3
    const float *lnk_g, *lnk_b, *lnv_g, *lnv_b, *b2k, *b2v;
    const float *lnk_bf, *lnv_bf, *fold;   // beta (/ gamma where folded into the W2 images), fold flags (see AttnW)
    float* hb;
    int maxn;
};
constexpr int PG_TRIP_TC_SINGLE_CHUNK_ATOMS = 33;   // n - 1 <= 32 staged rows: the single-chunk instantiation serves the molecule
int pg_launch_trip_pr(const TripTcArgs& a, cudaStream_t s);               // per-edge partials P, R (elementwise, HBM-bound)
int pg_launch_trip_tc(const TripTcArgs& a, int num_sms, cudaStream_t s);  // the tcgen05 triplet kernel proper
size_t pg_trip_tc_smem(int maxn);

int pg_launch_knn_attn(const KnnAttnArgs& a, int feat, int pos, cudaStream_t s);
int pg_launch_bond_attn(const BondAttnArgs& a, int pos, cudaStream_t s);
int pg_launch_trip(const TripArgs& a, cudaStream_t s);
int pg_launch_edge_weight(const PlanDev& d, const float* x, const int* knn_src, const float* w1t, const float* b1,
                          const float* g, const float* b, const float* w2, const float* b2, float* ew, cudaStream_t s);

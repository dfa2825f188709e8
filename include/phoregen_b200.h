/*
 * phoregen_b200 — C ABI of the B200-native PhoreGen sampling hot path.
 *
 * Drop-in boundary: the reference has no FFI layer; its seam is the Python class
 * interface (models/__init__.py:5-35, models/uni_denoiser.py:396-430, models/diffusion.py:175-246,
 * 390-525) plus the checkpoint state_dict.  The Python mirror in `phoregen_b200/` keeps those
 * signatures and calls the entry points below through ctypes.  Signatures use only plain pointers,
 * sizes and an opaque `void* stream` (a cudaStream_t); no torch types.
 *
 * Conventions
 *   - every `d_*` pointer is DEVICE memory, every `h_*` pointer is HOST memory;
 *   - all functions return 0 on success, a negative PG_E* code otherwise; `pg_last_error()` gives the text;
 *   - no function allocates device memory: work space is passed in (`pg_plan_workspace_bytes`);
 *   - all launches are stream-ordered on `stream`; no function synchronises except the ones documented
 *     as "synchronous" (plan creation reads one validation flag back);
 *   - allocation failures reported by the caller must surface to Python as RuntimeError containing
 *     "out of memory" (reference callers pattern-match that: sample_all.py:96, run/run.py:145).
 */
#ifndef PHOREGEN_B200_H
#define PHOREGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_EINVAL -1      /* bad argument / unsupported topology */
#define PG_ECUDA -2       /* CUDA runtime error (see pg_last_error) */
#define PG_EWORKSPACE -3  /* workspace too small */
#define PG_ELIMIT -4      /* a compiled-in limit (atoms / pharmacophore nodes per graph) exceeded */

#define PG_HIDDEN 128
#define PG_HEADS 16
#define PG_KNN 32
#define PG_NUM_LAYERS 6
#define PG_NODE_CLASSES 12
#define PG_EDGE_CLASSES 6
#define PG_PHORE_FEAT 18
#define PG_MAX_ATOMS 128       /* ligand atoms per molecule (reference caps at 78: diffusion.py:30) */
#define PG_MAX_CTX_NODES 384   /* ligand + pharmacophore nodes per graph */

typedef struct PgModel PgModel;
typedef struct PgPlan PgPlan;

int pg_version(void);
const char* pg_last_error(void);

/* ---------------------------------------------------------------- packed weights
 * The Python host packs the reference state_dict (641 keys, SURVEY.md §8(b) "Checkpoint") into one
 * fp32 device blob; slot i starts at d_blob + h_offsets[i] (in floats).  Slot names/sizes are
 * self-described so the packer and the library cannot drift apart.
 * Contract of the slot contents (phoregen_b200/weights.py is the reference packer):
 *   - first Linears of the attention MLPs are split along their concatenated input (DESIGN.md "Factorisation") and every
 *     128-column block of them ("L*.n1/e1/n2/e2.wt|b", "L*.nk|pk.tab_k|v", "L*.tr.wrkj|wrji|wa") has its mean over the
 *     128 output channels removed: the LayerNorm that follows subtracts it anyway, and the tensor-core attention kernels
 *     rely on mean-free pre-activations (they only accumulate the second moment);
 *   - "<S>.w2k.bf" / "<S>.w2v.bf" are bf16 hi|lo images [2][out][128] (LayerNorm gain folded in where "<S>.fold" says so),
 *     "*.wt.bf" the pre-swizzled 64-column GEMM images, "L*.tr.wa.bf" the angle-slab image of csrc/pg_trip_tc.cu. */
int pg_weight_slot_count(void);
const char* pg_weight_slot_name(int slot);      /* e.g. "L3.trip.w2k" or "G.ew.w1t" */
int64_t pg_weight_slot_numel(int slot);
int pg_model_create(PgModel** out, const float* d_blob, const int64_t* h_offsets, int n_slots);
void pg_model_destroy(PgModel* m);

/* ---------------------------------------------------------------- batch plan (static topology of one batch)
 * Replaces, for a whole trajectory: make_edge_data (utils/sample_utils.py:40-54), compose_context index maps
 * (models/common.py:166-208), fully_connect_two_graphs (common.py:329-356) and BondUpdateLayer.triplets
 * (uni_denoiser.py:101-121; arithmetic for complete ligand graphs).
 * Graph g has h_num_phore[g] pharmacophore nodes followed by h_num_atoms[g] ligand atoms in context order.
 * edge_order: 0 = sampling order (triu pairs then flipped), 1 = training order (dst-major); 2 = caller supplies
 * d_ref_edge_index [2,E_b] int64 in ligand numbering (validated to be the complete directed graph). */
int64_t pg_plan_workspace_bytes(int n_graphs, const int32_t* h_num_atoms, const int32_t* h_num_phore);
int pg_plan_create(PgPlan** out, int n_graphs, const int32_t* h_num_atoms, const int32_t* h_num_phore,
                   int edge_order, const int64_t* d_ref_edge_index, void* d_workspace, int64_t workspace_bytes,
                   void* stream);                                   /* synchronous */
void pg_plan_destroy(PgPlan* p);
int64_t pg_plan_num_ligand_atoms(const PgPlan* p);
int64_t pg_plan_num_phore_nodes(const PgPlan* p);
int64_t pg_plan_num_bond_edges(const PgPlan* p);
int64_t pg_plan_num_knn_edges(const PgPlan* p);
int64_t pg_plan_num_triplets(const PgPlan* p);
/* G1: write the reference-order edge list (ligand numbering) and edge->graph map, int64 like the reference. */
int pg_plan_export_bond_edges(const PgPlan* p, int64_t* d_edge_index /*[2,E_b]*/, int64_t* d_edge_batch /*[E_b]*/,
                              void* stream);
/* B1: materialise the five triplet index arrays of BondUpdateLayer.triplets in the reference's order, for the
 * reference-order edge list in CONTEXT numbering (test / interop only; the fused kernels never read them). */
int pg_plan_export_triplets(const PgPlan* p, int64_t* d_idx_i, int64_t* d_idx_j, int64_t* d_idx_k,
                            int64_t* d_idx_kj, int64_t* d_idx_ji, void* stream);

/* ---------------------------------------------------------------- K1: kNN graph (torch_cluster.knn_graph semantics)
 * d_x [N,3] context coordinates.  Output CSR by destination: edges of node v are
 * [v_rowptr[v], v_rowptr[v+1]); d_src[e] = neighbour (ascending distance, ties -> lower index).
 * mode 0: k = 32 over all nodes of the graph; mode 1: k = 3 over ligand atoms only (common.py:301). */
int pg_knn_graph(const PgPlan* p, const float* d_x, int mode, int64_t* d_edge_index /*[2,E]*/, void* stream);

/* ---------------------------------------------------------------- U1: denoiser forward (uni_denoiser.py:396-430)
 * d_h [N,128], d_x [N,3], d_h_bond [E_b,128] in the plan's REFERENCE edge order, d_phore_norm [P,3].
 * Outputs d_h_out [N,128], d_x_out [N,3], d_h_bond_out [E_b,128] (reference edge order). */
int pg_denoiser_forward(const PgModel* m, PgPlan* p, const float* d_h, const float* d_x, const float* d_h_bond,
                        const float* d_phore_norm, float* d_h_out, float* d_x_out, float* d_h_bond_out,
                        void* stream);

/* ---------------------------------------------------------------- E2: pharmacophore embedding + encoder
 * (diffusion.py:186-191).  d_h_phore [P,18], d_pos_phore [P,3] -> d_out [P,128]. */
int pg_phore_encode(const PgModel* m, PgPlan* p, const float* d_h_phore, const float* d_pos_phore, float* d_out,
                    void* stream);

/* ---------------------------------------------------------------- PhoreDiff.forward (diffusion.py:175-246)
 * d_h_node [Nl,12] f32, d_pos [Nl,3], d_h_edge [E_b,6] f32 (reference edge order), d_time_step [G] int64,
 * d_h_phore_emb [P,128] (output of pg_phore_encode; step-invariant in sample()), d_pos_phore, d_phore_norm.
 * Outputs logits_node [Nl,12], pos [Nl,3], logits_edge [E_b,6]. */
int pg_phorediff_forward(const PgModel* m, PgPlan* p, const float* d_h_node, const float* d_pos,
                         const float* d_h_edge, const int64_t* d_time_step, const float* d_h_phore_emb,
                         const float* d_pos_phore, const float* d_phore_norm, float* d_logits_node,
                         float* d_pos_out, float* d_logits_edge, void* stream);

/* ---------------------------------------------------------------- T1-T3: posterior + Gumbel sampling + position update
 * (transition.py:285-315, common.py:425-431, transition.py:44-63).  Categorical part for `rows` rows of K classes:
 *   log_v0 = log_softmax(d_pred); d_log_vt <- q_v_posterior(log_v0, d_log_vt, t) (in place);
 *   class  = argmax(gumbel(u) + d_log_vt) ; d_onehot [rows,K] f32 and d_cls [rows] int32 written.
 * d_uniform [rows,K] supplies the draws; if NULL, Philox4x32-10(seed, row, *d_step_counter) is used.
 * d_row_graph [rows] int32 maps rows to graphs (time_step is per graph).
 * Optional trajectory logging (diffusion.py:418-426,510-512) in compact form: slot (*d_step_counter + 1) of
 * d_traj_cls [T+1,rows] uint8 receives the sampled class (the reference stores its one-hot as f32). */
int pg_categorical_step(int rows, int K, const float* d_pred, float* d_log_vt, const float* d_q_mats,
                        const float* d_tq_onestep, const int64_t* d_time_step, const int32_t* d_row_graph,
                        const float* d_uniform, uint64_t seed, uint32_t stream_id, const int64_t* d_step_counter,
                        float* d_onehot, int32_t* d_cls, uint8_t* d_traj_cls /*[T+1,rows] or NULL*/,
                        const uint64_t* d_graph_seed /*[G] or NULL*/, const int64_t* d_graph_row0 /*[G] or NULL*/, void* stream);
/* Random-stream addressing of the Philox draws (all entry points below): with d_graph_seed == NULL the counter is the batch
 * row under `seed`; with d_graph_seed [G] (one 64-bit seed per molecule) and d_graph_row0 [G] (first row of each graph)
 * the counter is the row's index inside its own molecule under that molecule's seed, so a molecule draws the same numbers
 * in any batch and on any rank (molecule-sharded sampling reproduces a single-GPU run bit for bit). */
/* x_prev = coef_x0[t] x_recon + coef_xt[t] x_t - grad (+ std[t] z unless t == 0); d_normal NULL -> Philox. */
int pg_position_step(int rows, const float* d_x_t, const float* d_x_recon, const float* d_energy_grad,
                     const float* d_coef_x0, const float* d_coef_xt, const float* d_std,
                     const int64_t* d_time_step, const int32_t* d_row_graph, const float* d_normal, uint64_t seed,
                     uint32_t stream_id, const int64_t* d_step_counter, float* d_x_prev,
                     float* d_traj_pos /*[T+1,rows,3] or NULL*/, const float* d_center /*[3], [G,3] or NULL*/,
                     int center_per_graph /*1: d_center holds one centre per graph (multi-pharmacophore batches)*/,
                     const uint64_t* d_graph_seed, const int64_t* d_graph_row0, void* stream);
/* T4: initial state of the reverse trajectory (reference models/transition.py:65-69,331-339 `sample_init`;
 * models/diffusion.py:406-408).  Categorical: class = argmax(gumbel(u) + d_log_prior[K]) per row, one-hot f32, int32 class
 * and log one-hot (log(clamp(onehot, 1e-30))) written; d_uniform [rows,K] supplies the draws, NULL -> Philox.
 * Positions: d_pos = z - centre, z standard normal (d_normal [rows,3] or Philox). */
int pg_sample_init(int rows, int K, const float* d_log_prior, const int32_t* d_row_graph, const float* d_uniform, uint64_t seed,
                   uint32_t stream_id, float* d_onehot, int32_t* d_cls, float* d_log_vt, const uint64_t* d_graph_seed,
                   const int64_t* d_graph_row0, void* stream);
int pg_position_init(int rows, const int32_t* d_row_graph, const float* d_normal, uint64_t seed, uint32_t stream_id,
                     const float* d_center, int center_per_graph, float* d_pos, const uint64_t* d_graph_seed,
                     const int64_t* d_graph_row0, void* stream);
/* T5: closed-form gradient of the guidance energies (utils/sample_utils.py:135-165; diffusion.py:476-502).
 * flags bit0 = atom_prox(min_d,max_d), bit1 = center_prox(d_phore_center), bit2 = add onto d_grad instead of overwriting
 * it (the reference sums one gradient per pos_guidance_opt entry: diffusion.py:479-501), bit3 = d_phore_center is [G,3]
 * (one pharmacophore per graph) instead of [3].  d_edge_cls: sampled classes, reference edge order.  Output d_grad [Nl,3]. */
int pg_guidance_grad(const PgPlan* p, const float* d_pos, const int32_t* d_edge_cls, int flags, float min_d,
                     float max_d, const float* d_phore_center, float* d_grad,
                     int norm_graphs /*the energies are means over the call's n_graphs (sample_utils.py:155,165); 0 = this
                                       plan's G like the reference, > 0 = a job-wide constant so that the drift of a molecule
                                       does not depend on how a sharded job is batched*/,
                     void* stream);

/* O2 / D2: atom-count heads (reference models/diffusion.py:148-163 `predict_atom_count`, and the interval of
 * `sample_nodes` :374-380).  d_h_phore_emb [P,128] is pg_phore_encode's output, d_h_phore [P,18] the raw features
 * (column ex_col == 1 marks exclusion spheres: 12 for zinc_300 / pdbbind, else 10).  Outputs per graph: count_l, count_u
 * [G] f32 and, if d_lo / d_hi are given, round(count * (max_atom - min_atom) + min_atom) as int32 (round half to even,
 * like torch.round). */
int pg_atom_count(const PgModel* m, PgPlan* p, const float* d_h_phore_emb, const float* d_h_phore, int ex_col,
                  float min_atom, float max_atom, float* d_count_l, float* d_count_u, int32_t* d_lo, int32_t* d_hi,
                  void* stream);

/* ---------------------------------------------------------------- M1 building block: K = 128 contraction
 * C[M, 128*ntiles128] = pro(A)[M,128] @ W + bias (+ resid): the Linear layers of models/common.py:99-119 after the
 * factorisation of DESIGN.md §3.  prologue 0: A ; 1: A + A2 ; 2: ReLU(LayerNorm(A + A2[gather])) (A2 / gather optional).
 * impl 0 = tcgen05 bf16x3 kernel (d_w_bf16_tiles: pre-swizzled bf16 hi/lo images, weights.bf16_tiles64),
 * impl 1 = fp32 FFMA reference kernel (d_wt: fp32 [128][128*ntiles128], k-major). */
int pg_gemm_k128(int impl, int prologue, int64_t M, const float* d_a, int64_t lda, const float* d_a2, int64_t lda2,
                 const int32_t* d_gather, const float* d_ln_g, const float* d_ln_b, const float* d_wt,
                 const float* d_w_bf16_tiles, const float* d_bias, const float* d_resid, int64_t ldr, float* d_c, int64_t ldc,
                 int ntiles128, void* stream);

/* plan accessors used by the host mirror */
const int32_t* pg_plan_ligand_graph(const PgPlan* p);   /* device [Nl]  atom -> graph */
const int32_t* pg_plan_edge_graph(const PgPlan* p);     /* device [E_b] reference-order edge -> graph */
int64_t pg_plan_kernel_launches(const PgPlan* p);       /* kernels launched through this plan so far */
/* Per-kernel-class device timing with CUDA events on the launching stream (eager launches only, not under
 * stream capture).  Classes: 0 dense GEMM, 1 kNN attention, 2 bond attention, 3 triplet, 4 kNN graph build, 5 other. */
int pg_plan_timing_enable(PgPlan* p, int on);
int pg_plan_timing_read(PgPlan* p, int kernel_class, double* ms_total, int64_t* launches);   /* synchronous */

#ifdef __cplusplus
}
#endif
#endif

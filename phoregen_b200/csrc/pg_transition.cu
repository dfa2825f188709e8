// Categorical posterior + Gumbel sampling, Gaussian position posterior and guidance gradient
// (SURVEY.md §8(a) rows T1, T2, T3, T5).  HBM-bound elementwise kernels, one thread per row.
#include "pg_plan.h"

namespace {
// Random streams.  Legacy addressing: Philox counter (batch row, step, stream, block) under the call's seed - the draws of
// a molecule then depend on where it sits in the batch.  Per-molecule addressing (graph_seed != NULL): the key is the
// molecule's own 64-bit seed and the counter its LOCAL row (row - first row of its graph), so a molecule draws the same
// numbers in any batch, on any rank (sharded sampling reproduces a single-GPU run bit for bit).
struct RngAddr { uint32_t row; uint2 key; };
__device__ __forceinline__ RngAddr rng_addr(int r, int g, uint64_t seed, const uint64_t* __restrict__ graph_seed,
                                            const int64_t* __restrict__ graph_row0) {
    RngAddr a;
    if (graph_seed) { const uint64_t s = graph_seed[g]; a.key = make_uint2((uint32_t)s, (uint32_t)(s >> 32)); a.row = (uint32_t)(r - (int)graph_row0[g]); }
    else { a.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); a.row = (uint32_t)r; }
    return a;
}

// T1 + T2 (transition.py:285-315, common.py:425-431, diffusion.py:453-466):
//   log_v0 = log_softmax(pred); out = clamp(log(exp(log_vt) Q_t^T + eps)) + clamp(log(exp(log_v0) Qbar_{t-1} + eps));
//   out -= logsumexp(out); t == 0 -> log_v0.  The SOFT posterior is written back as the next log_vt (reference quirk 1).
//   class = argmax(-log(-log(u + 1e-30) + 1e-30) + out), first maximum wins.
template <int K>
__global__ void __launch_bounds__(256) categorical_step_kernel(int rows, const float* __restrict__ pred, float* __restrict__ log_vt,
                                                               const float* __restrict__ q_mats, const float* __restrict__ tq,
                                                               const int64_t* __restrict__ tstep, const int* __restrict__ row_graph,
                                                               const float* __restrict__ uniform, uint64_t seed, uint32_t stream_id,
                                                               const int64_t* __restrict__ step_counter, float* __restrict__ onehot,
                                                               int* __restrict__ cls, uint8_t* __restrict__ traj,
                                                               const uint64_t* __restrict__ graph_seed, const int64_t* __restrict__ graph_row0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int gr = row_graph[r];
    const int t = (int)tstep[gr];
    float x[K], lv0[K], pvt[K], out[K];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = pred[(size_t)r * K + k]; m = fmaxf(m, x[k]); }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < K; k++) se += expf(x[k] - m);
    const float lse = m + logf(se);
#pragma unroll
    for (int k = 0; k < K; k++) { lv0[k] = x[k] - lse; pvt[k] = expf(log_vt[(size_t)r * K + k]); }
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) out[k] = lv0[k];
    } else {
        const float* T1 = tq + (size_t)t * K * K;               // transpose of the one-step matrix at t
        const float* Q2 = q_mats + (size_t)(t - 1) * K * K;     // cumulative matrix at t-1
        float p0[K];
#pragma unroll
        for (int k = 0; k < K; k++) p0[k] = expf(lv0[k]);
        float mm = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; k++) {
            float f1 = 0.f, f2 = 0.f;
#pragma unroll
            for (int j = 0; j < K; j++) {
                f1 = fmaf(pvt[j], __ldg(T1 + j * K + k), f1);
                f2 = fmaf(p0[j], __ldg(Q2 + j * K + k), f2);
            }
            out[k] = fmaxf(logf(f1 + 1e-30f), -32.f) + fmaxf(logf(f2 + 1e-30f), -32.f);
            mm = fmaxf(mm, out[k]);
        }
        float s2 = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) s2 += expf(out[k] - mm);
        const float l2 = mm + logf(s2);
#pragma unroll
        for (int k = 0; k < K; k++) out[k] -= l2;
    }
    // Gumbel arg-max
    float u[K];
    if (uniform) {
#pragma unroll
        for (int k = 0; k < K; k++) u[k] = uniform[(size_t)r * K + k];
    } else {
        const uint32_t step = (uint32_t)(*step_counter);
        const RngAddr ra = rng_addr(r, gr, seed, graph_seed, graph_row0);
#pragma unroll
        for (int b = 0; b < (K + 3) / 4; b++) {
            const uint4 rnd = philox4x32_10(make_uint4(ra.row, step, stream_id, (uint32_t)b), ra.key);
            const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (b * 4 + i < K) u[b * 4 + i] = u32_to_unit(w[i]);
        }
    }
    int best = 0;
    float bv = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float gmb = -logf(-logf(u[k] + 1e-30f) + 1e-30f) + out[k];
        if (gmb > bv) { bv = gmb; best = k; }
        log_vt[(size_t)r * K + k] = out[k];
    }
    if (cls) cls[r] = best;
    if (traj) traj[(size_t)(*step_counter + 1) * rows + r] = (uint8_t)best;   // trajectory slot i+1 (diffusion.py:510-512)
    if (onehot) {
#pragma unroll
        for (int k = 0; k < K; k++) onehot[(size_t)r * K + k] = (k == best) ? 1.f : 0.f;
    }
}

// T3 (transition.py:44-63)
__global__ void __launch_bounds__(256) position_step_kernel(int rows, const float* __restrict__ x_t, const float* __restrict__ x_recon,
                                                            const float* __restrict__ grad, const float* __restrict__ c0,
                                                            const float* __restrict__ ct, const float* __restrict__ sd,
                                                            const int64_t* __restrict__ tstep, const int* __restrict__ row_graph,
                                                            const float* __restrict__ normal, uint64_t seed, uint32_t stream_id,
                                                            const int64_t* __restrict__ step_counter, float* __restrict__ x_prev,
                                                            float* __restrict__ traj, const float* __restrict__ center,
                                                            int center_per_graph, const uint64_t* __restrict__ graph_seed,
                                                            const int64_t* __restrict__ graph_row0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int gr = row_graph[r];
    const int t = (int)tstep[gr];
    if (center && center_per_graph) center += (size_t)gr * 3;
    const float a = c0[t], b = ct[t], sg = sd[t];
    float z[3] = {0.f, 0.f, 0.f};
    if (t != 0) {
        if (normal) {
            z[0] = normal[(size_t)r * 3]; z[1] = normal[(size_t)r * 3 + 1]; z[2] = normal[(size_t)r * 3 + 2];
        } else {   // Box-Muller on Philox draws
            const RngAddr adr = rng_addr(r, gr, seed, graph_seed, graph_row0);
            const uint4 rnd = philox4x32_10(make_uint4(adr.row, (uint32_t)(*step_counter), stream_id, 0u), adr.key);
            const float u1 = 1.0f - u32_to_unit(rnd.x), u2 = u32_to_unit(rnd.y), u3 = 1.0f - u32_to_unit(rnd.z), u4 = u32_to_unit(rnd.w);
            const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
            z[0] = ra * cospif(2.0f * u2); z[1] = ra * sinpif(2.0f * u2); z[2] = rb * cospif(2.0f * u4);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float mu = a * x_recon[(size_t)r * 3 + k] + b * x_t[(size_t)r * 3 + k];
        if (grad) mu -= grad[(size_t)r * 3 + k];
        const float xp = (t == 0) ? mu : mu + sg * z[k];
        x_prev[(size_t)r * 3 + k] = xp;
        if (traj) traj[((size_t)(*step_counter + 1) * rows + r) * 3 + k] = xp + (center ? center[k] : 0.f);
    }
}

// T4 (transition.py:65-69,331-339; diffusion.py:406-408): initial state.  Categorical rows: class = arg-max(gumbel(u) +
// log prior), one-hot, log one-hot (index_to_log_onehot, common.py:398-402: log(clamp(onehot, 1e-30))).  Positions:
// standard normal minus the centre.  Draws: supplied, or Philox with the addressing above and step counter 0xFFFFFFFF
// (a slot no reverse step uses).
template <int K>
__global__ void __launch_bounds__(256) categorical_init_kernel(int rows, const float* __restrict__ log_prior, const int* __restrict__ row_graph,
                                                               const float* __restrict__ uniform, uint64_t seed, uint32_t stream_id,
                                                               float* __restrict__ onehot, int* __restrict__ cls, float* __restrict__ log_vt,
                                                               const uint64_t* __restrict__ graph_seed, const int64_t* __restrict__ graph_row0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float u[K];
    if (uniform) {
#pragma unroll
        for (int k = 0; k < K; k++) u[k] = uniform[(size_t)r * K + k];
    } else {
        const RngAddr ra = rng_addr(r, row_graph[r], seed, graph_seed, graph_row0);
#pragma unroll
        for (int b = 0; b < (K + 3) / 4; b++) {
            const uint4 rnd = philox4x32_10(make_uint4(ra.row, 0xFFFFFFFFu, stream_id, (uint32_t)b), ra.key);
            const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (b * 4 + i < K) u[b * 4 + i] = u32_to_unit(w[i]);
        }
    }
    int best = 0;
    float bv = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float gmb = -logf(-logf(u[k] + 1e-30f) + 1e-30f) + log_prior[k];
        if (gmb > bv) { bv = gmb; best = k; }
    }
    cls[r] = best;
#pragma unroll
    for (int k = 0; k < K; k++) {
        onehot[(size_t)r * K + k] = (k == best) ? 1.f : 0.f;
        log_vt[(size_t)r * K + k] = (k == best) ? 0.f : -69.07755279f;       // logf(1e-30f)
    }
}
__global__ void __launch_bounds__(256) position_init_kernel(int rows, const int* __restrict__ row_graph, const float* __restrict__ normal,
                                                            uint64_t seed, uint32_t stream_id, const float* __restrict__ center,
                                                            int center_per_graph, float* __restrict__ pos,
                                                            const uint64_t* __restrict__ graph_seed, const int64_t* __restrict__ graph_row0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int gr = row_graph[r];
    float z[3];
    if (normal) {
        z[0] = normal[(size_t)r * 3]; z[1] = normal[(size_t)r * 3 + 1]; z[2] = normal[(size_t)r * 3 + 2];
    } else {
        const RngAddr ra = rng_addr(r, gr, seed, graph_seed, graph_row0);
        const uint4 rnd = philox4x32_10(make_uint4(ra.row, 0xFFFFFFFFu, stream_id, 0u), ra.key);
        const float u1 = 1.0f - u32_to_unit(rnd.x), u2 = u32_to_unit(rnd.y), u3 = 1.0f - u32_to_unit(rnd.z), u4 = u32_to_unit(rnd.w);
        const float ra_ = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
        z[0] = ra_ * cospif(2.0f * u2); z[1] = ra_ * sinpif(2.0f * u2); z[2] = rb * cospif(2.0f * u4);
    }
    if (center && center_per_graph) center += (size_t)gr * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) pos[(size_t)r * 3 + k] = z[k] - (center ? center[k] : 0.f);
}

// T5: closed-form gradient of the guidance energies.  One CTA per graph, deterministic (fixed reduction order).
//  atom_prox  (sample_utils.py:135-157): E = mean_g mean_{edges with class>0} relu(len-max_d) + relu(min_d-len)
//  center_prox(sample_utils.py:160-165): E = mean_g | centroid_g - phore_center |
__global__ void __launch_bounds__(128) guidance_kernel(PlanDev d, const int* __restrict__ inv_perm, const float* __restrict__ pos,
                                                       const int* __restrict__ edge_cls, int flags, float min_d, float max_d,
                                                       const float* __restrict__ center, float* __restrict__ grad, int norm_graphs) {
    const int g = blockIdx.x, tid = threadIdx.x;
    const int n = d.g_n[g], a0 = d.lig_off[g];
    const long long e0 = d.eoff[g];
    __shared__ float red[128];
    __shared__ float cen[3];
    if (!(flags & 4)) {                           // bit 2: accumulate onto the gradient of a previous drift entry
        for (int i = tid; i < n * 3; i += blockDim.x) grad[(size_t)a0 * 3 + i] = 0.f;
    }
    __syncthreads();
    const float invG = 1.0f / (float)(norm_graphs > 0 ? norm_graphs : d.G);
    if (flags & 8) center += (size_t)g * 3;       // one pharmacophore centre per graph
    if (flags & 1) {
        // number of bonded directed edges of this graph
        float c = 0.f;
        for (int r = tid; r < n * (n - 1); r += blockDim.x) c += edge_cls[inv_perm[e0 + r]] > 0 ? 1.f : 0.f;
        red[tid] = c;
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
        const float cnt = red[0];
        __syncthreads();
        if (cnt > 0.f) {
            // atom i gathers over all its incident directed edges (as dst in the internal order, and as src by symmetry)
            for (int i = tid; i < n; i += blockDim.x) {
                float g0 = 0.f, g1 = 0.f, g2 = 0.f;
                const float* pi = pos + (size_t)(a0 + i) * 3;
                for (int t = 0; t < n - 1; t++) {
                    const int j = t + (t >= i);
                    const float* pj = pos + (size_t)(a0 + j) * 3;
                    const float d0 = pi[0] - pj[0], d1 = pi[1] - pj[1], d2 = pi[2] - pj[2];
                    const float len = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
                    const float sgn = (len > max_d ? 1.f : 0.f) - (len < min_d ? 1.f : 0.f);
                    // edge j->i (slot in dst-major order) and edge i->j
                    const int c_in = edge_cls[inv_perm[e0 + (long long)i * (n - 1) + t]];
                    const int c_out = edge_cls[inv_perm[e0 + (long long)j * (n - 1) + (i - (i > j))]];
                    const float wgt = ((c_in > 0 ? 1.f : 0.f) + (c_out > 0 ? 1.f : 0.f)) * sgn / fmaxf(len, 1e-30f);
                    g0 = fmaf(wgt, d0, g0); g1 = fmaf(wgt, d1, g1); g2 = fmaf(wgt, d2, g2);
                }
                const float sc = invG / cnt;
                grad[(size_t)(a0 + i) * 3] += g0 * sc; grad[(size_t)(a0 + i) * 3 + 1] += g1 * sc; grad[(size_t)(a0 + i) * 3 + 2] += g2 * sc;
            }
        }
        __syncthreads();
    }
    if (flags & 2) {
        if (tid < 3) {
            float s = 0.f;
            for (int i = 0; i < n; i++) s += pos[(size_t)(a0 + i) * 3 + tid];
            cen[tid] = s / (float)n - center[tid];
        }
        __syncthreads();
        const float nr = sqrtf(cen[0] * cen[0] + cen[1] * cen[1] + cen[2] * cen[2]);
        const float sc = invG / ((float)n * fmaxf(nr, 1e-30f));
        for (int i = tid; i < n * 3; i += blockDim.x) grad[(size_t)a0 * 3 + i] += cen[i % 3] * sc;
    }
}
}  // namespace

extern "C" int pg_categorical_step(int rows, int K, const float* d_pred, float* d_log_vt, const float* d_q_mats,
                                   const float* d_tq_onestep, const int64_t* d_time_step, const int32_t* d_row_graph,
                                   const float* d_uniform, uint64_t seed, uint32_t stream_id, const int64_t* d_step_counter,
                                   float* d_onehot, int32_t* d_cls, uint8_t* d_traj, const uint64_t* d_graph_seed,
                                   const int64_t* d_graph_row0, void* stream) {
    if (rows <= 0) return PG_OK;
    if (d_traj && !d_step_counter) { pg_set_error("pg_categorical_step: trajectory output needs the step counter"); return PG_EINVAL; }
    if (!d_uniform && !d_step_counter) { pg_set_error("pg_categorical_step: need uniforms or a step counter"); return PG_EINVAL; }
    const unsigned grid = (unsigned)((rows + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (K == PG_NODE_CLASSES)
        categorical_step_kernel<PG_NODE_CLASSES><<<grid, 256, 0, s>>>(rows, d_pred, d_log_vt, d_q_mats, d_tq_onestep, d_time_step,
                                                                      d_row_graph, d_uniform, seed, stream_id, d_step_counter, d_onehot, d_cls, d_traj,
                                                                      d_graph_seed, d_graph_row0);
    else if (K == PG_EDGE_CLASSES)
        categorical_step_kernel<PG_EDGE_CLASSES><<<grid, 256, 0, s>>>(rows, d_pred, d_log_vt, d_q_mats, d_tq_onestep, d_time_step,
                                                                      d_row_graph, d_uniform, seed, stream_id, d_step_counter, d_onehot, d_cls, d_traj,
                                                                      d_graph_seed, d_graph_row0);
    else { pg_set_error("pg_categorical_step: K must be 12 or 6"); return PG_EINVAL; }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_position_step(int rows, const float* d_x_t, const float* d_x_recon, const float* d_energy_grad,
                                const float* d_coef_x0, const float* d_coef_xt, const float* d_std, const int64_t* d_time_step,
                                const int32_t* d_row_graph, const float* d_normal, uint64_t seed, uint32_t stream_id,
                                const int64_t* d_step_counter, float* d_x_prev, float* d_traj, const float* d_center,
                                int center_per_graph, const uint64_t* d_graph_seed, const int64_t* d_graph_row0, void* stream) {
    if (rows <= 0) return PG_OK;
    if (d_traj && !d_step_counter) { pg_set_error("pg_position_step: trajectory output needs the step counter"); return PG_EINVAL; }
    if (!d_normal && !d_step_counter) { pg_set_error("pg_position_step: need normals or a step counter"); return PG_EINVAL; }
    position_step_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, d_x_t, d_x_recon, d_energy_grad, d_coef_x0, d_coef_xt, d_std, d_time_step, d_row_graph, d_normal, seed, stream_id,
        d_step_counter, d_x_prev, d_traj, d_center, center_per_graph, d_graph_seed, d_graph_row0);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_sample_init(int rows, int K, const float* d_log_prior, const int32_t* d_row_graph, const float* d_uniform,
                              uint64_t seed, uint32_t stream_id, float* d_onehot, int32_t* d_cls, float* d_log_vt,
                              const uint64_t* d_graph_seed, const int64_t* d_graph_row0, void* stream) {
    if (rows <= 0) return PG_OK;
    if (!d_log_prior || !d_row_graph || !d_onehot || !d_cls || !d_log_vt || (!d_graph_seed) != (!d_graph_row0)) { pg_set_error("pg_sample_init: bad argument"); return PG_EINVAL; }
    const unsigned grid = (unsigned)((rows + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (K == PG_NODE_CLASSES)
        categorical_init_kernel<PG_NODE_CLASSES><<<grid, 256, 0, s>>>(rows, d_log_prior, d_row_graph, d_uniform, seed, stream_id, d_onehot, d_cls, d_log_vt, d_graph_seed, d_graph_row0);
    else if (K == PG_EDGE_CLASSES)
        categorical_init_kernel<PG_EDGE_CLASSES><<<grid, 256, 0, s>>>(rows, d_log_prior, d_row_graph, d_uniform, seed, stream_id, d_onehot, d_cls, d_log_vt, d_graph_seed, d_graph_row0);
    else { pg_set_error("pg_sample_init: K must be 12 or 6"); return PG_EINVAL; }
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_position_init(int rows, const int32_t* d_row_graph, const float* d_normal, uint64_t seed, uint32_t stream_id,
                                const float* d_center, int center_per_graph, float* d_pos, const uint64_t* d_graph_seed,
                                const int64_t* d_graph_row0, void* stream) {
    if (rows <= 0) return PG_OK;
    if (!d_row_graph || !d_pos || (!d_graph_seed) != (!d_graph_row0)) { pg_set_error("pg_position_init: bad argument"); return PG_EINVAL; }
    position_init_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, d_row_graph, d_normal, seed, stream_id, d_center,
                                                                                          center_per_graph, d_pos, d_graph_seed, d_graph_row0);
    PG_LAUNCH_CHECK();
    return PG_OK;
}

extern "C" int pg_guidance_grad(const PgPlan* p, const float* d_pos, const int32_t* d_edge_cls, int flags, float min_d,
                                float max_d, const float* d_phore_center, float* d_grad, int norm_graphs, void* stream) {
    if ((flags & 1) && !d_edge_cls) { pg_set_error("pg_guidance_grad: atom_prox needs edge classes"); return PG_EINVAL; }
    if ((flags & 2) && !d_phore_center) { pg_set_error("pg_guidance_grad: center_prox needs the phore centre"); return PG_EINVAL; }
    guidance_kernel<<<(unsigned)p->d.G, 128, 0, (cudaStream_t)stream>>>(p->d, p->inv_perm, d_pos, d_edge_cls, flags, min_d, max_d,
                                                                        d_phore_center, d_grad, norm_graphs);
    PG_LAUNCH_CHECK();
    const_cast<PgPlan*>(p)->launches++;
    return PG_OK;
}

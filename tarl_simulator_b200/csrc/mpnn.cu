// mpnn.cu — learned-MPNN path, HBM-bound parts (sm_100a): policy embedding gather fwd/bwd and the fused
// GraphDistribution (segmented softmax / log-prob / entropy / mode / inverse-CDF sample) fwd/bwd.
//
// Reference semantics: /root/reference/src/agents/mpnn_agent.py:117-217 (MPNNPolicyNet active path) and
// /root/reference/src/reinforcement_learning.py:15-96 (GraphDistribution), with the declared divergences D1/D2/D3/D7
// of SURVEY.md §8c (see oracle/mpnn_port.py). fp32 throughout; every reduction has a fixed order (no float atomics
// on the common path), so results are run-to-run deterministic.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr float kLogEps = 1e-8f;  // src/reinforcement_learning.py:27

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

// ------------------------------------------------------------------------------------------------ policy embedding
// node pass: idx[b,n] = ROAD_INDEX >= 0 ? ROAD_INDEX : n (D2);  emb[b,n] = W[idx]
__global__ void __launch_bounds__(kThreads) k_policy_node(const float* __restrict__ w, int rows,
                                                          const float* __restrict__ nf, int64_t nf_bs, int64_t nf_rs,
                                                          int ridx_col, int B, int N, float* __restrict__ emb,
                                                          int32_t* __restrict__ idx, int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * N) return;
    const int b = (int)(i / N), n = (int)(i % N);
    const float r = nf[b * nf_bs + n * nf_rs + ridx_col];
    long long k = (long long)r;  // .to(torch.long): truncation
    if (k < 0) k = n;
    if (k >= rows) {  // nn.Embedding raises IndexError
        atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_EMBED_RANGE);
        k = 0;
    }
    idx[i] = (int32_t)k;
    emb[i] = w[k];
}

// edge pass: logits[b,e] = emb[b, dst[e]]
__global__ void __launch_bounds__(kThreads) k_policy_edge(const float* __restrict__ emb, const int32_t* __restrict__ dst,
                                                          int B, int N, int E, float* __restrict__ logits) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int n = dst[e];
    for (int b = 0; b < B; ++b) logits[(int64_t)b * E + e] = emb[(int64_t)b * N + n];
}

// backward node pass: G[b,n] = sum over in-edges (ascending edge id) of grad_logits[b,e]
__global__ void __launch_bounds__(kThreads) k_policy_node_grad(tarl_csr in, const float* __restrict__ gl, int B, int E,
                                                               float* __restrict__ G) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= in.n_rows) return;
    const int k0 = in.ptr[n], k1 = in.ptr[n + 1];
    for (int b = 0; b < B; ++b) {
        float acc = 0.0f;
        for (int k = k0; k < k1; ++k) acc += gl[(int64_t)b * E + in.eid[k]];
        G[(int64_t)b * in.n_rows + n] = acc;
    }
}

// backward weight pass: gradW[idx[b,n]] += G[b,n]. Batch rows that share the index of row 0 (always, in practice: the
// ROAD_INDEX column is static) are summed in registers in batch order; with an injective index map that makes every
// weight row the target of exactly one atomicAdd, i.e. deterministic.
__global__ void __launch_bounds__(kThreads) k_policy_weight_grad(const float* __restrict__ G, const int32_t* __restrict__ idx,
                                                                 int B, int N, float* __restrict__ gw) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int k0 = idx[n];
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
        const int k = idx[(int64_t)b * N + n];
        const float g = G[(int64_t)b * N + n];
        if (k == k0) acc += g;
        else atomicAdd(&gw[k], g);
    }
    atomicAdd(&gw[k0], acc);
}

// ------------------------------------------------------------------------------------------------ GraphDistribution
__device__ __forceinline__ float load_action(const void* a, int dtype, int64_t i) {
    switch (dtype) {
        case TARL_ACTION_U8: return (float)static_cast<const uint8_t*>(a)[i];
        case TARL_ACTION_I64: return (float)static_cast<const long long*>(a)[i];
        default: return static_cast<const float*>(a)[i];
    }
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    T r = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : T(0);
    if (wid == 0) {
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread 0
}

// One thread per (batch row, source group). Softmax over the group's edges in ascending edge id, then whatever of
// {proba, entropy, log_prob, mode} was asked for. Per-block partial sums keep the [B] reductions deterministic.
__global__ void __launch_bounds__(kThreads) k_gd_forward(tarl_csr grp, const float* __restrict__ logits, float temp,
                                                         int E, const void* __restrict__ action, int action_dtype,
                                                         float* __restrict__ proba, float* __restrict__ mode,
                                                         float* __restrict__ part_ent, float* __restrict__ part_lp,
                                                         int32_t* __restrict__ part_bad) {
    __shared__ float sm_f[kThreads / 32];
    __shared__ int sm_i[kThreads / 32];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const float* lg = logits + (int64_t)b * E;
    float ent = 0.0f, lp = 0.0f;
    int bad = 0;
    if (g < grp.n_rows) {
        const int k0 = grp.ptr[g], k1 = grp.ptr[g + 1];
        float mx = -FLT_MAX;
        for (int k = k0; k < k1; ++k) mx = fmaxf(mx, lg[grp.eid[k]] / temp);
        float den = 0.0f;
        for (int k = k0; k < k1; ++k) den += expf(lg[grp.eid[k]] / temp - mx);
        float asum = 0.0f, best = -FLT_MAX;
        int best_e = -1;
        for (int k = k0; k < k1; ++k) {
            const int e = grp.eid[k];
            const float p = expf(lg[e] / temp - mx) / den;
            const float l = logf(p + kLogEps);
            ent -= p * l;
            if (action != nullptr) {
                const float a = load_action(action, action_dtype, (int64_t)b * E + e);
                lp += a * l;
                asum += a;
            }
            if (proba != nullptr) proba[(int64_t)b * E + e] = p;
            if (p > best) { best = p; best_e = e; }
        }
        if (action != nullptr && asum != 1.0f) bad = 1;   // not exactly one selected edge in this group (:86-88)
        if (mode != nullptr && best_e >= 0) mode[(int64_t)b * E + best_e] = 1.0f;
    }
    const int nb = gridDim.x;
    if (part_ent != nullptr) {
        const float s = block_sum(ent, sm_f);
        if (threadIdx.x == 0) part_ent[(int64_t)b * nb + blockIdx.x] = s;
    }
    if (part_lp != nullptr) {
        const float s = block_sum(lp, sm_f);
        const int c = block_sum(bad, sm_i);
        if (threadIdx.x == 0) {
            part_lp[(int64_t)b * nb + blockIdx.x] = s;
            part_bad[(int64_t)b * nb + blockIdx.x] = c;
        }
    }
}

// second stage: one block per batch row sums the per-block partials in a fixed order
__global__ void __launch_bounds__(kThreads) k_gd_finish(const float* __restrict__ part_ent, const float* __restrict__ part_lp,
                                                        const int32_t* __restrict__ part_bad, int nb,
                                                        float* __restrict__ entropy, float* __restrict__ log_prob) {
    __shared__ float sm_f[kThreads / 32];
    __shared__ int sm_i[kThreads / 32];
    const int b = blockIdx.x;
    if (part_ent != nullptr) {
        float a = 0.0f;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) a += part_ent[(int64_t)b * nb + i];
        a = block_sum(a, sm_f);
        if (threadIdx.x == 0) entropy[b] = a;
    }
    if (part_lp != nullptr) {
        float a = 0.0f;
        int c = 0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) {
            a += part_lp[(int64_t)b * nb + i];
            c += part_bad[(int64_t)b * nb + i];
        }
        a = block_sum(a, sm_f);
        c = block_sum(c, sm_i);
        if (threadIdx.x == 0) log_prob[b] = (c == 0) ? a : -INFINITY;   // :89-91
    }
}

// grad_logits[b,e] = inv_t * ( g_lp[b] * (a_e r_e - p_e S_a)  -  g_ent[b] * p_e (c_e - S_c) ),
//   r_e = p_e/(p_e+eps), c_e = log(p_e+eps) + r_e, S_a = sum_e a_e r_e, S_c = sum_e p_e c_e  (sums over the group).
// Rows whose action was impossible carry log_prob = -inf assigned as a constant in the reference: no gradient.
__global__ void __launch_bounds__(kThreads) k_gd_backward(tarl_csr grp, const float* __restrict__ logits, float temp,
                                                          int E, const void* __restrict__ action, int action_dtype,
                                                          const float* __restrict__ g_lp, const float* __restrict__ g_ent,
                                                          const float* __restrict__ log_prob, float* __restrict__ grad) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (g >= grp.n_rows) return;
    const float* lg = logits + (int64_t)b * E;
    float wl = (g_lp != nullptr && action != nullptr) ? g_lp[b] : 0.0f;
    if (log_prob != nullptr && log_prob[b] == -INFINITY) wl = 0.0f;
    const float we = (g_ent != nullptr) ? g_ent[b] : 0.0f;
    const int k0 = grp.ptr[g], k1 = grp.ptr[g + 1];
    float mx = -FLT_MAX;
    for (int k = k0; k < k1; ++k) mx = fmaxf(mx, lg[grp.eid[k]] / temp);
    float den = 0.0f;
    for (int k = k0; k < k1; ++k) den += expf(lg[grp.eid[k]] / temp - mx);
    float Sa = 0.0f, Sc = 0.0f;
    for (int k = k0; k < k1; ++k) {
        const int e = grp.eid[k];
        const float p = expf(lg[e] / temp - mx) / den;
        const float r = p / (p + kLogEps);
        const float c = logf(p + kLogEps) + r;
        if (wl != 0.0f) Sa += load_action(action, action_dtype, (int64_t)b * E + e) * r;
        Sc += p * c;
    }
    for (int k = k0; k < k1; ++k) {
        const int e = grp.eid[k];
        const float p = expf(lg[e] / temp - mx) / den;
        const float r = p / (p + kLogEps);
        const float c = logf(p + kLogEps) + r;
        float v = -we * p * (c - Sc);
        if (wl != 0.0f) v += wl * (load_action(action, action_dtype, (int64_t)b * E + e) * r - p * Sa);
        grad[(int64_t)b * E + e] = v / temp;
    }
}

// inverse-CDF sample, one uniform per (row, group): first edge (ascending edge id, D3) with u < cumulative proba (:62-80)
__global__ void __launch_bounds__(kThreads) k_gd_sample(tarl_csr grp, const float* __restrict__ logits, float temp, int E,
                                                        const float* __restrict__ u, long long* __restrict__ onehot) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (g >= grp.n_rows) return;
    const float* lg = logits + (int64_t)b * E;
    const int k0 = grp.ptr[g], k1 = grp.ptr[g + 1];
    float mx = -FLT_MAX;
    for (int k = k0; k < k1; ++k) mx = fmaxf(mx, lg[grp.eid[k]] / temp);
    float den = 0.0f;
    for (int k = k0; k < k1; ++k) den += expf(lg[grp.eid[k]] / temp - mx);
    const float ug = u[(int64_t)b * grp.n_rows + g];
    float cum = 0.0f;
    for (int k = k0; k < k1; ++k) {
        const int e = grp.eid[k];
        cum += expf(lg[e] / temp - mx) / den;
        if (ug < cum) { onehot[(int64_t)b * E + e] = 1; break; }
    }
}

int check_csr(const tarl_csr* c) {
    if (c == nullptr || c->n_rows < 0 || c->n_edges < 0) return TARL_E_BADARG;
    if (c->n_rows > 0 && c->ptr == nullptr) return TARL_E_BADARG;
    if (c->n_edges > 0 && c->eid == nullptr) return TARL_E_BADARG;
    return TARL_OK;
}

}  // namespace

extern "C" {

int tarl_policy_embed_forward(const float* emb_weight, int32_t emb_rows, const float* node_features,
                              int64_t nf_batch_stride, int64_t nf_row_stride, int32_t road_index_col, int32_t batch,
                              int32_t n_nodes, const int32_t* edge_dst, int32_t n_edges, float* node_emb,
                              int32_t* node_idx, float* logits, int32_t* flags, void* stream) {
    if (batch < 0 || n_nodes < 0 || n_edges < 0 || emb_rows <= 0 || flags == nullptr) return TARL_E_BADARG;
    if (batch == 0 || n_nodes == 0) return TARL_OK;
    if (!emb_weight || !node_features || !node_emb || !node_idx || (n_edges > 0 && (!edge_dst || !logits)))
        return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    k_policy_node<<<blocks_for((int64_t)batch * n_nodes), kThreads, 0, s>>>(
        emb_weight, emb_rows, node_features, nf_batch_stride, nf_row_stride, road_index_col, batch, n_nodes, node_emb,
        node_idx, flags);
    if (n_edges > 0)
        k_policy_edge<<<blocks_for(n_edges), kThreads, 0, s>>>(node_emb, edge_dst, batch, n_nodes, n_edges, logits);
    return launch_status();
}

int tarl_policy_embed_backward(const tarl_csr* by_target, const float* grad_logits, const int32_t* node_idx,
                               int32_t batch, float* node_grad, float* grad_weight, int32_t emb_rows, void* stream) {
    int rc = check_csr(by_target);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || emb_rows <= 0 || !grad_weight) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)emb_rows, s) != cudaSuccess) return TARL_E_LAUNCH;
    if (batch == 0 || by_target->n_rows == 0) return TARL_OK;
    if (!node_idx || !node_grad || (by_target->n_edges > 0 && !grad_logits)) return TARL_E_BADARG;
    const int nb = blocks_for(by_target->n_rows);
    k_policy_node_grad<<<nb, kThreads, 0, s>>>(*by_target, grad_logits, batch, by_target->n_edges, node_grad);
    k_policy_weight_grad<<<nb, kThreads, 0, s>>>(node_grad, node_idx, batch, by_target->n_rows, grad_weight);
    return launch_status();
}

int32_t tarl_graphdist_partial_count(int32_t n_groups) { return n_groups > 0 ? blocks_for(n_groups) : 0; }

int tarl_graphdist_forward(const tarl_csr* groups, const float* logits, float temperature, int32_t batch,
                           const void* action, int32_t action_dtype, float* proba, float* mode, float* entropy,
                           float* log_prob, float* partials, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (action != nullptr && (action_dtype < 0 || action_dtype > 2))) return TARL_E_BADARG;
    if (log_prob != nullptr && action == nullptr) return TARL_E_BADARG;
    if (batch == 0) return TARL_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int E = groups->n_edges, K = groups->n_rows;
    if (E > 0 && logits == nullptr) return TARL_E_BADARG;
    if (mode != nullptr && E > 0 &&
        cudaMemsetAsync(mode, 0, sizeof(float) * (size_t)batch * E, s) != cudaSuccess) return TARL_E_LAUNCH;
    const int nb = K > 0 ? blocks_for(K) : 0;
    float* part_ent = nullptr; float* part_lp = nullptr; int32_t* part_bad = nullptr;
    if (entropy != nullptr || log_prob != nullptr) {
        if (partials == nullptr && nb > 0) return TARL_E_WORKSPACE;
        part_ent = entropy ? partials : nullptr;
        part_lp = log_prob ? partials + (size_t)batch * nb : nullptr;
        part_bad = log_prob ? reinterpret_cast<int32_t*>(partials + 2 * (size_t)batch * nb) : nullptr;
    }
    if (nb > 0) {
        dim3 grid(nb, batch);
        k_gd_forward<<<grid, kThreads, 0, s>>>(*groups, logits, temperature, E, action, action_dtype, proba, mode,
                                               part_ent, part_lp, part_bad);
    }
    if (entropy != nullptr || log_prob != nullptr)
        k_gd_finish<<<batch, kThreads, 0, s>>>(part_ent, part_lp, part_bad, nb, entropy, log_prob);
    return launch_status();
}

int tarl_graphdist_backward(const tarl_csr* groups, const float* logits, float temperature, int32_t batch,
                            const void* action, int32_t action_dtype, const float* grad_log_prob,
                            const float* grad_entropy, const float* log_prob, float* grad_logits, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (action != nullptr && (action_dtype < 0 || action_dtype > 2))) return TARL_E_BADARG;
    if (batch == 0 || groups->n_edges == 0) return TARL_OK;
    if (!logits || !grad_logits) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // edges whose source has no group cannot exist (every edge has a source), so every grad entry is written
    dim3 grid(blocks_for(groups->n_rows), batch);
    k_gd_backward<<<grid, kThreads, 0, s>>>(*groups, logits, temperature, groups->n_edges, action, action_dtype,
                                            grad_log_prob, grad_entropy, log_prob, grad_logits);
    return launch_status();
}

int tarl_graphdist_sample(const tarl_csr* groups, const float* logits, float temperature, int32_t batch,
                          const float* uniforms, int64_t* onehot, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0) return TARL_E_BADARG;
    if (batch == 0 || groups->n_edges == 0) return TARL_OK;
    if (!logits || !uniforms || !onehot) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(onehot, 0, sizeof(int64_t) * (size_t)batch * groups->n_edges, s) != cudaSuccess)
        return TARL_E_LAUNCH;
    dim3 grid(blocks_for(groups->n_rows), batch);
    k_gd_sample<<<grid, kThreads, 0, s>>>(*groups, logits, temperature, groups->n_edges, uniforms,
                                          reinterpret_cast<long long*>(onehot));
    return launch_status();
}

}  // extern "C"

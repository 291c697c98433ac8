// core_step.cu — the per-timestep network step on the reference's row layout, in place (sm_100a).
//
// Semantics: SURVEY.md Appendix A == /root/reference/src/direction_mpnn.py:44-196 + src/response_mpnn.py:42-127.
// Compiled with -fmad=false: every fp32 add/sub/mul/div below must round exactly like the reference's ATen ops.
//
// Instead of the reference's four [E, 3*Nmax+7] row gathers, every link publishes two 16-byte summaries of its
// PRE-step row once (k_offer); the per-edge phases only gather those (they stay L2-resident: 32 B/link), and every
// write to x happens after all reads of pre-step state are done (the reference gets that from materialised gathers).
//
//   k_offer          per link      reads own row (head triplet, statics, old tail id)  -> recA, recB, dtt
//   k_select_append  per link d    scans in-edges in ascending original edge id: masks, prob sum, Gumbel arg-max
//                                  (strict >, lowest edge id wins), then the tail write on row d       -> post
//   k_respond_shift  per link u    OR over out-edges of "tail(d) == head(u)" on post-append summaries, writes
//                                  delta_tt for its out-edges, warp-cooperative FIFO shift of popping rows
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace {

constexpr int kThreads = 256;

// torch.maximum / torch.clamp(min=) propagate NaN; fmaxf does not.
__device__ __forceinline__ float max_propagate_nan(float a, float b) {
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

struct Workspace {
    float4* recA;  // {head_id, SEL, room = MAXN-NUM, flag bits}           pre-step, gathered per in-edge
    float4* recB;  // {ROAD_INDEX, NUM, new exit time t+tt, old tail id}   pre-step, own link only
    float4* post;  // {NUM, tail id, head id, 0}                           after the direction phase
    float* dtt;    // max(dep_head - arr_head - FFTT, 0) per link (identical for all of its out-edges)
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

inline Workspace carve(void* base, int32_t n) {
    char* p = static_cast<char*>(base);
    Workspace w;
    w.recA = reinterpret_cast<float4*>(p); p += align16(sizeof(float4) * (size_t)n);
    w.recB = reinterpret_cast<float4*>(p); p += align16(sizeof(float4) * (size_t)n);
    w.post = reinterpret_cast<float4*>(p); p += align16(sizeof(float4) * (size_t)n);
    w.dtt = reinterpret_cast<float*>(p);
    return w;
}

enum : int { kA1 = 1, kA2 = 2, kFree = 4, kBad = 8 };

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_offer(float* __restrict__ x, int64_t stride, int N, int Nmax,
                                                    const float* __restrict__ cc, const float* __restrict__ sel_in,
                                                    float t, Workspace w, int32_t* __restrict__ flags) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float* row = x + (int64_t)n * stride;
    const int c0 = 3 * Nmax;
    const float head_id = row[0], head_arr = row[Nmax], head_dep = row[2 * Nmax];
    const float maxn = row[c0], num = row[c0 + 1], fftt = row[c0 + 2], ridx = row[c0 + 6];
    float sel;
    if (sel_in != nullptr) {  // this step's routing decision, applied first (src/reinforcement_learning.py:231)
        sel = sel_in[n];
        row[c0 + 5] = sel;
    } else {
        sel = row[c0 + 5];
    }
    float ccn;
    if (cc != nullptr) {
        ccn = cc[n];
    } else {  // src/simulation_core_model.py:60-67
        const float crit = (row[c0 + 4] * fftt) / 3600.0f;
        ccn = fftt * ((maxn + 10.0f) - crit);
    }
    const bool bad = !(num >= 0.0f) || !(num < (float)Nmax);
    int q = bad ? 0 : (int)num;
    const int tail_col = max(q - 1, 0);
    const float old_tail = row[tail_col];
    // u-side terms of the two masks, src/direction_mpnn.py:81-89
    const bool a1 = (head_dep <= t) && (num > 0.0f);
    const bool a2 = ((head_dep - t) < -10.0f) && ((maxn - 3.0f) <= num);
    const bool fr = num < (maxn - 3.0f);
    // src/direction_mpnn.py:185-191: t + max(FFTT, cc / (MAXN + 10 - NUM))
    const float tt = max_propagate_nan(fftt, ccn / ((maxn + 10.0f) - num));
    const float dep_new = t + tt;
    const float d = max_propagate_nan((head_dep - head_arr) - fftt, 0.0f);  // :94-95
    const int bits = (a1 ? kA1 : 0) | (a2 ? kA2 : 0) | (fr ? kFree : 0) | (bad ? kBad : 0);
    w.recA[n] = make_float4(head_id, sel, maxn - num, __int_as_float(bits));
    w.recB[n] = make_float4(ridx, num, dep_new, old_tail);
    w.dtt[n] = d;
    if (bad) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_select_append(tarl_dual_csr g, float* __restrict__ x, int64_t stride,
                                                            int Nmax, const float* __restrict__ attr,
                                                            const float* __restrict__ noise, float t, Workspace w,
                                                            int32_t* __restrict__ flags) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= g.n_links) return;
    const float4 A = w.recA[d];
    const float4 B = w.recB[d];
    const int bits_d = __float_as_int(A.w);
    const bool free_d = bits_d & kFree;
    const float room_d = A.z, ridx_d = B.x;

    float best = -FLT_MAX;  // numeric_limits<float>::lowest(), torch-scatter scatter_max init
    int arg = -1;
    float psum = 0.0f;
    const int k1 = g.in_ptr[d + 1];
    for (int k = g.in_ptr[d]; k < k1; ++k) {  // ascending original edge id
        const int u = g.in_src[k];
        const int e = g.in_eid[k];
        const float4 U = w.recA[u];
        const int bits_u = __float_as_int(U.w);
        const bool match = (U.y == ridx_d);
        const bool m = ((bits_u & kA1) && free_d && match) || ((bits_u & kA2) && (U.z <= room_d) && match);
        const float p = attr[e] * (m ? 1.0f : 0.0f);
        psum += p;
        const float un = noise[e];
        const float gum = -logf(-logf(un));          // src/direction_mpnn.py:137
        const float s = logf(p + 1e-12f) + gum;      // :138
        if (s > best) { best = s; arg = u; }         // strict: the lowest edge id wins ties
    }
    float chosen = 0.0f;
    if (psum > 0.0f) {  // :142-144
        if (arg < 0) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_NO_WINNER);
        else chosen = w.recA[arg].x;
    }
    float num_post = B.y, tail_post = B.w, head_post = A.x;
    if (!(bits_d & kBad)) {  // src/direction_mpnn.py:171-195 — on EVERY link
        const int q = (int)B.y;
        float* row = x + (int64_t)d * stride;
        row[q] = chosen;
        row[Nmax + q] = t;
        row[2 * Nmax + q] = B.z;
        if (chosen != 0.0f) {
            num_post = B.y + 1.0f;
            row[3 * Nmax + 1] = num_post;
            tail_post = chosen;
        } else if (q == 0) {
            tail_post = 0.0f;  // slot 0 was just overwritten with id 0
        }
        if (q == 0) head_post = chosen;
    }
    w.post[d] = make_float4(num_post, tail_post, head_post, 0.0f);
}

// delta_travel_time for the direction-only entry point (the fused step emits it from k_respond_shift)
__global__ void __launch_bounds__(kThreads) k_emit_delta_tt(tarl_dual_csr g, Workspace w, float* __restrict__ delta_tt) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= g.n_links) return;
    const float v = w.dtt[u];
    const int k1 = g.out_ptr[u + 1];
    for (int k = g.out_ptr[u]; k < k1; ++k) delta_tt[g.out_eid != nullptr ? g.out_eid[k] : k] = v;
}

// post-append summaries straight from x, for the response-only entry point
__global__ void __launch_bounds__(kThreads) k_snapshot(const float* __restrict__ x, int64_t stride, int N, int Nmax,
                                                       Workspace w, int32_t* __restrict__ flags) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* row = x + (int64_t)n * stride;
    const float num = row[3 * Nmax + 1];
    const bool bad = !(num >= 0.0f) || !(num <= (float)Nmax);
    const int cnt = bad ? 0 : (int)num;
    w.post[n] = make_float4(num, row[max(cnt - 1, 0)], row[0], 0.0f);
    if (bad) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_respond_shift(tarl_dual_csr g, float* __restrict__ x, int64_t stride,
                                                            int Nmax, Workspace w, float* __restrict__ delta_tt,
                                                            uint8_t* __restrict__ pop, int32_t* __restrict__ flags) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool accept = false;
    float num_u = 0.0f;
    if (u < g.n_links) {
        const float4 P = w.post[u];
        num_u = P.x;
        const bool has_up = (long long)P.x > 0;
        const long long head = (long long)P.z;
        const float dv = (delta_tt != nullptr) ? w.dtt[u] : 0.0f;
        const int k1 = g.out_ptr[u + 1];
        for (int k = g.out_ptr[u]; k < k1; ++k) {
            if (delta_tt != nullptr) delta_tt[g.out_eid != nullptr ? g.out_eid[k] : k] = dv;
            const float4 D = w.post[g.out_dst[k]];
            // src/response_mpnn.py:66-83
            accept = accept || (has_up && ((long long)D.x > 0) && ((long long)D.y == head));
        }
        pop[u] = accept ? 1 : 0;
    }
    // FIFO shift of the popping rows, one warp per row (src/response_mpnn.py:119-122)
    unsigned todo = __ballot_sync(0xffffffffu, accept);
    if (todo != 0 && lane == 0) flags[TARL_FLAG_ANY_POP] = 1;
    const int span = Nmax - 1;
    while (todo) {
        const int src_lane = __ffs(todo) - 1;
        todo &= todo - 1;
        const int r = __shfl_sync(0xffffffffu, u, src_lane);
        const float nr = __shfl_sync(0xffffffffu, num_u, src_lane);
        float* row = x + (int64_t)r * stride;
        for (int seg = 0; seg < 3; ++seg) {
            float* q = row + seg * Nmax;
            for (int base = 0; base < span; base += 32) {  // ascending chunks: a chunk reads only cells no earlier chunk wrote
                const int k = base + lane;
                float v = 0.0f;
                if (k < span) v = q[k + 1];
                __syncwarp();
                if (k < span) q[k] = v;
                __syncwarp();
            }
        }
        if (lane == 0) row[3 * Nmax + 1] = nr - 1.0f;
    }
}

inline int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

int check_common(const tarl_dual_csr* g, const float* x, int32_t nmax, const void* ws, size_t ws_bytes,
                 const int32_t* flags) {
    if (g == nullptr || flags == nullptr || nmax < 2 || g->n_links < 0 || g->n_edges < 0) return TARL_E_BADARG;
    if (g->n_links > 0 && (x == nullptr || g->in_ptr == nullptr || g->out_ptr == nullptr)) return TARL_E_BADARG;
    if (g->n_edges > 0 && (g->in_src == nullptr || g->in_eid == nullptr || g->out_dst == nullptr))
        return TARL_E_BADARG;
    if (g->n_links > 0 && (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 15) != 0)) return TARL_E_WORKSPACE;
    if (ws_bytes < tarl_core_workspace_bytes(g->n_links)) return TARL_E_WORKSPACE;
    return TARL_OK;
}

int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

}  // namespace

extern "C" {

int tarl_abi_version(void) { return TARL_ABI_VERSION; }

const char* tarl_error_string(int code) {
    switch (code) {
        case TARL_OK: return "ok";
        case TARL_E_BADARG: return "bad argument (null pointer, negative size or nmax < 2)";
        case TARL_E_WORKSPACE: return "workspace missing, misaligned or smaller than tarl_core_workspace_bytes()";
        case TARL_E_LAUNCH: return "CUDA kernel launch failed";
        default: return "unknown error";
    }
}

size_t tarl_core_workspace_bytes(int32_t n_links) {
    const size_t n = n_links > 0 ? (size_t)n_links : 0;
    return 3 * align16(sizeof(float4) * n) + align16(sizeof(float) * n);
}

int tarl_direction_forward(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax,
                           const float* edge_attr, const float* cc, const float* noise, const float* sel, float t,
                           float* delta_tt,
                           int32_t* flags, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(g, x, nmax, workspace, workspace_bytes, flags);
    if (rc != TARL_OK) return rc;
    if (g->n_edges > 0 && (edge_attr == nullptr || noise == nullptr)) return TARL_E_BADARG;
    if (g->n_links == 0) return TARL_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(workspace, g->n_links);
    const int nb = blocks_for(g->n_links);
    k_offer<<<nb, kThreads, 0, s>>>(x, x_row_stride, g->n_links, nmax, cc, sel, t, w, flags);
    k_select_append<<<nb, kThreads, 0, s>>>(*g, x, x_row_stride, nmax, edge_attr, noise, t, w, flags);
    if (delta_tt != nullptr && g->n_edges > 0) k_emit_delta_tt<<<nb, kThreads, 0, s>>>(*g, w, delta_tt);
    return launch_status();
}

int tarl_response_forward(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax, uint8_t* pop,
                          int32_t* flags, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(g, x, nmax, workspace, workspace_bytes, flags);
    if (rc != TARL_OK) return rc;
    if (g->n_links == 0) return TARL_OK;
    if (pop == nullptr) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(workspace, g->n_links);
    const int nb = blocks_for(g->n_links);
    k_snapshot<<<nb, kThreads, 0, s>>>(x, x_row_stride, g->n_links, nmax, w, flags);
    k_respond_shift<<<nb, kThreads, 0, s>>>(*g, x, x_row_stride, nmax, w, nullptr, pop, flags);
    return launch_status();
}

int tarl_core_step_phases(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax,
                          const float* edge_attr, const float* cc, const float* noise, const float* sel, float t,
                          float* delta_tt,
                          uint8_t* pop, int32_t* flags, void* workspace, size_t workspace_bytes, void* stream,
                          uint32_t phase_mask) {
    int rc = check_common(g, x, nmax, workspace, workspace_bytes, flags);
    if (rc != TARL_OK) return rc;
    if (g->n_edges > 0 && (edge_attr == nullptr || noise == nullptr)) return TARL_E_BADARG;
    if (g->n_links == 0) return TARL_OK;
    if (pop == nullptr) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Workspace w = carve(workspace, g->n_links);
    const int nb = blocks_for(g->n_links);
    if (phase_mask & TARL_PHASE_OFFER)
        k_offer<<<nb, kThreads, 0, s>>>(x, x_row_stride, g->n_links, nmax, cc, sel, t, w, flags);
    if (phase_mask & TARL_PHASE_SELECT_APPEND)
        k_select_append<<<nb, kThreads, 0, s>>>(*g, x, x_row_stride, nmax, edge_attr, noise, t, w, flags);
    if (phase_mask & TARL_PHASE_RESPOND_SHIFT)
        k_respond_shift<<<nb, kThreads, 0, s>>>(*g, x, x_row_stride, nmax, w, delta_tt, pop, flags);
    return launch_status();
}

int tarl_core_step(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax, const float* edge_attr,
                   const float* cc, const float* noise, const float* sel, float t, float* delta_tt, uint8_t* pop,
                   int32_t* flags, void* workspace, size_t workspace_bytes, void* stream) {
    return tarl_core_step_phases(g, x, x_row_stride, nmax, edge_attr, cc, noise, sel, t, delta_tt, pop, flags, workspace,
                                 workspace_bytes, stream, TARL_PHASE_ALL);
}

}  // extern "C"

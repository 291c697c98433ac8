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
#include "engine_common.cuh"
#include "tile_map.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kLogEps = 1e-8f;  // src/reinforcement_learning.py:27

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

// ------------------------------------------------------------------------------------------------ policy embedding
// Layout: everything the policy kernels produce is NODE-major / EDGE-major with the batch row innermost (element (b, n)
// of emb / idx / node_grad at n*B + b, element (b, e) of logits at e*B + b), so that the B rows of one node or edge are
// one contiguous vector: gathers by node id fetch all rows in one sector and edge ids are read once for all rows.
// node pass: idx[n,b] = ROAD_INDEX >= 0 ? ROAD_INDEX : n (D2);  emb[n,b] = W[idx]. Tiled (tile_map.cuh): the ROAD_INDEX
// column of the row-major [B, N, C] observation is read with the node innermost (the 32 lanes of a warp stay inside
// one ~1 KB span of one sample), idx / emb are written with the row innermost (one contiguous B-vector per node).
__global__ void __launch_bounds__(tarl::kTileThreads) k_policy_node(const float* __restrict__ w, int rows,
                                                                    const float* __restrict__ nf, int64_t nf_bs,
                                                                    int64_t nf_rs, int ridx_col, int B, int Bp, int N,
                                                                    float* __restrict__ emb, int32_t* __restrict__ idx,
                                                                    int32_t* __restrict__ flags) {
    __shared__ float sm_w[tarl::kTileSmem];
    __shared__ int32_t sm_k[tarl::kTileSmem];
    const tarl::Tile t = tarl::tile_here(B, Bp);
    tarl::tile_walk_nodes(t, [&](int r, int j) {
        const int n = t.n0 + j, b = t.b0 + r;
        if (n >= N || r >= t.nrows) return;
        long long k = (long long)nf[b * nf_bs + n * nf_rs + ridx_col];  // .to(torch.long): truncation
        if (k < 0) k = n;
        if (k >= rows) {  // nn.Embedding raises IndexError
            atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_EMBED_RANGE);
            k = 0;
        }
        sm_k[tarl::tile_slot(t, r, j)] = (int32_t)k;
        sm_w[tarl::tile_slot(t, r, j)] = w[k];
    });
    __syncthreads();
    tarl::tile_walk_rows(t, [&](int r, int j) {
        const int n = t.n0 + j;
        if (n >= N || r >= t.nrows) return;
        const int64_t i = (int64_t)n * B + t.b0 + r;
        idx[i] = sm_k[tarl::tile_slot(t, r, j)];
        emb[i] = sm_w[tarl::tile_slot(t, r, j)];
    });
}

// edge pass: logits[e, :] = emb[dst[e], :]. One thread per (edge, chunk of 4 rows) when B % 4 == 0 (the threads of one
// edge copy one contiguous B-vector with 128-bit accesses), else one thread per (edge, row).
__global__ void __launch_bounds__(kThreads) k_policy_edge(const float* __restrict__ emb, const int32_t* __restrict__ dst,
                                                          int B, int E, float* __restrict__ logits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((B & 3) == 0) {
        const int C = B >> 2;
        if (i >= (int64_t)E * C) return;
        const int e = (int)(i / C), c = (int)(i - (int64_t)e * C);
        reinterpret_cast<float4*>(logits)[i] = reinterpret_cast<const float4*>(emb)[(int64_t)dst[e] * C + c];
    } else {
        if (i >= (int64_t)E * B) return;
        const int e = (int)(i / B), b = (int)(i - (int64_t)e * B);
        logits[i] = emb[(int64_t)dst[e] * B + b];
    }
}

// backward node pass: G[n,b] = sum over in-edges (ascending edge id) of grad_logits[b,e]. Edge-major gradient with
// B % 4 == 0: one thread per (node, chunk of 4 rows), 128-bit gathers; any other strides: one thread per (node, row).
__global__ void __launch_bounds__(kThreads) k_policy_node_grad(tarl_csr in, const float* __restrict__ gl, int64_t g_sb,
                                                               int64_t g_se, int B, float* __restrict__ G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g_sb == 1 && g_se == B && (B & 3) == 0) {
        const int C = B >> 2;
        if (i >= (int64_t)in.n_rows * C) return;
        const int n = (int)(i / C), c = (int)(i - (int64_t)n * C);
        const int k1 = in.ptr[n + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = in.ptr[n]; k < k1; ++k) {
            const float4 v = reinterpret_cast<const float4*>(gl)[(int64_t)in.eid[k] * C + c];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(G)[i] = acc;
    } else {
        if (i >= (int64_t)in.n_rows * B) return;
        const int n = (int)(i / B), b = (int)(i - (int64_t)n * B);
        const int k1 = in.ptr[n + 1];
        float acc = 0.0f;
        for (int k = in.ptr[n]; k < k1; ++k) acc += gl[b * g_sb + in.eid[k] * g_se];
        G[i] = acc;
    }
}

// backward weight pass: gradW[idx[n,b]] += G[n,b]. A group of Bp lanes (Bp = B rounded up to a power of two, <= 32) owns
// one node and reads its rows coalesced. Rows that share the index of row 0 (always, in practice: the ROAD_INDEX column
// is static) are summed by a fixed shuffle tree; with an injective index map every weight row is then the target of
// exactly one atomicAdd, i.e. the result is deterministic. Rows with a different index fall back to their own atomicAdd.
__global__ void __launch_bounds__(kThreads) k_policy_weight_grad(const float* __restrict__ G, const int32_t* __restrict__ idx,
                                                                 int B, int Bp, int N, float* __restrict__ gw) {
    const int sh = 31 - __clz(Bp);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)(i >> sh), lane_b = (int)(i & (Bp - 1));
    const bool node_ok = n < N;
    float acc = 0.0f;
    int k0 = 0;
    if (node_ok) {
        k0 = idx[(int64_t)n * B];
        for (int b = lane_b; b < B; b += Bp) {
            const int k = idx[(int64_t)n * B + b];
            const float g = G[(int64_t)n * B + b];
            if (k == k0) acc += g;
            else atomicAdd(&gw[k], g);
        }
    }
    for (int o = Bp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (node_ok && lane_b == 0) atomicAdd(&gw[k0], acc);
}

// ------------------------------------------------------------------------------------------------ GraphDistribution
// Thread mapping: one thread per (source group, batch row) with the BATCH ROW INNERMOST — lanes [0, Bp) of a warp hold
// the Bp rows of one group (Bp = batch rounded up to a power of two, <= 32), the next Bp lanes the next group. With
// edge-major tensors (element (b, e) at e*B + b: what MPNNPolicyNet emits) the Bp lanes of a group read one
// contiguous B-vector per edge and the group's edge ids are loaded once per group, not once per row. Any strides are
// accepted (a row-major [B,E] tensor is correct, just less coalesced). A group's values are cached in registers
// (kCache edges) so that logits are read once and every exp is evaluated once; longer groups recompute their tail.
// 5 = the out-degree of a road link in the full graph of a 4-way network (4 turns incl. the U-turn + its DEST edge):
// every unrolled, predicated slot beyond the real degree costs registers and issue slots (6 -> 5: backward 0.69 ->
// 0.54 ms, forward 0.41 -> 0.37 ms at 32 rows x 6.0 M edges).
constexpr int kCache = 5;

struct View {          // element (b, e) of a [B, E] tensor at base[b*sb + e*se]
    int64_t sb, se;
};

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

// Sums v over the threads of the block that share (threadIdx.x % Bp); thread b < Bp returns row b's sum (fixed order).
template <typename T>
__device__ __forceinline__ T block_sum_rows(T v, int Bp, T* smem /* [kThreads/32 * 32] */) {
    for (int o = 16; o >= Bp; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane < Bp) smem[wid * 32 + lane] = v;
    __syncthreads();
    T r = T(0);
    if (threadIdx.x < Bp)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += smem[w * 32 + threadIdx.x];
    return r;
}

struct GroupRows {     // what thread (g, b) needs to walk its group
    int g, b, k0, k1;
    bool live;
};

// Bp is a power of two: tile `tile` of the grid covers groups [tile*kThreads/Bp, (tile+1)*kThreads/Bp).
__device__ __forceinline__ GroupRows locate(const tarl_csr& grp, int B, int Bp, int b_chunk, int tile) {
    const int sh = 31 - __clz(Bp);
    const int64_t i = (int64_t)tile * blockDim.x + threadIdx.x;
    GroupRows t;
    t.g = (int)(i >> sh);
    t.b = b_chunk * 32 + (int)(i & (Bp - 1));
    t.live = t.g < grp.n_rows && t.b < B;
    t.k0 = t.live ? grp.ptr[t.g] : 0;
    t.k1 = t.live ? grp.ptr[t.g + 1] : 0;
    return t;
}

// Softmax of one group for one row. The first kCache edges live in registers (ex[j] = exp(z_j - max), statically
// indexed: every loop over them is fully unrolled and predicated on j < deg); longer groups recompute the tail.
struct Soft {
    float ex[kCache];
    float mx, inv_den;
    int deg;
};

// Fast-math forms on the per-(edge, row) path, all far inside the 1e-5 relative / 1e-6 absolute parity bar of the
// tests: __fdividef 2 ulp; __expf on arguments <= 0 (softmax after max subtraction) has relative error ~|x|*1e-7;
// __logf has absolute error <= 2^-21.4, the size of one ulp of log(p) for |log p| ~ 1 and negligible next to the
// p * log(p) products it feeds. With them the kernels are bandwidth- instead of instruction-bound.
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float flog(float x) { return __logf(x); }

__device__ __forceinline__ void soft_load(Soft& s, const tarl_csr& grp, const float* __restrict__ lg, View lv, int b, int k0,
                                          int k1, float temp) {
    s.deg = k1 - k0;
    s.mx = -FLT_MAX;
    const float* row = lg + b * lv.sb;
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        s.ex[j] = -FLT_MAX;
        if (j < s.deg) {
            s.ex[j] = fdiv(row[grp.eid[k0 + j] * lv.se], temp);
            s.mx = fmaxf(s.mx, s.ex[j]);
        }
    }
    for (int k = k0 + kCache; k < k1; ++k) s.mx = fmaxf(s.mx, fdiv(row[grp.eid[k] * lv.se], temp));
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < kCache; ++j)
        if (j < s.deg) { s.ex[j] = fexp(s.ex[j] - s.mx); den += s.ex[j]; }
    for (int k = k0 + kCache; k < k1; ++k) den += fexp(fdiv(row[grp.eid[k] * lv.se], temp) - s.mx);
    s.inv_den = fdiv(1.0f, den);
}

__device__ __forceinline__ float tail_p(const Soft& s, const float* __restrict__ lg, View lv, int b, int e, float temp) {
    return fexp(fdiv(lg[b * lv.sb + e * lv.se], temp) - s.mx) * s.inv_den;
}

struct FwdAcc {
    float ent, lp, asum, best;
    int best_e;
};

__device__ __forceinline__ void fwd_edge(FwdAcc& a, float p, int e, int b, const void* __restrict__ action, View av,
                                         int action_dtype, float* __restrict__ proba, View pv) {
    const float l = flog(p + kLogEps);
    a.ent -= p * l;
    if (action != nullptr) {
        const float x = load_action(action, action_dtype, b * av.sb + e * av.se);
        a.lp += x * l;
        a.asum += x;
    }
    if (proba != nullptr) proba[b * pv.sb + e * pv.se] = p;
    if (p > a.best) { a.best = p; a.best_e = e; }
}

// Softmax over the group's edges in ascending edge id, then whatever of {proba, entropy, log_prob, mode} was asked
// for. Each CTA walks tiles of groups with a fixed stride and leaves one partial sum per (CTA, row): the [B]
// reductions stay deterministic and the second stage stays small.
__global__ void __launch_bounds__(kThreads) k_gd_forward(tarl_csr grp, const float* __restrict__ logits, View lv,
                                                         float temp, int B, int Bp, int n_tiles,
                                                         const void* __restrict__ action, View av, int action_dtype,
                                                         float* __restrict__ proba, View pv, float* __restrict__ mode,
                                                         View mv, float* __restrict__ part_ent,
                                                         float* __restrict__ part_lp, int32_t* __restrict__ part_bad) {
    __shared__ float sm_f[kThreads];
    __shared__ int sm_i[kThreads];
    float ent = 0.0f, lp = 0.0f;
    int bad = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const GroupRows t = locate(grp, B, Bp, blockIdx.y, tile);
        if (!t.live || t.k1 == t.k0) continue;
        Soft s;
        soft_load(s, grp, logits, lv, t.b, t.k0, t.k1, temp);
        FwdAcc a = {0.0f, 0.0f, 0.0f, -FLT_MAX, -1};
#pragma unroll
        for (int j = 0; j < kCache; ++j)
            if (j < s.deg) fwd_edge(a, s.ex[j] * s.inv_den, grp.eid[t.k0 + j], t.b, action, av, action_dtype, proba, pv);
        for (int k = t.k0 + kCache; k < t.k1; ++k) {
            const int e = grp.eid[k];
            fwd_edge(a, tail_p(s, logits, lv, t.b, e, temp), e, t.b, action, av, action_dtype, proba, pv);
        }
        ent += a.ent; lp += a.lp;
        if (action != nullptr && a.asum != 1.0f) bad = 1;   // not exactly one selected edge in this group (:86-88)
        if (mode != nullptr && a.best_e >= 0) mode[t.b * mv.sb + a.best_e * mv.se] = 1.0f;
    }
    const int nb = gridDim.x;
    const int row = blockIdx.y * 32 + threadIdx.x;      // valid for threadIdx.x < Bp
    if (part_ent != nullptr) {
        const float v = block_sum_rows(ent, Bp, sm_f);
        if (threadIdx.x < Bp && row < B) part_ent[(int64_t)row * nb + blockIdx.x] = v;
    }
    if (part_lp != nullptr) {
        const float v = block_sum_rows(lp, Bp, sm_f);
        const int c = block_sum_rows(bad, Bp, sm_i);
        if (threadIdx.x < Bp && row < B) {
            part_lp[(int64_t)row * nb + blockIdx.x] = v;
            part_bad[(int64_t)row * nb + blockIdx.x] = c;
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
__global__ void __launch_bounds__(kThreads) k_gd_backward(tarl_csr grp, const float* __restrict__ logits, View lv,
                                                          float temp, int B, int Bp, const void* __restrict__ action,
                                                          View av, int action_dtype, const float* __restrict__ g_lp,
                                                          const float* __restrict__ g_ent,
                                                          const float* __restrict__ log_prob, float* __restrict__ grad,
                                                          View gv) {
    const GroupRows t = locate(grp, B, Bp, blockIdx.y, blockIdx.x);
    if (!t.live || t.k1 == t.k0) return;
    float wl = (g_lp != nullptr && action != nullptr) ? g_lp[t.b] : 0.0f;
    if (log_prob != nullptr && log_prob[t.b] == -INFINITY) wl = 0.0f;
    const float we = (g_ent != nullptr) ? g_ent[t.b] : 0.0f;
    Soft s;
    soft_load(s, grp, logits, lv, t.b, t.k0, t.k1, temp);
    float Sa = 0.0f, Sc = 0.0f;
    float cc[kCache], ar[kCache];       // c_e and a_e r_e of the cached edges
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        cc[j] = 0.0f; ar[j] = 0.0f;
        if (j < s.deg) {
            const float p = s.ex[j] * s.inv_den;
            const float r = fdiv(p, p + kLogEps);
            cc[j] = flog(p + kLogEps) + r;
            if (wl != 0.0f) ar[j] = load_action(action, action_dtype, t.b * av.sb + grp.eid[t.k0 + j] * av.se) * r;
            Sa += ar[j];
            Sc += p * cc[j];
        }
    }
    for (int k = t.k0 + kCache; k < t.k1; ++k) {
        const int e = grp.eid[k];
        const float p = tail_p(s, logits, lv, t.b, e, temp);
        const float r = fdiv(p, p + kLogEps);
        if (wl != 0.0f) Sa += load_action(action, action_dtype, t.b * av.sb + e * av.se) * r;
        Sc += p * (flog(p + kLogEps) + r);
    }
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg) {
            const float p = s.ex[j] * s.inv_den;
            float v = -we * p * (cc[j] - Sc);
            if (wl != 0.0f) v += wl * (ar[j] - p * Sa);
            grad[t.b * gv.sb + grp.eid[t.k0 + j] * gv.se] = fdiv(v, temp);
        }
    }
    for (int k = t.k0 + kCache; k < t.k1; ++k) {
        const int e = grp.eid[k];
        const float p = tail_p(s, logits, lv, t.b, e, temp);
        const float r = fdiv(p, p + kLogEps);
        const float c = flog(p + kLogEps) + r;
        float v = -we * p * (c - Sc);
        if (wl != 0.0f) v += wl * (load_action(action, action_dtype, t.b * av.sb + e * av.se) * r - p * Sa);
        grad[t.b * gv.sb + e * gv.se] = fdiv(v, temp);
    }
}

// inverse-CDF sample, one uniform per (row, group): first edge (ascending edge id, D3) with u < cumulative proba (:62-80)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_gd_sample(tarl_csr grp, const float* __restrict__ logits, View lv, float temp,
                                                        int B, int Bp, const float* __restrict__ u, View uv,
                                                        T* __restrict__ onehot, View ov) {
    const GroupRows t = locate(grp, B, Bp, blockIdx.y, blockIdx.x);
    if (!t.live || t.k1 == t.k0) return;
    Soft s;
    soft_load(s, grp, logits, lv, t.b, t.k0, t.k1, temp);
    const float ug = u[t.b * uv.sb + t.g * uv.se];
    float cum = 0.0f;
    int hit = -1;
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg && hit < 0) {
            cum += s.ex[j] * s.inv_den;
            if (ug < cum) hit = grp.eid[t.k0 + j];
        }
    }
    for (int k = t.k0 + kCache; k < t.k1 && hit < 0; ++k) {
        const int e = grp.eid[k];
        cum += tail_p(s, logits, lv, t.b, e, temp);
        if (ug < cum) hit = e;
    }
    // Every edge of the group is written (1 at the hit, 0 elsewhere): with edge-major output the lanes of a warp store
    // consecutive bytes of one edge's row vector, whereas writing only the hit — a different edge in every row —
    // would dirty one sector per (group, row). Every edge has a source, so the whole one-hot tensor is written here.
    for (int k = t.k0; k < t.k1; ++k) {
        const int e = grp.eid[k];
        onehot[t.b * ov.sb + e * ov.se] = T(e == hit ? 1 : 0);
    }
}

// ------------------------------------------------------------------------------------------------ 4 rows per thread
// Fast path of the two training-time kernels for the layout the package itself produces: edge-major logits / gradient
// (row stride 1, edge stride B), B in {4, 8, 16, 32} per grid.y slice, action absent or edge-major bytes. One thread
// owns (group, 4 consecutive rows): every logits access is one 128-bit load, the action of an edge is one 32-bit load,
// and edge ids / predicates / addresses are paid once per four rows.
struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 ld4(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    return F4{{t.x, t.y, t.z, t.w}};
}
// logits of one edge for 4 rows: edge-major [E][B], or ONE row shared by every batch row (kBcast: an
// expanded [1, E] tensor — the active policy path does not depend on the dynamic observation, so a rollout over R
// replicas has a single logits row)
template <bool kBcast>
__device__ __forceinline__ F4 ld4_logits(const float* lg, int B, int e, int row0) {
    if (kBcast) { const float z = lg[e]; return F4{{z, z, z, z}}; }
    return ld4(lg + (int64_t)e * B + row0);
}
__device__ __forceinline__ void st4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

struct Soft4 {
    F4 ex[kCache];
    F4 mx, inv_den;
    int eid[kCache];
    uint32_t aw[kCache];   // the 4 action bytes of each cached edge, fetched together with its logits
    int deg;
};

// s.eid[0 .. min(deg, kCache)) and s.deg are filled by the caller (prefetched one tile ahead)
template <bool kBcast = false>
__device__ __forceinline__ void soft4_load(Soft4& s, const tarl_csr& grp, const float* __restrict__ lg,
                                           const uint8_t* __restrict__ act, int B, int row0, int k0, int k1,
                                           float inv_t) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s.mx.v[q] = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        s.aw[j] = 0u;
        if (act != nullptr && j < s.deg) s.aw[j] = *reinterpret_cast<const uint32_t*>(act + (int64_t)s.eid[j] * B + row0);
    }
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg) {
            s.ex[j] = ld4_logits<kBcast>(lg, B, s.eid[j], row0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { s.ex[j].v[q] *= inv_t; s.mx.v[q] = fmaxf(s.mx.v[q], s.ex[j].v[q]); }
        }
    }
    for (int k = k0 + kCache; k < k1; ++k) {
        const F4 z = ld4_logits<kBcast>(lg, B, grp.eid[k], row0);
#pragma unroll
        for (int q = 0; q < 4; ++q) s.mx.v[q] = fmaxf(s.mx.v[q], z.v[q] * inv_t);
    }
    F4 den = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { s.ex[j].v[q] = fexp(s.ex[j].v[q] - s.mx.v[q]); den.v[q] += s.ex[j].v[q]; }
        }
    }
    for (int k = k0 + kCache; k < k1; ++k) {
        const F4 z = ld4_logits<kBcast>(lg, B, grp.eid[k], row0);
#pragma unroll
        for (int q = 0; q < 4; ++q) den.v[q] += fexp(z.v[q] * inv_t - s.mx.v[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) s.inv_den.v[q] = fdiv(1.0f, den.v[q]);
}

template <bool kBcast = false>
__device__ __forceinline__ F4 tail4_p(const Soft4& s, const float* __restrict__ lg, int B, int row0, int e, float inv_t) {
    F4 z = ld4_logits<kBcast>(lg, B, e, row0);
#pragma unroll
    for (int q = 0; q < 4; ++q) z.v[q] = fexp(z.v[q] * inv_t - s.mx.v[q]) * s.inv_den.v[q];
    return z;
}

__device__ __forceinline__ F4 action4_of(uint32_t w) {
    return F4{{(float)(w & 0xffu), (float)((w >> 8) & 0xffu), (float)((w >> 16) & 0xffu), (float)(w >> 24)}};
}
__device__ __forceinline__ F4 action4(const uint8_t* __restrict__ act, int B, int row0, int e) {
    return action4_of(*reinterpret_cast<const uint32_t*>(act + (int64_t)e * B + row0));
}

// (group, 4-row chunk) of this thread in tile `tile`; C = chunks per group (power of two), rows offset by grid.y * 32
struct Where4 {
    int g, row0, k0, k1;
    bool live;
};
__device__ __forceinline__ Where4 locate4(const tarl_csr& grp, int B, int C, int tile) {
    const int sh = 31 - __clz(C);
    const int64_t i = (int64_t)tile * blockDim.x + threadIdx.x;
    Where4 t;
    t.g = (int)(i >> sh);
    t.row0 = blockIdx.y * 32 + 4 * (int)(i & (C - 1));
    t.live = t.g < grp.n_rows && t.row0 < B;
    t.k0 = t.live ? grp.ptr[t.g] : 0;
    t.k1 = t.live ? grp.ptr[t.g + 1] : 0;
    return t;
}

// Software pipeline of the persistent em4 kernels: a thread walks tiles with a fixed stride; while tile i computes, the
// edge ids of tile i+1 and the group bounds of tile i+2 are already in flight, so that only ONE level of the
// ptr -> edge id -> logits chain (the logits / action gathers) is exposed per tile.
struct Stage4 {
    int row0, k0, deg;     // deg == 0: nothing to do (dead lane, empty group, or past the last tile)
    int g;
};
__device__ __forceinline__ Stage4 stage4_bounds(const tarl_csr& grp, int B, int C, int tile, int n_tiles) {
    Stage4 st = {0, 0, 0, 0};
    if (tile < n_tiles) {
        const Where4 t = locate4(grp, B, C, tile);
        st.row0 = t.row0; st.k0 = t.k0; st.deg = t.k1 - t.k0; st.g = t.g;
    }
    return st;
}
__device__ __forceinline__ void stage4_eids(const tarl_csr& grp, const Stage4& st, int (&eid)[kCache]) {
#pragma unroll
    for (int j = 0; j < kCache; ++j) eid[j] = (j < st.deg) ? grp.eid[st.k0 + j] : 0;
}

__global__ void __launch_bounds__(kThreads, 3) k_gd_forward_em4(tarl_csr grp, const float* __restrict__ logits, float inv_t,
                                                             int B, int C, int n_tiles, const uint8_t* __restrict__ action,
                                                             float* __restrict__ part_ent, float* __restrict__ part_lp,
                                                             int32_t* __restrict__ part_bad) {
    __shared__ float sm_f[kThreads];
    __shared__ int sm_i[kThreads];
    F4 ent = {{0.f, 0.f, 0.f, 0.f}}, lp = {{0.f, 0.f, 0.f, 0.f}};
    int bad[4] = {0, 0, 0, 0};
    const int stride = gridDim.x;
    Stage4 nxt = stage4_bounds(grp, B, C, blockIdx.x, n_tiles);
    int eid_nxt[kCache];
    stage4_eids(grp, nxt, eid_nxt);
    Stage4 nxt2 = stage4_bounds(grp, B, C, blockIdx.x + stride, n_tiles);
    for (int tile = blockIdx.x; tile < n_tiles; tile += stride) {
        struct { int row0, k0, k1; } t = {nxt.row0, nxt.k0, nxt.k0 + nxt.deg};
        Soft4 s;
        s.deg = nxt.deg;
#pragma unroll
        for (int j = 0; j < kCache; ++j) s.eid[j] = eid_nxt[j];
        nxt = nxt2;
        stage4_eids(grp, nxt, eid_nxt);
        nxt2 = stage4_bounds(grp, B, C, tile + 2 * stride, n_tiles);
        if (s.deg == 0) continue;
        soft4_load(s, grp, logits, action, B, t.row0, t.k0, t.k1, inv_t);
        F4 asum = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int j = 0; j < kCache; ++j) {
            if (j < s.deg) {
                const F4 a = action4_of(s.aw[j]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float p = s.ex[j].v[q] * s.inv_den.v[q];
                    const float l = flog(p + kLogEps);
                    ent.v[q] -= p * l;
                    lp.v[q] += a.v[q] * l;
                    asum.v[q] += a.v[q];
                }
            }
        }
        for (int k = t.k0 + kCache; k < t.k1; ++k) {
            const int e = grp.eid[k];
            const F4 p4 = tail4_p(s, logits, B, t.row0, e, inv_t);
            F4 a = {{0.f, 0.f, 0.f, 0.f}};
            if (action != nullptr) a = action4(action, B, t.row0, e);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float l = flog(p4.v[q] + kLogEps);
                ent.v[q] -= p4.v[q] * l;
                lp.v[q] += a.v[q] * l;
                asum.v[q] += a.v[q];
            }
        }
        if (action != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (asum.v[q] != 1.0f) bad[q] = 1;
        }
    }
    // per-row partials: lanes that share (threadIdx.x % C) hold the same four rows
    const int nb = gridDim.x;
    const int c = threadIdx.x & (C - 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = blockIdx.y * 32 + 4 * c + q;
        if (part_ent != nullptr) {
            const float v = block_sum_rows(ent.v[q], C, sm_f);
            if (threadIdx.x < C && row < B) part_ent[(int64_t)row * nb + blockIdx.x] = v;
        }
        if (part_lp != nullptr) {
            const float v = block_sum_rows(lp.v[q], C, sm_f);
            const int n = block_sum_rows(bad[q], C, sm_i);
            if (threadIdx.x < C && row < B) {
                part_lp[(int64_t)row * nb + blockIdx.x] = v;
                part_bad[(int64_t)row * nb + blockIdx.x] = n;
            }
        }
    }
}

// Sampling on the same fast path: inverse CDF per (group, row) with one uniform each, the one-hot written as one
// 32-bit store per (edge, 4 rows) — every edge has a source, so the whole tensor is written — and, optionally, the
// log-probability of what was drawn (sum over the groups of log(p_hit + eps); a group without a hit makes the row
// -inf, as log_prob() of that action would, src/reinforcement_learning.py:86-91) accumulated in the same pass, so that
// a rollout needs no second sweep over the logits.
template <bool kBcast>
__global__ void __launch_bounds__(kThreads, 2) k_gd_sample_em4(tarl_csr grp, const float* __restrict__ logits,
                                                            float inv_t, int B, int C, int n_tiles,
                                                            const float* __restrict__ u,
                                                            int64_t u_sb, int64_t u_sg, uint8_t* __restrict__ onehot,
                                                            float* __restrict__ part_lp, int32_t* __restrict__ part_bad) {
    __shared__ float sm_f[kThreads];
    __shared__ int sm_i[kThreads];
    F4 lp = {{0.f, 0.f, 0.f, 0.f}};
    int bad[4] = {0, 0, 0, 0};
    const int stride = gridDim.x;
    Stage4 nxt = stage4_bounds(grp, B, C, blockIdx.x, n_tiles);
    int eid_nxt[kCache];
    stage4_eids(grp, nxt, eid_nxt);
    Stage4 nxt2 = stage4_bounds(grp, B, C, blockIdx.x + stride, n_tiles);
    for (int tile = blockIdx.x; tile < n_tiles; tile += stride) {
        struct { int row0, k0, k1, g; } t = {nxt.row0, nxt.k0, nxt.k0 + nxt.deg, nxt.g};
        Soft4 s;
        s.deg = nxt.deg;
#pragma unroll
        for (int j = 0; j < kCache; ++j) s.eid[j] = eid_nxt[j];
        nxt = nxt2;
        stage4_eids(grp, nxt, eid_nxt);
        nxt2 = stage4_bounds(grp, B, C, tile + 2 * stride, n_tiles);
        if (s.deg == 0) continue;
        float ug[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ug[q] = u[(t.row0 + q) * u_sb + t.g * u_sg];
        soft4_load<kBcast>(s, grp, logits, nullptr, B, t.row0, t.k0, t.k1, inv_t);
        F4 cum = {{0.f, 0.f, 0.f, 0.f}};
        int hit[4] = {-1, -1, -1, -1};         // position inside the group
        float ph[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kCache; ++j) {
            if (j < s.deg) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float p = s.ex[j].v[q] * s.inv_den.v[q];
                    cum.v[q] += p;
                    if (hit[q] < 0 && ug[q] < cum.v[q]) { hit[q] = j; ph[q] = p; }
                }
            }
        }
        for (int k = t.k0 + kCache; k < t.k1; ++k) {
            const F4 p4 = tail4_p<kBcast>(s, logits, B, t.row0, grp.eid[k], inv_t);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cum.v[q] += p4.v[q];
                if (hit[q] < 0 && ug[q] < cum.v[q]) { hit[q] = k - t.k0; ph[q] = p4.v[q]; }
            }
        }
#pragma unroll
        for (int j = 0; j < kCache; ++j) {
            if (j < s.deg) {
                const uint32_t w = (hit[0] == j ? 1u : 0u) | (hit[1] == j ? 1u << 8 : 0u) | (hit[2] == j ? 1u << 16 : 0u) |
                                   (hit[3] == j ? 1u << 24 : 0u);
                *reinterpret_cast<uint32_t*>(onehot + (int64_t)s.eid[j] * B + t.row0) = w;
            }
        }
        for (int k = t.k0 + kCache; k < t.k1; ++k) {
            const int j = k - t.k0;
            const uint32_t w = (hit[0] == j ? 1u : 0u) | (hit[1] == j ? 1u << 8 : 0u) | (hit[2] == j ? 1u << 16 : 0u) |
                               (hit[3] == j ? 1u << 24 : 0u);
            *reinterpret_cast<uint32_t*>(onehot + (int64_t)grp.eid[k] * B + t.row0) = w;
        }
        if (part_lp != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (hit[q] < 0) bad[q] = 1;
                else lp.v[q] += flog(ph[q] + kLogEps);
            }
        }
    }
    if (part_lp != nullptr) {
        const int nb = gridDim.x;
        const int c = threadIdx.x & (C - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int row = blockIdx.y * 32 + 4 * c + q;
            const float v = block_sum_rows(lp.v[q], C, sm_f);
            const int n = block_sum_rows(bad[q], C, sm_i);
            if (threadIdx.x < C && row < B) {
                part_lp[(int64_t)row * nb + blockIdx.x] = v;
                part_bad[(int64_t)row * nb + blockIdx.x] = n;
            }
        }
    }
}

// Rollout sampling: ONE logits row for every batch row (the active policy path does not read the dynamic observation,
// so R replicas share their logits). The softmax of a group is then the same for every replica: a CTA computes the
// cumulative distribution of 32 consecutive groups ONCE (phase 0, shared memory), and the per-replica work left is one
// uniform, <= 8 compares and the one-hot bytes (phase 1: thread = (group, 4 replicas), replica innermost — 128-bit
// uniform loads, 32-bit one-hot stores, a group's row vector contiguous). Optionally the draw is APPLIED in the same
// pass — SELECTED_ROAD[replica, source node] = target of the drawn edge (src/reinforcement_learning.py:223-231) —
// which needs the node innermost: the hit positions are parked as bytes in shared memory and phase 2 walks them with
// the 32 groups on the lanes (128 contiguous bytes of SELECTED_ROAD per replica). The one-hot is still written (the
// trajectory keeps it for the PPO update), but nobody has to read it back to step the environment.
// Arithmetic per (group, row) is exactly that of k_gd_sample / k_gd_sample_em4 (z * inv_t, max, __expf, sum in
// ascending edge id, one reciprocal, cumulative sum), so the three kernels draw the same edges from the same uniforms.
constexpr int kBcGroups = 32;         // groups per tile
constexpr int kBcDeg = 8;             // edges per group kept in shared memory; longer groups take the slow walk
constexpr int kBcRows = 1024;         // batch rows per grid.y slice (4 per thread)
constexpr int kBcPitch = kBcRows + 4; // bytes: lane stride of 257 words -> phase 2 reads are bank-conflict-free
constexpr uint8_t kBcBig = 0xfe;  // hit byte of a long group (applied directly); 0xff = no hit; < kBcDeg = position

struct BcApply {
    const int32_t* group_node;   // [K] source node of each group
    const int32_t* edge_dst;     // [E] target node of each edge (the value SELECTED_ROAD receives)
    float* sel_links;            // [batch, n_links]
    float* sel_sources;          // [batch, n_nodes - n_links]
    int n_links, n_nodes;
    const float* prev_links;     // nullptr, or the SELECTED_ROAD arrays of the PREVIOUS step when this call writes a fresh
    const float* prev_sources;   // pair of buffers: a group without a hit then carries its previous value over
};

__global__ void __launch_bounds__(kThreads, 4) k_gd_sample_bcast(tarl_csr grp, const float* __restrict__ lg, float inv_t, int B,
                                                              int n_tiles, const float* __restrict__ u, int64_t u_sb,
                                                              int64_t u_sg, uint8_t* __restrict__ onehot,
                                                              float* __restrict__ part_lp, int32_t* __restrict__ part_bad,
                                                              BcApply ap, uint32_t seed_lo, uint32_t seed_hi,
                                                              const unsigned long long* __restrict__ seed_dev,
                                                              uint32_t draw_id, int row_offset) {
    __shared__ float sm_z[kBcGroups][kBcDeg], sm_cdf[kBcGroups][kBcDeg], sm_logp[kBcGroups][kBcDeg];
    __shared__ int sm_eid[kBcGroups][kBcDeg], sm_dst[kBcGroups][kBcDeg];
    __shared__ int sm_deg[kBcGroups], sm_k0[kBcGroups];
    __shared__ float sm_red[kThreads];
    __shared__ int sm_redi[kThreads];
    __shared__ __align__(16) uint8_t sm_hit[kBcGroups * kBcPitch];
    const bool apply = ap.sel_links != nullptr;
    const int row_base = blockIdx.y * kBcRows;
    const int rows_here = min(kBcRows, B - row_base);             // multiple of 4
    const int chunks = rows_here >> 2;
    int cp = 1;                                                   // chunks rounded up to a power of two (<= 256)
    while (cp < chunks) cp <<= 1;
    const int gpar = kThreads / cp;                               // groups handled side by side in phase 1
    const int c = threadIdx.x & (cp - 1), gsub = threadIdx.x / cp;
    const int row0 = row_base + 4 * c;
    const bool rows_live = c < chunks;
    // u == nullptr: the uniforms are drawn here — Philox4x32-10 keyed by the seed, counter (group, 4-row chunk): one
    // call = the four uniforms of this thread (a stream of its own, D4; torch.rand + its read back cost a third of
    // this kernel's traffic)
    const bool u_own = u == nullptr;
    if (u_own && seed_dev != nullptr) {
        const unsigned long long k = *seed_dev;
        seed_lo = (uint32_t)k; seed_hi = (uint32_t)(k >> 32);
    }
    const bool u_vec = !u_own && u_sb == 1 && (u_sg & 3) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0;
    float lp[4] = {0.f, 0.f, 0.f, 0.f};
    int bad[4] = {0, 0, 0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int g0 = tile * kBcGroups;
        __syncthreads();                                          // previous tile's phase 2 is done with shared memory
        // ---- phase 0: the 32 groups' distributions, once for every row
        {
            const int gl = threadIdx.x >> 3, j = threadIdx.x & 7, g = g0 + gl;
            int k0 = 0, deg = 0;
            if (g < grp.n_rows) { k0 = grp.ptr[g]; deg = grp.ptr[g + 1] - k0; }
            if (j == 0) { sm_deg[gl] = deg; sm_k0[gl] = k0; }
            if (j < deg && deg <= kBcDeg) {
                const int e = grp.eid[k0 + j];
                sm_eid[gl][j] = e;
                sm_z[gl][j] = lg[e] * inv_t;
                if (apply) sm_dst[gl][j] = ap.edge_dst[e];
            }
        }
        __syncthreads();
        if (threadIdx.x < kBcGroups) {
            const int gl = threadIdx.x, deg = sm_deg[gl];
            if (deg > 0 && deg <= kBcDeg) {
                float mx = -FLT_MAX;
                for (int j = 0; j < deg; ++j) mx = fmaxf(mx, sm_z[gl][j]);
                float den = 0.0f;
                for (int j = 0; j < deg; ++j) { const float ex = fexp(sm_z[gl][j] - mx); sm_z[gl][j] = ex; den += ex; }
                const float inv_den = fdiv(1.0f, den);
                float cum = 0.0f;
                for (int j = 0; j < deg; ++j) {
                    const float p = sm_z[gl][j] * inv_den;
                    cum += p;
                    sm_cdf[gl][j] = cum;
                    sm_logp[gl][j] = flog(p + kLogEps);
                }
            }
        }
        __syncthreads();
        // ---- phase 1: one uniform per (group, row) -> hit position, one-hot bytes, log-probability of the draw
        // the uniforms of the NEXT group are requested before this group's compares and stores (one load in flight
        // per thread would otherwise be all the memory parallelism this phase has)
        float4 u_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u_vec && rows_live && g0 + gsub < grp.n_rows)
            u_nxt = *reinterpret_cast<const float4*>(u + (int64_t)(g0 + gsub) * u_sg + row0);
        for (int gl = gsub; gl < kBcGroups; gl += gpar) {
            const int g = g0 + gl, deg = sm_deg[gl];
            const float4 t4 = u_nxt;
            if (u_vec && rows_live && gl + gpar < kBcGroups && g + gpar < grp.n_rows)
                u_nxt = *reinterpret_cast<const float4*>(u + (int64_t)(g + gpar) * u_sg + row0);
            if (!rows_live || g >= grp.n_rows || deg == 0) continue;
            float ug[4];
            if (u_own) {
                tarl::philox4x32_10((uint32_t)g, (uint32_t)((row_offset + row0) >> 2), draw_id, 0x53414d50u, seed_lo, seed_hi, ug);
            } else if (u_vec) {
                ug[0] = t4.x; ug[1] = t4.y; ug[2] = t4.z; ug[3] = t4.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) ug[q] = u[(int64_t)(row0 + q) * u_sb + (int64_t)g * u_sg];
            }
            int hit[4] = {-1, -1, -1, -1};
            if (deg <= kBcDeg) {
#pragma unroll
                for (int j = 0; j < kBcDeg; ++j) {
                    if (j < deg) {
                        const float cum = sm_cdf[gl][j];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (hit[q] < 0 && ug[q] < cum) hit[q] = j;
                    }
                }
#pragma unroll
                for (int j = 0; j < kBcDeg; ++j) {
                    if (j < deg) {
                        const uint32_t w = (hit[0] == j ? 1u : 0u) | (hit[1] == j ? 1u << 8 : 0u) |
                                           (hit[2] == j ? 1u << 16 : 0u) | (hit[3] == j ? 1u << 24 : 0u);
                        *reinterpret_cast<uint32_t*>(onehot + (int64_t)sm_eid[gl][j] * B + row0) = w;
                    }
                }
                if (part_lp != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (hit[q] < 0) bad[q] = 1;
                        else lp[q] += sm_logp[gl][hit[q]];
                    }
                }
                if (apply)
                    *reinterpret_cast<uint32_t*>(sm_hit + gl * kBcPitch + 4 * c) =
                        (uint32_t)(hit[0] & 0xff) | (uint32_t)(hit[1] & 0xff) << 8 | (uint32_t)(hit[2] & 0xff) << 16 |
                        (uint32_t)(hit[3] & 0xff) << 24;                      // -1 -> kBcNoHit
            } else {
                // a group longer than the shared-memory cache: every thread walks the logits row itself (same order)
                const int k0 = sm_k0[gl], k1 = k0 + deg;
                float mx = -FLT_MAX;
                for (int k = k0; k < k1; ++k) mx = fmaxf(mx, lg[grp.eid[k]] * inv_t);
                float den = 0.0f;
                for (int k = k0; k < k1; ++k) den += fexp(lg[grp.eid[k]] * inv_t - mx);
                const float inv_den = fdiv(1.0f, den);
                float cum = 0.0f, ph[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = k0; k < k1; ++k) {
                    const float p = fexp(lg[grp.eid[k]] * inv_t - mx) * inv_den;
                    cum += p;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (hit[q] < 0 && ug[q] < cum) { hit[q] = k - k0; ph[q] = p; }
                }
                for (int k = k0; k < k1; ++k) {
                    const int j = k - k0;
                    const uint32_t w = (hit[0] == j ? 1u : 0u) | (hit[1] == j ? 1u << 8 : 0u) |
                                       (hit[2] == j ? 1u << 16 : 0u) | (hit[3] == j ? 1u << 24 : 0u);
                    *reinterpret_cast<uint32_t*>(onehot + (int64_t)grp.eid[k] * B + row0) = w;
                }
                if (part_lp != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (hit[q] < 0) bad[q] = 1;
                        else lp[q] += flog(ph[q] + kLogEps);
                    }
                }
                if (apply) {
                    *reinterpret_cast<uint32_t*>(sm_hit + gl * kBcPitch + 4 * c) = 0x01010101u * kBcBig;
                    const int node = ap.group_node[g];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool lk = node < ap.n_links;
                        const int64_t at = lk ? (int64_t)(row0 + q) * ap.n_links + node
                                              : (int64_t)(row0 + q) * (ap.n_nodes - ap.n_links) + (node - ap.n_links);
                        if (hit[q] < 0) {
                            if (ap.prev_links != nullptr) { if (lk) ap.sel_links[at] = ap.prev_links[at]; else ap.sel_sources[at] = ap.prev_sources[at]; }
                            continue;
                        }
                        const float v = (float)ap.edge_dst[grp.eid[k0 + hit[q]]];
                        if (lk) ap.sel_links[at] = v; else ap.sel_sources[at] = v;
                    }
                }
            }
        }
        if (!apply) continue;
        __syncthreads();
        // ---- phase 2: SELECTED_ROAD with the node innermost (lanes = the tile's 32 groups, warps stride over the rows)
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int g = g0 + lane;
            const int deg = sm_deg[lane];
            const bool live = g < grp.n_rows && deg > 0 && deg <= kBcDeg;
            const int node = live ? ap.group_node[g] : 0;
            const bool is_link = node < ap.n_links;
            float* const base = is_link ? ap.sel_links + node : ap.sel_sources + (node - ap.n_links);
            const float* const prev = ap.prev_links == nullptr ? nullptr
                                      : (is_link ? ap.prev_links + node : ap.prev_sources + (node - ap.n_links));
            const int64_t pitch = is_link ? ap.n_links : ap.n_nodes - ap.n_links;
            for (int r = warp; r < rows_here; r += kThreads / 32) {
                if (!live) continue;
                const uint8_t h = sm_hit[lane * kBcPitch + r];
                if (h < kBcDeg) base[(int64_t)(row_base + r) * pitch] = (float)sm_dst[lane][h];
                else if (prev != nullptr && h != kBcBig) base[(int64_t)(row_base + r) * pitch] = prev[(int64_t)(row_base + r) * pitch];
            }
        }
    }
    if (part_lp != nullptr) {
        const int nb = gridDim.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __syncthreads();
            sm_red[threadIdx.x] = lp[q];
            sm_redi[threadIdx.x] = bad[q];
            __syncthreads();
            if (threadIdx.x < chunks) {                            // c == threadIdx.x, gsub == 0
                float v = 0.0f;
                int n = 0;
                for (int i = 0; i < gpar; ++i) { v += sm_red[i * cp + threadIdx.x]; n += sm_redi[i * cp + threadIdx.x]; }
                const int row = row_base + 4 * threadIdx.x + q;
                part_lp[(int64_t)row * nb + blockIdx.x] = v;
                part_bad[(int64_t)row * nb + blockIdx.x] = n;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_gd_backward_em4(tarl_csr grp, const float* __restrict__ logits, float inv_t,
                                                              int B, int C, int n_tiles,
                                                              const uint8_t* __restrict__ action,
                                                              const float* __restrict__ g_lp, const float* __restrict__ g_ent,
                                                              const float* __restrict__ log_prob, float* __restrict__ grad) {
  const int stride = gridDim.x;
  Stage4 nxt = stage4_bounds(grp, B, C, blockIdx.x, n_tiles);
  int eid_nxt[kCache];
  stage4_eids(grp, nxt, eid_nxt);
  Stage4 nxt2 = stage4_bounds(grp, B, C, blockIdx.x + stride, n_tiles);
  for (int tile = blockIdx.x; tile < n_tiles; tile += stride) {
    struct { int row0, k0, k1; } t = {nxt.row0, nxt.k0, nxt.k0 + nxt.deg};
    Soft4 s;
    s.deg = nxt.deg;
#pragma unroll
    for (int j = 0; j < kCache; ++j) s.eid[j] = eid_nxt[j];
    nxt = nxt2;
    stage4_eids(grp, nxt, eid_nxt);
    nxt2 = stage4_bounds(grp, B, C, tile + 2 * stride, n_tiles);
    if (s.deg == 0) continue;
    F4 wl, we;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b = t.row0 + q;
        wl.v[q] = (g_lp != nullptr && action != nullptr) ? g_lp[b] : 0.0f;
        if (log_prob != nullptr && log_prob[b] == -INFINITY) wl.v[q] = 0.0f;
        we.v[q] = (g_ent != nullptr) ? g_ent[b] : 0.0f;
    }
    soft4_load(s, grp, logits, action, B, t.row0, t.k0, t.k1, inv_t);
    F4 Sa = {{0.f, 0.f, 0.f, 0.f}}, Sc = {{0.f, 0.f, 0.f, 0.f}};
    // pass 1: the two group sums; pass 2 recomputes r_e and c_e instead of keeping them (48 registers less: occupancy)
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg) {
            const F4 a = action4_of(s.aw[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float p = s.ex[j].v[q] * s.inv_den.v[q];
                const float r = fdiv(p, p + kLogEps);
                Sa.v[q] += a.v[q] * r;
                Sc.v[q] += p * (flog(p + kLogEps) + r);
            }
        }
    }
    for (int k = t.k0 + kCache; k < t.k1; ++k) {
        const int e = grp.eid[k];
        const F4 p4 = tail4_p(s, logits, B, t.row0, e, inv_t);
        F4 a = {{0.f, 0.f, 0.f, 0.f}};
        if (action != nullptr) a = action4(action, B, t.row0, e);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r = fdiv(p4.v[q], p4.v[q] + kLogEps);
            Sa.v[q] += a.v[q] * r;
            Sc.v[q] += p4.v[q] * (flog(p4.v[q] + kLogEps) + r);
        }
    }
#pragma unroll
    for (int j = 0; j < kCache; ++j) {
        if (j < s.deg) {
            const F4 a = action4_of(s.aw[j]);
            F4 out;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float p = s.ex[j].v[q] * s.inv_den.v[q];
                const float r = fdiv(p, p + kLogEps);
                const float c = flog(p + kLogEps) + r;
                out.v[q] = (-we.v[q] * p * (c - Sc.v[q]) + wl.v[q] * (a.v[q] * r - p * Sa.v[q])) * inv_t;
            }
            st4(grad + (int64_t)s.eid[j] * B + t.row0, out);
        }
    }
    for (int k = t.k0 + kCache; k < t.k1; ++k) {
        const int e = grp.eid[k];
        const F4 p4 = tail4_p(s, logits, B, t.row0, e, inv_t);
        F4 a = {{0.f, 0.f, 0.f, 0.f}};
        if (action != nullptr) a = action4(action, B, t.row0, e);
        F4 out;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r = fdiv(p4.v[q], p4.v[q] + kLogEps);
            const float c = flog(p4.v[q] + kLogEps) + r;
            out.v[q] = (-we.v[q] * p4.v[q] * (c - Sc.v[q]) + wl.v[q] * (a.v[q] * r - p4.v[q] * Sa.v[q])) * inv_t;
        }
        st4(grad + (int64_t)e * B + t.row0, out);
    }
  }
}

// The fast path applies when the tensors are edge-major, 16-byte aligned, B is 4, 8, 16 or a multiple of 32, and the
// action (if any) is edge-major bytes.
inline bool em4_ok(const tarl_rows* lg, int B, const tarl_rows* act, int act_dtype, const tarl_rows* extra = nullptr) {
    if (!(B == 4 || B == 8 || B == 16 || (B % 32) == 0)) return false;
    auto em = [&](const tarl_rows* r, size_t elem) {
        return r->row_stride == 1 && r->col_stride == B && (reinterpret_cast<uintptr_t>(r->data) % (elem * 4)) == 0;
    };
    if (lg == nullptr || !em(lg, 4)) return false;
    if (extra != nullptr && !em(extra, 4)) return false;
    if (act != nullptr && (act_dtype != TARL_ACTION_U8 || !em(act, 1))) return false;
    return true;
}
inline int em4_chunks(int B) { return (B >= 32 ? 32 : B) >> 2; }
inline dim3 em4_grid(int K, int B) { return dim3(blocks_for((int64_t)K * em4_chunks(B)), (B + 31) / 32); }

inline int pow2_rows(int B) {
    int p = 1;
    while (p < B && p < 32) p <<= 1;
    return p;
}
inline dim3 gd_grid(int K, int B) {      // one CTA per tile of kThreads/Bp groups
    const int Bp = pow2_rows(B);
    return dim3(blocks_for((int64_t)K * Bp), (B + 31) / 32);
}
constexpr int kBwdMaxCtas = 148 * 8;      // persistent backward (tile stride): a few waves of resident CTAs
constexpr int kFwdMaxCtas = 148 * 16;     // forward CTAs walk tiles with a stride: this bounds the partials per row
// The caps count CTAs of the WHOLE grid: with many row chunks (grid.y = B / 32 = 32 for a 1024-replica rollout) each
// chunk gets cap / grid.y CTAs in x, so that a CTA still walks tens of tiles — the prefetch pipeline has something to
// prefetch and the per-CTA block reductions of the partial sums are amortised.
inline unsigned em4_cap(int cap, const dim3& g) {
    const unsigned per = (unsigned)cap / (g.y > 0 ? g.y : 1);
    return per > 0 ? per : 1;
}
inline dim3 gd_fwd_grid(int K, int B) {
    dim3 g = gd_grid(K, B);
    if (g.x > (unsigned)kFwdMaxCtas) g.x = kFwdMaxCtas;
    return g;
}

int check_csr(const tarl_csr* c) {
    if (c == nullptr || c->n_rows < 0 || c->n_edges < 0) return TARL_E_BADARG;
    if (c->n_rows > 0 && c->ptr == nullptr) return TARL_E_BADARG;
    if (c->n_edges > 0 && c->eid == nullptr) return TARL_E_BADARG;
    return TARL_OK;
}

inline View view_of(const tarl_rows* r) { return r != nullptr ? View{r->row_stride, r->col_stride} : View{0, 0}; }
template <typename T>
inline T* data_of(const tarl_rows* r) { return r != nullptr ? static_cast<T*>(r->data) : nullptr; }


// launch of k_gd_sample_bcast (+ the log-probability finish); apply.sel_links == nullptr: sample only
int launch_sample_bcast(const tarl_csr* groups, const float* logits_row, float temperature, int batch,
                        const tarl_rows* uniforms, uint8_t* onehot, float* log_prob, float* partials, const BcApply& apply,
                        uint64_t seed, const uint64_t* seed_dev, uint32_t draw_id, int row_offset, cudaStream_t s) {
    if (log_prob != nullptr && partials == nullptr) return TARL_E_WORKSPACE;
    const int n_tiles = (groups->n_rows + kBcGroups - 1) / kBcGroups;
    int nb = tarl_graphdist_partial_count(groups->n_rows, batch);           // bounds the partials per row
    if (nb > n_tiles) nb = n_tiles;
    if (nb < 1) nb = 1;
    const dim3 grid((unsigned)nb, (unsigned)((batch + kBcRows - 1) / kBcRows));
    float* part_lp = log_prob ? partials + (size_t)batch * nb : nullptr;     // same layout as the forward's
    int32_t* part_bad = log_prob ? reinterpret_cast<int32_t*>(partials + 2 * (size_t)batch * nb) : nullptr;
    k_gd_sample_bcast<<<grid, kThreads, 0, s>>>(*groups, logits_row, 1.0f / temperature, batch, n_tiles,
                                                uniforms ? data_of<const float>(uniforms) : nullptr,
                                                uniforms ? uniforms->row_stride : 0, uniforms ? uniforms->col_stride : 0,
                                                onehot, part_lp, part_bad, apply, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                reinterpret_cast<const unsigned long long*>(seed_dev), draw_id, row_offset);
    if (log_prob != nullptr) k_gd_finish<<<batch, kThreads, 0, s>>>(nullptr, part_lp, part_bad, nb, nullptr, log_prob);
    return launch_status();
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
    k_policy_node<<<tarl::tile_grid(n_nodes, batch), tarl::kTileThreads, 0, s>>>(
        emb_weight, emb_rows, node_features, nf_batch_stride, nf_row_stride, road_index_col, batch,
        tarl::tile_rows_pow2(batch), n_nodes, node_emb, node_idx, flags);
    if (n_edges > 0)
        k_policy_edge<<<blocks_for((int64_t)n_edges * ((batch & 3) == 0 ? batch >> 2 : batch)), kThreads, 0, s>>>(
            node_emb, edge_dst, batch, n_edges, logits);
    return launch_status();
}

int tarl_policy_embed_backward(const tarl_csr* by_target, const tarl_rows* grad_logits, const int32_t* node_idx,
                               int32_t batch, float* node_grad, float* grad_weight, int32_t emb_rows, void* stream) {
    int rc = check_csr(by_target);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || emb_rows <= 0 || !grad_weight) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)emb_rows, s) != cudaSuccess) return TARL_E_LAUNCH;
    if (batch == 0 || by_target->n_rows == 0) return TARL_OK;
    if (!node_idx || !node_grad || (by_target->n_edges > 0 && (!grad_logits || !grad_logits->data))) return TARL_E_BADARG;
    const View gv = view_of(grad_logits);
    const bool vec = gv.sb == 1 && gv.se == batch && (batch & 3) == 0;
    k_policy_node_grad<<<blocks_for((int64_t)by_target->n_rows * (vec ? batch >> 2 : batch)), kThreads, 0, s>>>(
        *by_target, data_of<const float>(grad_logits), gv.sb, gv.se, batch, node_grad);
    int Bp = 1;
    while (Bp < batch && Bp < 32) Bp <<= 1;
    k_policy_weight_grad<<<blocks_for((int64_t)by_target->n_rows * Bp), kThreads, 0, s>>>(node_grad, node_idx, batch, Bp,
                                                                                          by_target->n_rows, grad_weight);
    return launch_status();
}

int32_t tarl_graphdist_partial_count(int32_t n_groups, int32_t batch) {
    if (n_groups <= 0 || batch <= 0) return 0;        // an upper bound over both forward variants
    const unsigned a = gd_fwd_grid(n_groups, batch).x, b = em4_grid(n_groups, batch).x;
    const unsigned m = a > b ? a : b;
    return (int32_t)(m > (unsigned)kFwdMaxCtas ? kFwdMaxCtas : m);
}

int tarl_graphdist_forward(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                           const tarl_rows* action, int32_t action_dtype, const tarl_rows* proba, const tarl_rows* mode,
                           float* entropy, float* log_prob, float* partials, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (action != nullptr && (action_dtype < 0 || action_dtype > 2))) return TARL_E_BADARG;
    if (log_prob != nullptr && action == nullptr) return TARL_E_BADARG;
    if (batch == 0) return TARL_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int E = groups->n_edges, K = groups->n_rows;
    if (E > 0 && (logits == nullptr || logits->data == nullptr)) return TARL_E_BADARG;
    const bool fast = proba == nullptr && mode == nullptr && temperature != 0.0f && em4_ok(logits, batch, action, action_dtype);
    dim3 grid = K > 0 ? gd_fwd_grid(K, batch) : dim3(0, 1);
    int n_tiles = K > 0 ? (int)gd_grid(K, batch).x : 0;
    if (fast && K > 0) {
        grid = em4_grid(K, batch);
        n_tiles = (int)grid.x;
        if (grid.x > em4_cap(kFwdMaxCtas, grid)) grid.x = em4_cap(kFwdMaxCtas, grid);
    }
    const int nb = (int)grid.x;
    float* part_ent = nullptr; float* part_lp = nullptr; int32_t* part_bad = nullptr;
    if (entropy != nullptr || log_prob != nullptr) {
        if (partials == nullptr && nb > 0) return TARL_E_WORKSPACE;
        part_ent = entropy ? partials : nullptr;
        part_lp = log_prob ? partials + (size_t)batch * nb : nullptr;
        part_bad = log_prob ? reinterpret_cast<int32_t*>(partials + 2 * (size_t)batch * nb) : nullptr;
    }
    if (nb > 0 && fast)
        k_gd_forward_em4<<<grid, kThreads, 0, s>>>(*groups, data_of<const float>(logits), 1.0f / temperature, batch,
                                                   em4_chunks(batch), n_tiles, data_of<const uint8_t>(action), part_ent,
                                                   part_lp, part_bad);
    else if (nb > 0)
        k_gd_forward<<<grid, kThreads, 0, s>>>(*groups, data_of<const float>(logits), view_of(logits), temperature, batch,
                                               pow2_rows(batch), n_tiles, data_of<const void>(action), view_of(action),
                                               action_dtype,
                                               data_of<float>(proba), view_of(proba), data_of<float>(mode), view_of(mode),
                                               part_ent, part_lp, part_bad);
    if (entropy != nullptr || log_prob != nullptr)
        k_gd_finish<<<batch, kThreads, 0, s>>>(part_ent, part_lp, part_bad, nb, entropy, log_prob);
    return launch_status();
}

int tarl_graphdist_backward(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                            const tarl_rows* action, int32_t action_dtype, const float* grad_log_prob,
                            const float* grad_entropy, const float* log_prob, const tarl_rows* grad_logits, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (action != nullptr && (action_dtype < 0 || action_dtype > 2))) return TARL_E_BADARG;
    if (batch == 0 || groups->n_edges == 0) return TARL_OK;
    if (!logits || !logits->data || !grad_logits || !grad_logits->data) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // every edge has a source, hence a group: every grad entry is written
    if (temperature != 0.0f && em4_ok(logits, batch, action, action_dtype, grad_logits)) {
        dim3 grid = em4_grid(groups->n_rows, batch);
        const int n_tiles = (int)grid.x;
        if (grid.x > em4_cap(kBwdMaxCtas, grid)) grid.x = em4_cap(kBwdMaxCtas, grid);
        k_gd_backward_em4<<<grid, kThreads, 0, s>>>(
            *groups, data_of<const float>(logits), 1.0f / temperature, batch, em4_chunks(batch), n_tiles,
            data_of<const uint8_t>(action), grad_log_prob, grad_entropy, log_prob, data_of<float>(grad_logits));
        return launch_status();
    }
    k_gd_backward<<<gd_grid(groups->n_rows, batch), kThreads, 0, s>>>(
        *groups, data_of<const float>(logits), view_of(logits), temperature, batch, pow2_rows(batch),
        data_of<const void>(action), view_of(action), action_dtype, grad_log_prob, grad_entropy, log_prob,
        data_of<float>(grad_logits), view_of(grad_logits));
    return launch_status();
}

int tarl_graphdist_sample(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                          const tarl_rows* uniforms, const tarl_rows* onehot, int32_t onehot_dtype, float* log_prob,
                          float* partials, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (onehot_dtype != TARL_ACTION_U8 && onehot_dtype != TARL_ACTION_I64)) return TARL_E_BADARG;
    if (batch == 0 || groups->n_edges == 0) return TARL_OK;
    if (!logits || !logits->data || !uniforms || !uniforms->data || !onehot || !onehot->data) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // logits: edge-major, or one row broadcast over the batch (row stride 0, unit column stride)
    const bool bcast = batch > 1 && logits->row_stride == 0 && logits->col_stride == 1;
    tarl_rows lg_as_em = *logits;
    if (bcast) { lg_as_em.data = reinterpret_cast<void*>(uintptr_t(16)); lg_as_em.row_stride = 1; lg_as_em.col_stride = batch; }   // passes the layout test
    if (temperature != 0.0f && onehot_dtype == TARL_ACTION_U8 &&
        em4_ok(bcast ? &lg_as_em : logits, batch, onehot, onehot_dtype)) {
        dim3 grid = em4_grid(groups->n_rows, batch);
        const int n_tiles = (int)grid.x;
        if (grid.x > em4_cap(kFwdMaxCtas, grid)) grid.x = em4_cap(kFwdMaxCtas, grid);
        const int nb = (int)grid.x;
        if (log_prob != nullptr && partials == nullptr) return TARL_E_WORKSPACE;
        float* part_lp = log_prob ? partials + (size_t)batch * nb : nullptr;       // same layout as the forward's
        int32_t* part_bad = log_prob ? reinterpret_cast<int32_t*>(partials + 2 * (size_t)batch * nb) : nullptr;
        if (bcast)      // one logits row for every batch row: the group's distribution is computed once per CTA
            return launch_sample_bcast(groups, data_of<const float>(logits), temperature, batch, uniforms,
                                       data_of<uint8_t>(onehot), log_prob, partials, BcApply{}, 0, nullptr, 0u, 0, s);
        auto kernel = k_gd_sample_em4<false>;
        kernel<<<grid, kThreads, 0, s>>>(*groups, data_of<const float>(logits), 1.0f / temperature, batch,
                                         em4_chunks(batch), n_tiles, data_of<const float>(uniforms), uniforms->row_stride,
                                         uniforms->col_stride, data_of<uint8_t>(onehot), part_lp, part_bad);
        if (log_prob != nullptr) k_gd_finish<<<batch, kThreads, 0, s>>>(nullptr, part_lp, part_bad, nb, nullptr, log_prob);
        return launch_status();
    }
    if (log_prob != nullptr) return TARL_E_BADARG;       // the fused log-probability exists on the fast path only
    const dim3 grid = gd_grid(groups->n_rows, batch);
    if (onehot_dtype == TARL_ACTION_U8)
        k_gd_sample<uint8_t><<<grid, kThreads, 0, s>>>(*groups, data_of<const float>(logits), view_of(logits), temperature,
                                                       batch, pow2_rows(batch), data_of<const float>(uniforms),
                                                       view_of(uniforms), data_of<uint8_t>(onehot), view_of(onehot));
    else
        k_gd_sample<long long><<<grid, kThreads, 0, s>>>(*groups, data_of<const float>(logits), view_of(logits),
                                                         temperature, batch, pow2_rows(batch),
                                                         data_of<const float>(uniforms), view_of(uniforms),
                                                         data_of<long long>(onehot), view_of(onehot));
    return launch_status();
}

int tarl_graphdist_sample_apply(const tarl_csr* groups, const float* logits_row, float temperature, int32_t batch,
                                const tarl_rows* uniforms, uint8_t* onehot, float* log_prob, float* partials,
                                const int32_t* group_node, const int32_t* edge_dst, float* sel_links, float* sel_sources,
                                const float* prev_links, const float* prev_sources,
                                int32_t n_links, int32_t n_nodes, uint64_t seed, const uint64_t* seed_dev,
                                uint32_t draw_id, int32_t row_offset, void* stream) {
    int rc = check_csr(groups);
    if (rc != TARL_OK) return rc;
    if (batch < 0 || (batch & 3) != 0 || temperature == 0.0f || n_links < 0 || n_nodes < n_links) return TARL_E_BADARG;
    if (row_offset < 0 || (row_offset & 3) != 0) return TARL_E_BADARG;
    if (batch == 0 || groups->n_edges == 0) return TARL_OK;
    if (uniforms != nullptr && uniforms->data == nullptr) uniforms = nullptr;      // no uniforms: drawn in the kernel
    if (!logits_row || !onehot || (reinterpret_cast<uintptr_t>(onehot) & 3) != 0 ||
        !group_node || !edge_dst || !sel_links || (n_nodes > n_links && !sel_sources))
        return TARL_E_BADARG;
    if (prev_links != nullptr && n_nodes > n_links && prev_sources == nullptr) return TARL_E_BADARG;
    const BcApply ap = {group_node, edge_dst, sel_links, sel_sources, n_links, n_nodes, prev_links, prev_sources};
    return launch_sample_bcast(groups, logits_row, temperature, batch, uniforms, onehot, log_prob, partials, ap, seed,
                               seed_dev, draw_id, row_offset, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// value_net.cu — MPNNValueNet's message / mean-aggregate / update, forward and backward (sm_100a, fp32).
//
// Reference semantics: /root/reference/src/agents/mpnn_agent.py:267-402 with dropout off (eval mode):
//   x[b,n]   = [node_features[b,n,0:7] ‖ agent_features[agent_index[b,n], 0:9]]                     (:330-349)
//   msg[b,e] = tanh(w · [x[b, edge_index[1][e]] ‖ edge_features[b,e]] + w0)   flow target_to_source  (:385-386)
//   mean[b,n]= mean of msg over the edges whose edge_index[0] == n (0 for nodes without one)         (aggr='mean')
//   v[b,n]   = tanh(a * mean + c)                                                                    (:402)
// The 16 node terms of the message do not depend on the edge, so they are projected ONCE per node
// (proj[b,n] = w[0:16] · x[b,n]) instead of being gathered as [E,16] rows per edge as PyG does: per edge the kernels
// read one 4-byte projected scalar + the edge feature. Every reduction has a fixed order (segment sums in ascending
// edge id, per-block partials summed by a second kernel), so results are run-to-run deterministic.
// Layout: proj / mean / v / gm are NODE-major with the batch row innermost (element (b, n) at n*B + b) and one thread
// handles one (node, row) pair with the row innermost, so the B lanes of a node gather one contiguous B-vector per
// neighbour and read the node's edge list once. edge_features may be broadcast over the batch (batch stride 0).
//
// Train mode (nn.Dropout(0.05) on the [B*E, 17] message input, :278,385-386): every (row, edge) pair has its own
// 17-bit keep mask, so the per-node projection no longer factors out. The *_dropout entry points compute the message
// per (TARGET node, row) — the node's 16 inputs are loaded once and every in-edge applies its own mask — into an
// edge-major message buffer msg[e*B + b] that the source-side mean and the backward pass read back. The mask is
// either injected ([B, E] words, bit k = input k survives; parity tests replay the mask the reference drew) or drawn
// in the kernel: Philox4x32-10 keyed by the seed, counter (edge, row, draw), 12-bit fields compared with
// round(p * 4096) (two draws per pair; the forward pass stores the words edge-major and the backward pass reads them back) — a documented stream of its own (declared divergence D4, as for the core step's noise).
#include <cuda_runtime.h>
#include <stdint.h>

#include "tarl_b200.h"
#include "tile_map.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kNodeDim = 7, kAgentDim = 9, kIn = 16;
constexpr int kGrads = 20;   // w[0:17], w0, a, c

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

struct Inputs {
    const float* nf; int64_t nf_bs, nf_rs;     // node_features [B,N,>=7]
    const float* ef; int64_t ef_bs;            // edge_features [B,E], batch stride in elements (0 = shared by all rows)
    const long long* ai;                       // agent_index [B,N]
    const float* af; int af_rows;              // agent_features [rows,9]
    int B, N, E;
};

__device__ __forceinline__ void load_x(const Inputs& in, int b, int n, float x[kIn], int32_t* flags) {
    const float* p = in.nf + b * in.nf_bs + n * in.nf_rs;
#pragma unroll
    for (int c = 0; c < kNodeDim; ++c) x[c] = p[c];
    long long a = in.ai[(int64_t)b * in.N + n];
    if (a < 0) a += in.af_rows;                                   // torch advanced indexing wraps negatives
    if (a < 0 || a >= in.af_rows) {
        if (flags != nullptr) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_AGENT_RANGE);
        a = 0;
    }
    const float* q = in.af + a * kAgentDim;
#pragma unroll
    for (int c = 0; c < kAgentDim; ++c) x[kNodeDim + c] = q[c];
}

// proj[n,b] = w[0:16] . x[b,n]. Tiled (tile_map.cuh): x is assembled from the row-major observation with the node
// innermost (node_features rows and agent_index entries of 32 consecutive nodes are contiguous; agent_features is a
// small table), proj is written with the row innermost.
// The 7 node features of 32 consecutive nodes of one sample are 224 contiguous floats when the observation is dense
// ([B, N, 7] with unit strides): the warp loads them coalesced (7 instructions, 7 wavefronts) into its own shared
// buffer and every lane picks its 7 (stride 7: conflict-free), instead of 7 strided loads of 28 sectors each.
// n_first = first node of the warp's segment; lanes beyond the graph read nothing. Warp-synchronous.
constexpr int kWarpsPerCta = tarl::kTileThreads / 32;
__device__ __forceinline__ void load_nf_staged(const Inputs& in, bool dense, int b, int n_first, int lane, bool lane_live,
                                               float* __restrict__ warp_buf, float x[kNodeDim]) {
    if (dense) {
        const float* p = in.nf + b * in.nf_bs + (int64_t)n_first * kNodeDim;
        const int n_valid = min(32, in.N - n_first) * kNodeDim;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kNodeDim; ++k) {
            const int o = lane + 32 * k;
            if (o < n_valid) warp_buf[o] = p[o];
        }
        __syncwarp();
        if (lane_live) {
#pragma unroll
            for (int c = 0; c < kNodeDim; ++c) x[c] = warp_buf[lane * kNodeDim + c];
        }
    } else if (lane_live) {
        const float* p = in.nf + b * in.nf_bs + (int64_t)(n_first + lane) * in.nf_rs;
#pragma unroll
        for (int c = 0; c < kNodeDim; ++c) x[c] = p[c];
    }
}

__device__ __forceinline__ long long agent_row(const Inputs& in, int b, int n, int32_t* flags) {
    long long a = in.ai[(int64_t)b * in.N + n];
    if (a < 0) a += in.af_rows;                                   // torch advanced indexing wraps negatives
    if (a < 0 || a >= in.af_rows) {
        if (flags != nullptr) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_AGENT_RANGE);
        a = 0;
    }
    return a;
}

// pa[a] = w[7:16] . agent_features[a, 0:9]: the agent part of the message projection, once per agent row instead of
// once per (node, sample) — the projection then gathers one float per node instead of nine.
__global__ void __launch_bounds__(256) k_value_agent_project(const float* __restrict__ af, int af_rows,
                                                             const float* __restrict__ w, float* __restrict__ pa) {
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= af_rows) return;
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < kAgentDim; ++c) acc += w[kNodeDim + c] * af[(int64_t)a * kAgentDim + c];
    pa[a] = acc;
}

__global__ void __launch_bounds__(tarl::kTileThreads) k_value_project(Inputs in, int Bp, const float* __restrict__ w,
                                                                      const float* __restrict__ pa,
                                                                      float* __restrict__ proj,
                                                                      int32_t* __restrict__ flags) {
    __shared__ float sm[tarl::kTileSmem];
    __shared__ float sm_nf[kWarpsPerCta][32 * kNodeDim];
    const tarl::Tile t = tarl::tile_here(in.B, Bp);
    const bool dense = in.nf_rs == kNodeDim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float wr[kNodeDim];
#pragma unroll
    for (int c = 0; c < kNodeDim; ++c) wr[c] = w[c];
    tarl::tile_walk_nodes(t, [&](int r, int j) {
        const int n = t.n0 + j;
        const bool live = n < in.N && r < t.nrows;
        const bool row_live = r < t.nrows && t.n0 + (j - lane) < in.N;      // warp-uniform
        float x[kNodeDim];
        if (row_live) load_nf_staged(in, dense, t.b0 + r, n - lane, lane, live, sm_nf[warp], x);
        if (!live) return;
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < kNodeDim; ++c) acc += wr[c] * x[c];
        sm[tarl::tile_slot(t, r, j)] = acc + pa[agent_row(in, t.b0 + r, n, flags)];
    });
    __syncthreads();
    tarl::tile_walk_rows(t, [&](int r, int j) {
        const int n = t.n0 + j;
        if (n >= in.N || r >= t.nrows) return;
        proj[(int64_t)n * in.B + t.b0 + r] = sm[tarl::tile_slot(t, r, j)];
    });
}

// one thread per (source node, batch row): segment mean of tanh messages in ascending edge id, then the node update
__global__ void __launch_bounds__(kThreads) k_value_aggregate(tarl_csr by_src, Inputs in, const float* __restrict__ w,
                                                              const float* __restrict__ w0, const float* __restrict__ a,
                                                              const float* __restrict__ c, const float* __restrict__ proj,
                                                              float* __restrict__ mean, float* __restrict__ v) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)in.B * in.N) return;
    const int n = (int)(i / in.B), b = (int)(i % in.B);
    const float we = w[kIn], bias = w0[0];
    const float* ef = in.ef + b * in.ef_bs;
    const int k0 = by_src.ptr[n], k1 = by_src.ptr[n + 1];
    float acc = 0.0f;
    for (int k = k0; k < k1; ++k) acc += tanhf(proj[(int64_t)by_src.idx[k] * in.B + b] + we * ef[by_src.eid[k]] + bias);
    const float m = k1 > k0 ? acc / (float)(k1 - k0) : 0.0f;
    mean[i] = m;
    v[i] = tanhf(a[0] * m + c[0]);
}

// The same with one thread per (source node, 4 consecutive batch rows) when B % 4 == 0: 128-bit gathers of the
// projected B-vectors and 128-bit stores, and the node's edge list / edge features are read once per four rows.
__global__ void __launch_bounds__(kThreads) k_value_aggregate4(tarl_csr by_src, Inputs in, const float* __restrict__ w,
                                                               const float* __restrict__ w0, const float* __restrict__ a,
                                                               const float* __restrict__ c, const float* __restrict__ proj,
                                                               float* __restrict__ mean, float* __restrict__ v) {
    const int C = in.B >> 2;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)C * in.N) return;
    const int n = (int)(i / C), b0 = 4 * (int)(i % C);
    const float we = w[kIn], bias = w0[0];
    const int k0 = by_src.ptr[n], k1 = by_src.ptr[n + 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = k0; k < k1; ++k) {
        const float4 p = *reinterpret_cast<const float4*>(proj + (int64_t)by_src.idx[k] * in.B + b0);
        const int e = by_src.eid[k];
        float f[4];
        f[0] = in.ef[(int64_t)b0 * in.ef_bs + e];
        if (in.ef_bs == 0) { f[1] = f[2] = f[3] = f[0]; }
        else {
#pragma unroll
            for (int q = 1; q < 4; ++q) f[q] = in.ef[(int64_t)(b0 + q) * in.ef_bs + e];
        }
        acc[0] += tanhf(p.x + we * f[0] + bias);
        acc[1] += tanhf(p.y + we * f[1] + bias);
        acc[2] += tanhf(p.z + we * f[2] + bias);
        acc[3] += tanhf(p.w + we * f[3] + bias);
    }
    float m[4], vv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        m[q] = k1 > k0 ? acc[q] / (float)(k1 - k0) : 0.0f;
        vv[q] = tanhf(a[0] * m[q] + c[0]);
    }
    *reinterpret_cast<float4*>(mean + (int64_t)n * in.B + b0) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(v + (int64_t)n * in.B + b0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

__device__ __forceinline__ float warp_sum(float x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// sums `count` per-thread values over the block (fixed tree) and lets thread 0 write them to dst[0..count)
template <int kCount>
__device__ __forceinline__ void block_store(float (&vals)[kCount], float* __restrict__ dst) {
    __shared__ float sm[kThreads / 32][kCount];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kCount; ++j) {
        const float s = warp_sum(vals[j]);
        if (lane == 0) sm[wid][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < kCount) {
        float s = 0.0f;
        for (int wi = 0; wi < kThreads / 32; ++wi) s += sm[wi][threadIdx.x];
        dst[threadIdx.x] = s;
    }
}

// node update backward: dv = g_v (1 - v^2); partial sums of d a, d c; gm = dv * a / deg handed to the edge pass.
// grad_v: element (b, n) at b*gv_sb + n*gv_sn. Same tiles as the edge pass (one partial row per tile).
__global__ void __launch_bounds__(tarl::kTileThreads) k_value_node_grad(tarl_csr by_src, int B, int Bp, int N,
                                                                        const float* __restrict__ a,
                                                                        const float* __restrict__ mean,
                                                                        const float* __restrict__ v,
                                                                        const float* __restrict__ gv, int64_t gv_sb,
                                                                        int64_t gv_sn, const float* __restrict__ head_g,
                                                                        const float* __restrict__ head_w,
                                                                        float* __restrict__ gm,
                                                                        float* __restrict__ partials) {
    __shared__ float sm[tarl::kTileSmem];
    const tarl::Tile t = tarl::tile_here(B, Bp);
    float vals[2] = {0.0f, 0.0f};
    const float a0 = a[0];
    // grad_v normally arrives row-major [B, N] (autograd of a dense head): staged with the node innermost. With the
    // value head applied by tarl_value_head_forward it is the rank-1 product head_g[b] * head_w[n] and never exists.
    tarl::tile_walk_nodes(t, [&](int r, int j) {
        const int n = t.n0 + j;
        if (n >= N || r >= t.nrows) return;
        sm[tarl::tile_slot(t, r, j)] = head_g != nullptr ? head_g[t.b0 + r] * head_w[n] : gv[(t.b0 + r) * gv_sb + n * gv_sn];
    });
    __syncthreads();
    tarl::tile_walk_rows(t, [&](int r, int j) {
        const int n = t.n0 + j, b = t.b0 + r;
        if (n >= N || r >= t.nrows) return;
        const int64_t i = (int64_t)n * B + b;
        const float vv = v[i];
        const float dv = sm[tarl::tile_slot(t, r, j)] * (1.0f - vv * vv);
        vals[0] += dv * mean[i];
        vals[1] += dv;
        const int deg = by_src.ptr[n + 1] - by_src.ptr[n];
        gm[i] = deg > 0 ? dv * a0 / (float)deg : 0.0f;
    });
    block_store<2>(vals, partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kGrads + 18);
}

// message backward over (TARGET node, batch row) tiles. Walk 1, rows innermost: every in-edge's message is recomputed
// from the node's own projection and the gathered gm B-vectors -> gs[n,b] = sum of d z over the in-edges (ascending
// edge id), kept in shared memory. Walk 2, nodes innermost: the 16 input-weight gradients gs * x[b,n,:] with x read
// the way the observation is laid out (once per node and row, not once per edge).
__global__ void __launch_bounds__(tarl::kTileThreads) k_value_edge_grad(tarl_csr by_dst, Inputs in, int Bp, bool vec4,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ w0,
                                                                        const float* __restrict__ proj,
                                                                        const float* __restrict__ gm,
                                                                        float* __restrict__ partials) {
    __shared__ float sm[tarl::kTileSmem];
    const tarl::Tile t = tarl::tile_here(in.B, Bp);
    float vals[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) vals[j] = 0.0f;
    const float we = w[kIn], bias = w0[0];
    if (vec4) {
        // (target node, 4 consecutive rows) per thread: 128-bit loads of proj / gm, the in-edge list once per four rows
        const int C = t.Bp >> 2, shc = t.sh - 2;
        for (int p = threadIdx.x; p < (tarl::kTilePairs >> 2); p += tarl::kTileThreads) {
            const int r0 = 4 * (p & (C - 1)), j = p >> shc;
            const int n = t.n0 + j, b0 = t.b0 + r0;
            if (n >= in.N || r0 >= t.nrows) continue;
            const float4 pn = *reinterpret_cast<const float4*>(proj + (int64_t)n * in.B + b0);
            float gs[4] = {0.f, 0.f, 0.f, 0.f}, gwe = 0.0f;
            const int k1 = by_dst.ptr[n + 1];
            for (int k = by_dst.ptr[n]; k < k1; ++k) {
                const int e = by_dst.eid[k];
                const float4 g4 = *reinterpret_cast<const float4*>(gm + (int64_t)by_dst.idx[k] * in.B + b0);
                float f[4];
                f[0] = in.ef[(int64_t)b0 * in.ef_bs + e];
                if (in.ef_bs == 0) { f[1] = f[2] = f[3] = f[0]; }
                else {
#pragma unroll
                    for (int q = 1; q < 4; ++q) f[q] = in.ef[(int64_t)(b0 + q) * in.ef_bs + e];
                }
                const float pq[4] = {pn.x, pn.y, pn.z, pn.w}, gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float m = tanhf(pq[q] + we * f[q] + bias);
                    const float gz = gq[q] * (1.0f - m * m);
                    gs[q] += gz;
                    gwe += gz * f[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sm[tarl::tile_slot(t, r0 + q, j)] = gs[q];
                vals[17] += gs[q];
            }
            vals[16] += gwe;
        }
    } else {
        tarl::tile_walk_rows(t, [&](int r, int j) {
            const int n = t.n0 + j, b = t.b0 + r;
            if (n >= in.N || r >= t.nrows) return;
            const float pn = proj[(int64_t)n * in.B + b];
            const float* ef = in.ef + b * in.ef_bs;
            float gs = 0.0f, gwe = 0.0f;
            const int k1 = by_dst.ptr[n + 1];
            for (int k = by_dst.ptr[n]; k < k1; ++k) {
                const float f = ef[by_dst.eid[k]];
                const float m = tanhf(pn + we * f + bias);
                const float gz = gm[(int64_t)by_dst.idx[k] * in.B + b] * (1.0f - m * m);
                gs += gz;
                gwe += gz * f;
            }
            sm[tarl::tile_slot(t, r, j)] = gs;
            vals[16] += gwe;
            vals[17] += gs;
        });
    }
    __syncthreads();
    // (staging the node features through shared memory as k_value_project does was measured slower here: 0.78 ->
    // 1.13 ms; this walk is not bound by its load instructions)
    tarl::tile_walk_nodes(t, [&](int r, int j) {
        const int n = t.n0 + j;
        if (n >= in.N || r >= t.nrows) return;
        const float gs = sm[tarl::tile_slot(t, r, j)];
        if (gs != 0.0f) {
            float x[kIn];
            load_x(in, t.b0 + r, n, x, nullptr);
#pragma unroll
            for (int cI = 0; cI < kIn; ++cI) vals[cI] += gs * x[cI];
        }
    });
    block_store<18>(vals, partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kGrads);
}

// ---- the value head: out[b] = sum_n v[n, b] * w[n] (the node part of final_mlp, src/agents/mpnn_agent.py:359-361) -------
// v is node-major (element (b, n) at n*B + b). A CTA sums a block of kHeadNodes nodes: thread = (row b, node group), the
// rows of one node on consecutive lanes (coalesced); per-CTA partial rows, then a fixed-order sum over the CTAs.
constexpr int kHeadNodes = 2048;
__global__ void __launch_bounds__(kThreads) k_value_head(const float* __restrict__ v, int B, int N,
                                                         const float* __restrict__ w, float* __restrict__ partials) {
    __shared__ float sm[kThreads];
    const int n0 = blockIdx.x * kHeadNodes, n1 = min(n0 + kHeadNodes, N);
    for (int b0 = 0; b0 < B; b0 += kThreads) {                    // (one pass for B <= 256)
        const int rows = min(B - b0, kThreads);
        int Bq = 1;
        while (Bq < rows) Bq <<= 1;                               // rows of this pass rounded up to a power of two
        const int groups = kThreads / Bq;                         // node groups working side by side
        const int r = threadIdx.x & (Bq - 1), grp = threadIdx.x / Bq;
        float acc = 0.0f;
        if (r < rows) {
            for (int n = n0 + grp; n < n1; n += groups) acc += v[(int64_t)n * B + b0 + r] * w[n];
        }
        sm[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < rows) {
            float s = 0.0f;
            for (int gI = 0; gI < groups; ++gI) s += sm[gI * Bq + threadIdx.x];
            partials[(size_t)blockIdx.x * B + b0 + threadIdx.x] = s;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kThreads) k_value_head_finish(const float* __restrict__ partials, int n_parts, int B,
                                                                float* __restrict__ out) {
    const int b = blockIdx.x * kThreads + threadIdx.x;
    if (b >= B) return;
    float s = 0.0f;
    for (int i = 0; i < n_parts; ++i) s += partials[(size_t)i * B + b];
    out[b] = s;
}
// d w[n] = sum_b g[b] * v[n, b]: a warp per node at a time, rows on the lanes, shuffle tree (fixed order)
__global__ void __launch_bounds__(kThreads) k_value_head_wgrad(const float* __restrict__ v, int B, int N,
                                                               const float* __restrict__ g, float* __restrict__ gw) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), n_warps = gridDim.x * (kThreads / 32);
    for (int n = warp; n < N; n += n_warps) {
        float acc = 0.0f;
        for (int b = lane; b < B; b += 32) acc += g[b] * v[(int64_t)n * B + b];
        acc = warp_sum(acc);
        if (lane == 0) gw[n] = acc;
    }
}

// grads[j] = sum over all blocks of partials[.., j]: one CTA per j, strided accumulation then a fixed tree
__global__ void __launch_bounds__(kThreads) k_value_finish(const float* __restrict__ partials, int n_parts,
                                                           float* __restrict__ grads) {
    __shared__ float sm[kThreads];
    const int j = blockIdx.x;
    float s = 0.0f;
    for (int i = threadIdx.x; i < n_parts; i += kThreads) s += partials[(size_t)i * kGrads + j];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) grads[j] = sm[0];
}

// ---- train mode: message dropout ------------------------------------------------------------------------------------
struct Drop {
    const uint32_t* bits; int64_t bits_bs;   // injected keep words [B, E] (batch stride in elements), or nullptr
    const uint32_t* words;                   // backward: the words the forward pass drew, edge-major [E, B], or nullptr
    uint32_t seed_lo, seed_hi, thresh;       // in-kernel stream: input k of (row, edge) is dropped iff its 12-bit field < thresh
    float scale;                             // 1 / (1 - p) (0 when p == 1: everything dropped)
};

__device__ __forceinline__ void philox_raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;   // one IMAD.WIDE each
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0; c1 = (uint32_t)p1; c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// keep word of (row b, edge e): bit k set = message input k survives (k < 16 node inputs, k = 16 edge feature).
// Input k is dropped iff a 12-bit uniform u_k < thresh = round(p * 4096). For thresh <= 256 (p <= 1/16: the reference's
// 0.05) ONE Philox call serves all 17 inputs: 17 four-bit fields are the TOP nibbles of the u_k — a non-zero nibble
// means u_k >= 256 >= thresh, kept, with no further bits (15 of 16 fields) — and the rare zero nibbles take their low
// byte from the 7 spare bytes of the same 128 bits (u_k = low byte); a word with more than 7 zero nibbles (6e-6 of them)
// draws a second block. Exactly Bernoulli(thresh / 4096) per input, independent; half the instructions of the generic
// form below (two calls, 17 twelve-bit fields), which serves larger p.
// flags of the zero nibbles of x (bit k = nibble k of x is 0), k = 0..7, without a loop
__device__ __forceinline__ uint32_t zero_nibbles(uint32_t x) {
    uint32_t z = ~(x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;             // bit 4k
    z = (z | (z >> 3)) & 0x03030303u;                                             // 2 flags per byte
    z = (z | (z >> 6)) & 0x000f000fu;                                             // 4 flags per half word
    return (z | (z >> 12)) & 0xffu;
}

__device__ __forceinline__ uint32_t philox_keep_word(const Drop& d, int b, int e) {
    if (d.thresh <= 256u) {
        uint32_t r[4];
        philox_raw((uint32_t)e, (uint32_t)b, 0u, 0x44524f50u, d.seed_lo, d.seed_hi, r);
        // top nibbles: inputs 0-7 in r[0], 8-15 in r[1], 16 in the low nibble of r[2]
        uint32_t zero = zero_nibbles(r[0]) | (zero_nibbles(r[1]) << 8) | (((r[2] & 0xfu) == 0u) ? (1u << 16) : 0u);
        uint32_t word = 0x1ffffu;
        // spare bytes: r[3] (4), then r[2] >> 4 (3), as one shift register. 1.06 zero nibbles per word on average, but a
        // warp walks as many rounds as its unluckiest lane (3 - 4): the round is kept to a handful of instructions.
        uint64_t spare = (uint64_t)r[3] | ((uint64_t)(r[2] >> 4) << 32);
        if (__popc(zero) <= 7) {
            while (zero != 0u) {
                const uint32_t lowest = zero & (0u - zero);
                zero ^= lowest;
                if ((uint32_t)(spare & 0xffu) < d.thresh) word ^= lowest;
                spare >>= 8;
            }
            return word;
        }
        int used = 0;                                                             // more than 7 zero nibbles: 6e-6 of the words
        while (zero != 0u) {
            const int k = __ffs((int)zero) - 1;
            zero &= zero - 1u;
            uint32_t byte;
            if (used < 4) byte = (r[3] >> (8 * used)) & 0xffu;
            else if (used < 7) byte = (r[2] >> (4 + 8 * (used - 4))) & 0xffu;
            else {
                uint32_t q[4];
                philox_raw((uint32_t)e, (uint32_t)b, (uint32_t)(1 + ((used - 7) >> 4)), 0x44524f50u, d.seed_lo, d.seed_hi, q);
                const int i = (used - 7) & 15;
                byte = (q[i >> 2] >> (8 * (i & 3))) & 0xffu;
            }
            ++used;
            if (byte < d.thresh) word &= ~(1u << k);
        }
        return word;
    }
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {            // 2 draws x 2 sixty-four-bit halves x 5 twelve-bit fields; the first 17 are used
        uint32_t r[4];
        philox_raw((uint32_t)e, (uint32_t)b, (uint32_t)j, 0x44524f51u, d.seed_lo, d.seed_hi, r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t v = (uint64_t)r[2 * h] | ((uint64_t)r[2 * h + 1] << 32);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int k = 10 * j + 5 * h + q;
                if (k <= kIn) word |= ((uint32_t)((v >> (12 * q)) & 0xfffu) >= d.thresh ? 1u : 0u) << k;
            }
        }
    }
    return word;
}
// k = position of edge e in the by-target CSR: the buffers the train-mode passes hand each other (messages, drawn
// words, d z) are kept in THAT order — element (b, k) at k*B + b — so that a (target node, row) pair finds its in-edges'
// entries at addresses that depend on k alone: no edge-id load sits in front of them.
__device__ __forceinline__ uint32_t keep_word(const Drop& d, int B, int b, int e, int k) {
    if (d.bits != nullptr) return d.bits[(int64_t)b * d.bits_bs + e];            // injected, [B, E] by edge id
    if (d.words != nullptr) return d.words[(int64_t)k * B + b];                  // what the forward pass drew, [E, B] by k
    return philox_keep_word(d, b, e);
}

__global__ void __launch_bounds__(kThreads) k_value_dropout_bits(Drop d, int B, int E, uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)B * E) return;
    const int e = (int)(i / B), b = (int)(i % B);
    out[(int64_t)b * E + e] = philox_keep_word(d, b, e);
}

// The words of a training forward pass, by-target position major (element (b, k) at k*B + b): a streaming pass of its
// own — one thread per (position, row), 28 registers, full occupancy — instead of a Philox chain inside the message
// kernel's edge walk, where 24 warps per SM (66 KB of staged planes per CTA) could not hide it (message kernel 3.4 ms
// with the draw, of which the draw was two thirds of the instructions).
__global__ void __launch_bounds__(kThreads) k_value_keep_words(const int32_t* __restrict__ eid, int64_t total, int B,
                                                               Drop d, uint32_t* __restrict__ words) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    int k, b;
    if ((B & (B - 1)) == 0) { k = (int)(i >> (31 - __clz(B))); b = (int)i & (B - 1); }      // no 64-bit division
    else { k = (int)(i / B); b = (int)(i - (int64_t)k * B); }
    words[i] = philox_keep_word(d, b, eid[k]);
}

constexpr int kDropSmemBytes = (kIn * tarl::kTileSmem + (tarl::kTileThreads / 32) * 32 * kNodeDim) * (int)sizeof(float);
constexpr int kPairsPerThread = tarl::kTilePairs / tarl::kTileThreads;      // 4

// Walk 1 of both train-mode kernels, nodes innermost, TWO of a thread's pairs at a time (a pair is two dependent load
// levels — agent index -> agent row — and 66 KB of planes leave 24 warps per SM to hide them): store(c, slot, x_c).
// The gather of a pair's 9 agent features is what the L1 spends its time on in this walk: 9 scalar loads per warp, each
// on 32 different 36-byte rows = 9 x 32 tag look-ups, against 7 + 2 for everything else a warp loads here. pack =
// the agent table re-laid as 48-byte rows (k_value_pack_agents, once per call): three 128-bit loads per pair.
__global__ void __launch_bounds__(256) k_value_pack_agents(const float* __restrict__ af, int af_rows, float4* __restrict__ pack) {
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= af_rows) return;
    const float* q = af + (int64_t)a * kAgentDim;
    pack[3 * (int64_t)a] = make_float4(q[0], q[1], q[2], q[3]);
    pack[3 * (int64_t)a + 1] = make_float4(q[4], q[5], q[6], q[7]);
    pack[3 * (int64_t)a + 2] = make_float4(q[8], 0.0f, 0.0f, 0.0f);
}

template <typename Keep, typename Store>
__device__ __forceinline__ void stage_inputs(const tarl::Tile& t, const Inputs& in, const float4* __restrict__ pack,
                                             float* __restrict__ warp_buf, int32_t* flags, Keep keep, Store store) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int segs = t.TN >> 5;
    const int total = t.Bp * segs;
    constexpr int kWarps = tarl::kTileThreads / 32;
    const bool dense = in.nf_rs == kNodeDim;
    for (int u0 = warp; u0 < total; u0 += 2 * kWarps) {
        float x[2][kIn];
        int slot[2];
        bool live[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int u = u0 + h * kWarps;                        // warp-uniform
            const int r = u / segs, j = (u - r * segs) * 32 + lane;
            const int n = t.n0 + j;
            const bool row_live = u < total && r < t.nrows;       // warp-uniform
            live[h] = row_live && n < in.N && keep(n);
            slot[h] = tarl::tile_slot(t, r, j);
            if (pack == nullptr) {
                if (live[h]) load_x(in, t.b0 + r, n, x[h], flags);
                continue;
            }
            long long a = 0;
            if (live[h]) a = agent_row(in, t.b0 + r, n, flags);
            if (row_live) load_nf_staged(in, dense, t.b0 + r, n - lane, lane, live[h], warp_buf, x[h]);
            if (live[h]) {
                const float4 p0 = pack[3 * a], p1 = pack[3 * a + 1], p2 = pack[3 * a + 2];
                x[h][7] = p0.x; x[h][8] = p0.y; x[h][9] = p0.z; x[h][10] = p0.w;
                x[h][11] = p1.x; x[h][12] = p1.y; x[h][13] = p1.z; x[h][14] = p1.w;
                x[h][15] = p2.x;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!live[h]) continue;
#pragma unroll
            for (int c = 0; c < kIn; ++c) store(c, slot[h], x[h][c]);
        }
    }
}

// (target node, row) tiles (tile_map.cuh). Walk 1: the 16 message inputs of every pair are read the way the
// observation is laid out and parked in shared memory as products (x_k * scale) * w_k — 16 planes of the tile.
// Walk 2, rows innermost: msg = tanh(sum_k keep_k * product_k + keep_16 * (f * scale) * w_16 + w0) for every in-edge of
// the node (x * (mask / (1 - p)) as ATen's dropout computes it, then the 17-term dot product), written with the row
// innermost at the edge's by-target position k (see keep_word). A thread owns four pairs and walks them TOGETHER, two
// in-edges of each per round: 8 edge ids, 8 keep words, then 8 edge features are in flight at once, and the products
// are read from shared memory where they are used instead of being held in 16 registers per pair (which is what kept
// the walk to one pair, i.e. one dependent load chain, at a time: 2.4 ms for 32 rows x 6.0 M edges without the draw).
__global__ void __launch_bounds__(tarl::kTileThreads, 3) k_value_message_dropout(tarl_csr by_dst, Inputs in, int Bp, Drop d,
                                                                                 const float* __restrict__ w,
                                                                                 const float* __restrict__ w0,
                                                                                 float* __restrict__ msg,
                                                                                 uint32_t* __restrict__ words_out,
                                                                                 int32_t* __restrict__ flags,
                                                                                 const float4* __restrict__ pack) {
    extern __shared__ float xs[];                                 // [kIn][kTileSmem], then a 224-float buffer per warp
    const tarl::Tile t = tarl::tile_here(in.B, Bp);
    stage_inputs(t, in, pack, xs + kIn * tarl::kTileSmem + (threadIdx.x >> 5) * 32 * kNodeDim, flags, [&](int) { return true; },
                 [&](int c, int slot, float v) { xs[c * tarl::kTileSmem + slot] = (v * d.scale) * w[c]; });
    __syncthreads();
    const float we = w[kIn], bias = w0[0];
    int k0[kPairsPerThread], k1[kPairsPerThread];
#pragma unroll
    for (int q = 0; q < kPairsPerThread; ++q) {
        const int p = threadIdx.x + q * tarl::kTileThreads;
        const int r = p & (t.Bp - 1), n = t.n0 + (p >> t.sh);
        const bool live = n < in.N && r < t.nrows;
        k0[q] = live ? by_dst.ptr[n] : 0;
        k1[q] = live ? by_dst.ptr[n + 1] : 0;
    }
    for (;;) {
        bool any = false;
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) any = any || k0[q] < k1[q];
        if (!any) break;
        int e[kPairsPerThread][2];
        uint32_t word[kPairsPerThread][2];
        float f[kPairsPerThread][2];
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i) e[q][i] = k0[q] + i < k1[q] ? by_dst.eid[k0[q] + i] : -1;
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            const int b = t.b0 + ((threadIdx.x + q * tarl::kTileThreads) & (t.Bp - 1));
#pragma unroll
            for (int i = 0; i < 2; ++i) word[q][i] = e[q][i] >= 0 ? keep_word(d, in.B, b, e[q][i], k0[q] + i) : 0u;
        }
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            const int b = t.b0 + ((threadIdx.x + q * tarl::kTileThreads) & (t.Bp - 1));
#pragma unroll
            for (int i = 0; i < 2; ++i) f[q][i] = e[q][i] >= 0 ? in.ef[(int64_t)b * in.ef_bs + e[q][i]] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            const int p = threadIdx.x + q * tarl::kTileThreads;
            const int r = p & (t.Bp - 1);
            const int slot = tarl::tile_slot(t, r, p >> t.sh);
            const int b = t.b0 + r;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (e[q][i] < 0) continue;
                const int64_t o = (int64_t)(k0[q] + i) * in.B + b;
                if (words_out != nullptr) words_out[o] = word[q][i];             // backward reads them back
                float z = 0.0f;
#pragma unroll
                for (int c = 0; c < kIn; ++c) z += ((word[q][i] >> c) & 1u) ? xs[c * tarl::kTileSmem + slot] : 0.0f;
                z += ((word[q][i] >> kIn) & 1u) ? (f[q][i] * d.scale) * we : 0.0f;
                msg[o] = tanhf(z + bias);
            }
            k0[q] = min(k0[q] + 2, k1[q]);
        }
    }
}

// one thread per (source node, batch row): mean of the stored messages in ascending edge id, then the node update
__global__ void __launch_bounds__(kThreads) k_value_aggregate_msg(tarl_csr by_src, const int32_t* __restrict__ src_pos,
                                                                  int B, int N,
                                                                  const float* __restrict__ a, const float* __restrict__ c,
                                                                  const float* __restrict__ msg, float* __restrict__ mean,
                                                                  float* __restrict__ v) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)B * N) return;
    const int n = (int)(i / B), b = (int)(i % B);
    const int k0 = by_src.ptr[n], k1 = by_src.ptr[n + 1];
    float acc = 0.0f;
    for (int k = k0; k < k1; ++k) acc += msg[(int64_t)src_pos[k] * B + b];       // ascending edge id, as scatter-mean adds
    const float m = k1 > k0 ? acc / (float)(k1 - k0) : 0.0f;
    mean[i] = m;
    v[i] = tanhf(a[0] * m + c[0]);
}

// Message backward with dropout, in two passes.
// Pass 1, edge-parallel and streaming (one thread per by-target position k and 4 rows): d z = gm[source, b] * (1 - msg^2)
// overwrites the message in place; the two gradients that need nothing of the target node ride along —
// d w_16 += d z * keep_16 * scale * f and d w0 += d z — as per-CTA partial sums (fixed grid, fixed order).
constexpr int kDzCtas = 148 * 8;
__global__ void __launch_bounds__(kThreads) k_value_dz(tarl_csr by_dst, Inputs in, Drop d, float* __restrict__ msg,
                                                       const float* __restrict__ gm, float* __restrict__ partials) {
    const int B = in.B;
    const int64_t total = (int64_t)by_dst.n_edges * B;
    float vals[2] = {0.0f, 0.0f};
    const bool vec = (B & 3) == 0;
    const int64_t step = (int64_t)gridDim.x * kThreads;
    if (vec) {
        const int64_t quads = total >> 2;
        const int qpe = B >> 2;                                   // quads per edge
        const int qsh = (qpe & (qpe - 1)) == 0 ? 31 - __clz(qpe) : -1;
        // (two quads per round, with both edges' first-level loads requested before either's second level, was measured:
        // 59 -> 72 registers, 0.71 -> 0.89 ms)
        for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < quads; q += step) {
            const int k = qsh >= 0 ? (int)(q >> qsh) : (int)(q / qpe);
            const int b = (int)(q - (int64_t)k * qpe) << 2;
            const int e = by_dst.eid[k];
            const float4 m = *reinterpret_cast<const float4*>(msg + (int64_t)k * B + b);
            const float4 g = *reinterpret_cast<const float4*>(gm + (int64_t)by_dst.idx[k] * B + b);
            float4 z;
            z.x = g.x * (1.0f - m.x * m.x); z.y = g.y * (1.0f - m.y * m.y);
            z.z = g.z * (1.0f - m.z * m.z); z.w = g.w * (1.0f - m.w * m.w);
            *reinterpret_cast<float4*>(msg + (int64_t)k * B + b) = z;
            const float zz[4] = {z.x, z.y, z.z, z.w};
            uint32_t wv[4];
            if (d.bits == nullptr && d.words != nullptr) {        // the four rows' stored words are one 128-bit vector
                const uint4 w4 = *reinterpret_cast<const uint4*>(d.words + (int64_t)k * B + b);
                wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) wv[r] = keep_word(d, B, b + r, e, k);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if ((wv[r] >> kIn) & 1u) vals[0] += zz[r] * (in.ef[(int64_t)(b + r) * in.ef_bs + e] * d.scale);
                vals[1] += zz[r];
            }
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += step) {
            const int k = (int)(i / B), b = (int)(i - (int64_t)k * B);
            const int e = by_dst.eid[k];
            const float m = msg[i];
            const float z = gm[(int64_t)by_dst.idx[k] * B + b] * (1.0f - m * m);
            msg[i] = z;
            const uint32_t word = keep_word(d, B, b, e, k);
            if ((word >> kIn) & 1u) vals[0] += z * (in.ef[(int64_t)b * in.ef_bs + e] * d.scale);
            vals[1] += z;
        }
    }
    float all[kGrads];
#pragma unroll
    for (int j = 0; j < kGrads; ++j) all[j] = 0.0f;
    all[16] = vals[0]; all[17] = vals[1];
    block_store<kGrads>(all, partials + (size_t)blockIdx.x * kGrads);
}

// Pass 2, (target node, row) tiles as in the forward pass: d w_k += sum over in-edges of keep_k * d z * (scale * x_k).
// The in-edges' d z and keep words sit at k*B + b: nothing has to be loaded to find them. Like the forward walk, a
// thread takes its four pairs together, two in-edges of each per round (16 loads in flight), and multiplies by the
// staged input where the edge is consumed — one predicated FFMA per (edge, input) into the thread's 16 running sums,
// no per-pair accumulators. (History: the round-1 form walked edge id -> message, word, source -> gm one dependent load
// after the other, 4.7 ms at 7 % of the DRAM bandwidth for 32 rows x 6.0 M edges; one pair at a time with per-pair
// sums, however its loads were chunked, stayed at 2.0 - 2.3 ms.)
__global__ void __launch_bounds__(tarl::kTileThreads, 3) k_value_edge_grad_dropout(tarl_csr by_dst, Inputs in, int Bp, Drop d,
                                                                                const float* __restrict__ dz,
                                                                                float* __restrict__ partials,
                                                                                const float4* __restrict__ pack) {
    extern __shared__ float xs[];                                 // [kIn][kTileSmem], then a 224-float buffer per warp
    const tarl::Tile t = tarl::tile_here(in.B, Bp);
    // (a node nobody points at needs no inputs)
    stage_inputs(t, in, pack, xs + kIn * tarl::kTileSmem + (threadIdx.x >> 5) * 32 * kNodeDim, nullptr,
                 [&](int n) { return by_dst.ptr[n] != by_dst.ptr[n + 1]; },
                 [&](int c, int slot, float v) { xs[c * tarl::kTileSmem + slot] = v * d.scale; });
    __syncthreads();
    float vals[kIn];
#pragma unroll
    for (int j = 0; j < kIn; ++j) vals[j] = 0.0f;
    const bool stored = d.bits == nullptr && d.words != nullptr;  // words at k*B + b
    int k0[kPairsPerThread], k1[kPairsPerThread];
#pragma unroll
    for (int q = 0; q < kPairsPerThread; ++q) {
        const int p = threadIdx.x + q * tarl::kTileThreads;
        const int r = p & (t.Bp - 1), n = t.n0 + (p >> t.sh);
        const bool live = n < in.N && r < t.nrows;
        k0[q] = live ? by_dst.ptr[n] : 0;
        k1[q] = live ? by_dst.ptr[n + 1] : 0;
    }
    for (;;) {
        bool any = false;
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) any = any || k0[q] < k1[q];
        if (!any) break;
        float gz[kPairsPerThread][2];
        uint32_t word[kPairsPerThread][2];
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            const int b = t.b0 + ((threadIdx.x + q * tarl::kTileThreads) & (t.Bp - 1));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool live = k0[q] + i < k1[q];
                const int64_t o = (int64_t)(k0[q] + i) * in.B + b;
                gz[q][i] = live ? dz[o] : 0.0f;
                word[q][i] = 0u;
                if (live) word[q][i] = stored ? d.words[o] : keep_word(d, in.B, b, by_dst.eid[k0[q] + i], k0[q] + i);
            }
        }
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            const int p = threadIdx.x + q * tarl::kTileThreads;
            const int slot = tarl::tile_slot(t, p & (t.Bp - 1), p >> t.sh);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int c = 0; c < kIn; ++c)
                    if ((word[q][i] >> c) & 1u) vals[c] += gz[q][i] * xs[c * tarl::kTileSmem + slot];
            }
            k0[q] = min(k0[q] + 2, k1[q]);
        }
    }
    float all[18];
#pragma unroll
    for (int j = 0; j < kIn; ++j) all[j] = vals[j];
    all[16] = 0.0f; all[17] = 0.0f;
    block_store<18>(all, partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kGrads);
}

// both dropout kernels stage 16 planes of a tile: 66 KB of dynamic shared memory, opted in once per process
int drop_smem_ready() {
    static const int rc = [] {
        const bool ok =
            cudaFuncSetAttribute(k_value_message_dropout, cudaFuncAttributeMaxDynamicSharedMemorySize, kDropSmemBytes) == cudaSuccess &&
            cudaFuncSetAttribute(k_value_edge_grad_dropout, cudaFuncAttributeMaxDynamicSharedMemorySize, kDropSmemBytes) == cudaSuccess;
        return ok ? TARL_OK : TARL_E_LAUNCH;
    }();
    return rc;
}

Drop make_drop(const uint32_t* keep_bits, int64_t keep_batch_stride, uint64_t seed, float p,
               const uint32_t* drawn_words = nullptr) {
    Drop d;
    d.bits = keep_bits;
    d.bits_bs = keep_batch_stride;
    d.words = keep_bits == nullptr ? drawn_words : nullptr;
    d.seed_lo = (uint32_t)seed;
    d.seed_hi = (uint32_t)(seed >> 32);
    const double t = (double)p * 4096.0 + 0.5;
    d.thresh = t < 0.0 ? 0u : (t > 4096.0 ? 4096u : (uint32_t)t);
    d.scale = p < 1.0f ? 1.0f / (1.0f - p) : 0.0f;
    return d;
}

int check(const tarl_csr* c, int n_nodes) {
    if (c == nullptr || c->n_rows != n_nodes || c->n_edges < 0) return TARL_E_BADARG;
    if (n_nodes > 0 && c->ptr == nullptr) return TARL_E_BADARG;
    if (c->n_edges > 0 && (c->idx == nullptr || c->eid == nullptr)) return TARL_E_BADARG;
    return TARL_OK;
}

}  // namespace

extern "C" {

int32_t tarl_value_mp_partial_count(int32_t n_nodes, int32_t batch) {
    // one row per (node, row) tile + the rows of the train-mode d z pass (a fixed grid)
    return (n_nodes > 0 && batch > 0) ? tarl::tile_count(n_nodes, batch) + kDzCtas : 0;
}

int tarl_value_mp_forward(const tarl_csr* by_source, const float* node_features, int64_t nf_batch_stride,
                          int64_t nf_row_stride, const float* edge_features, int64_t ef_batch_stride,
                          const int64_t* agent_index, const float* agent_features, int32_t agent_rows,
                          const float* msg_weight, const float* msg_bias, const float* node_weight,
                          const float* node_bias, int32_t batch, int32_t n_nodes, float* agent_proj, float* proj,
                          float* mean, float* v, int32_t* flags, void* stream) {
    if (batch < 0 || n_nodes < 0 || agent_rows < 1) return TARL_E_BADARG;
    int rc = check(by_source, n_nodes);
    if (rc != TARL_OK) return rc;
    if (batch == 0 || n_nodes == 0) return TARL_OK;
    if (!node_features || !agent_index || !agent_features || !msg_weight || !msg_bias || !node_weight || !node_bias ||
        !agent_proj || !proj || !mean || !v || !flags || (by_source->n_edges > 0 && !edge_features))
        return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Inputs in = {node_features, nf_batch_stride, nf_row_stride, edge_features, ef_batch_stride,
                       reinterpret_cast<const long long*>(agent_index), agent_features, agent_rows, batch, n_nodes,
                       by_source->n_edges};
    const int nb = blocks_for((int64_t)batch * n_nodes);
    k_value_agent_project<<<(agent_rows + 255) / 256, 256, 0, s>>>(agent_features, agent_rows, msg_weight, agent_proj);
    k_value_project<<<tarl::tile_grid(n_nodes, batch), tarl::kTileThreads, 0, s>>>(in, tarl::tile_rows_pow2(batch), msg_weight,
                                                                                   agent_proj, proj, flags);
    const bool vec4 = (batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(mean) |
                                            reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec4)
        k_value_aggregate4<<<blocks_for((int64_t)(batch >> 2) * n_nodes), kThreads, 0, s>>>(
            *by_source, in, msg_weight, msg_bias, node_weight, node_bias, proj, mean, v);
    else
        k_value_aggregate<<<nb, kThreads, 0, s>>>(*by_source, in, msg_weight, msg_bias, node_weight, node_bias, proj, mean, v);
    return launch_status();
}

int tarl_value_mp_backward(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                           int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                           int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                           int32_t agent_rows, const float* msg_weight, const float* msg_bias, const float* node_weight,
                           int32_t batch, int32_t n_nodes, const float* proj, const float* mean, const float* v,
                           const float* grad_v, int64_t gv_batch_stride, int64_t gv_node_stride, const float* head_g,
                           const float* head_w, float* gm, float* partials, float* grads, void* stream) {
    if (batch < 0 || n_nodes < 0 || agent_rows < 1 || grads == nullptr) return TARL_E_BADARG;
    int rc = check(by_source, n_nodes);
    if (rc == TARL_OK) rc = check(by_target, n_nodes);
    if (rc != TARL_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (batch == 0 || n_nodes == 0) {
        return cudaMemsetAsync(grads, 0, sizeof(float) * kGrads, s) == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    }
    if (!node_features || !agent_index || !agent_features || !msg_weight || !msg_bias || !node_weight || !proj ||
        !mean || !v || !gm || !partials || (by_source->n_edges > 0 && !edge_features))
        return TARL_E_BADARG;
    if ((head_g != nullptr) != (head_w != nullptr) || (grad_v == nullptr && head_g == nullptr)) return TARL_E_BADARG;
    const Inputs in = {node_features, nf_batch_stride, nf_row_stride, edge_features, ef_batch_stride,
                       reinterpret_cast<const long long*>(agent_index), agent_features, agent_rows, batch, n_nodes,
                       by_source->n_edges};
    const dim3 grid = tarl::tile_grid(n_nodes, batch);
    const int Bp = tarl::tile_rows_pow2(batch);
    k_value_node_grad<<<grid, tarl::kTileThreads, 0, s>>>(*by_source, batch, Bp, n_nodes, node_weight, mean, v, grad_v,
                                                          gv_batch_stride, gv_node_stride, head_g, head_w, gm, partials);
    // 4 rows per thread when every row chunk of the tile is a whole multiple of 4 rows and the B-vectors are 16-byte aligned
    const bool vec4 = (batch & 3) == 0 && Bp >= 4 &&
                      ((reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(gm)) & 15) == 0;
    k_value_edge_grad<<<grid, tarl::kTileThreads, 0, s>>>(*by_target, in, Bp, vec4, msg_weight, msg_bias, proj, gm, partials);
    k_value_finish<<<kGrads, kThreads, 0, s>>>(partials, (int)(grid.x * grid.y), grads);
    return launch_status();
}

int tarl_value_mp_dropout_bits(uint64_t seed, float p, int32_t batch, int32_t n_edges, uint32_t* keep_bits, void* stream) {
    if (batch < 0 || n_edges < 0 || !(p >= 0.0f && p <= 1.0f)) return TARL_E_BADARG;
    if (batch == 0 || n_edges == 0) return TARL_OK;
    if (keep_bits == nullptr) return TARL_E_BADARG;
    k_value_dropout_bits<<<blocks_for((int64_t)batch * n_edges), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        make_drop(nullptr, 0, seed, p), batch, n_edges, keep_bits);
    return launch_status();
}

int tarl_value_mp_forward_dropout(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                                  int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                                  int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                                  int32_t agent_rows, const float* msg_weight, const float* msg_bias,
                                  const float* node_weight, const float* node_bias, int32_t batch, int32_t n_nodes,
                                  const uint32_t* keep_bits, int64_t keep_batch_stride, uint64_t seed, float p,
                                  const int32_t* source_pos, uint32_t* keep_words, float* agent_pack, float* msg,
                                  float* mean, float* v, int32_t* flags, void* stream) {
    if (batch < 0 || n_nodes < 0 || agent_rows < 1 || !(p >= 0.0f && p <= 1.0f)) return TARL_E_BADARG;
    int rc = check(by_source, n_nodes);
    if (rc == TARL_OK) rc = check(by_target, n_nodes);
    if (rc != TARL_OK) return rc;
    if (by_source->n_edges != by_target->n_edges) return TARL_E_BADARG;
    if (batch == 0 || n_nodes == 0) return TARL_OK;
    if (!node_features || !agent_index || !agent_features || !msg_weight || !msg_bias || !node_weight || !node_bias ||
        !mean || !v || !flags || (by_source->n_edges > 0 && (!edge_features || !msg || !source_pos)))
        return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Inputs in = {node_features, nf_batch_stride, nf_row_stride, edge_features, ef_batch_stride,
                       reinterpret_cast<const long long*>(agent_index), agent_features, agent_rows, batch, n_nodes,
                       by_source->n_edges};
    const int nb = blocks_for((int64_t)batch * n_nodes);
    if ((rc = drop_smem_ready()) != TARL_OK) return rc;
    if (agent_pack != nullptr) {
        if ((reinterpret_cast<uintptr_t>(agent_pack) & 15) != 0) return TARL_E_BADARG;
        k_value_pack_agents<<<(agent_rows + 255) / 256, 256, 0, s>>>(agent_features, agent_rows, reinterpret_cast<float4*>(agent_pack));
    }
    // words drawn here: a streaming pass writes them first and the message kernel reads them like the backward pass does
    const bool draw_first = keep_bits == nullptr && keep_words != nullptr && by_target->n_edges > 0;
    if (draw_first) {
        const int64_t total = (int64_t)by_target->n_edges * batch;
        k_value_keep_words<<<blocks_for(total), kThreads, 0, s>>>(by_target->eid, total, batch,
                                                                  make_drop(nullptr, 0, seed, p), keep_words);
    }
    k_value_message_dropout<<<tarl::tile_grid(n_nodes, batch), tarl::kTileThreads, kDropSmemBytes, s>>>(
        *by_target, in, tarl::tile_rows_pow2(batch),
        make_drop(keep_bits, keep_batch_stride, seed, p, draw_first ? keep_words : nullptr), msg_weight, msg_bias, msg,
        nullptr, flags, reinterpret_cast<const float4*>(agent_pack));
    k_value_aggregate_msg<<<nb, kThreads, 0, s>>>(*by_source, source_pos, batch, n_nodes, node_weight, node_bias, msg, mean, v);
    return launch_status();
}

int tarl_value_mp_backward_dropout(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                                   int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                                   int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                                   int32_t agent_rows, const float* node_weight, int32_t batch, int32_t n_nodes,
                                   const uint32_t* keep_bits, int64_t keep_batch_stride, uint64_t seed, float p,
                                   const uint32_t* keep_words, float* agent_pack, float* msg, const float* mean,
                                   const float* v, const float* grad_v,
                                   int64_t gv_batch_stride, int64_t gv_node_stride, const float* head_g,
                                   const float* head_w, float* gm, float* partials, float* grads, void* stream) {
    if (batch < 0 || n_nodes < 0 || agent_rows < 1 || grads == nullptr || !(p >= 0.0f && p <= 1.0f)) return TARL_E_BADARG;
    int rc = check(by_source, n_nodes);
    if (rc == TARL_OK) rc = check(by_target, n_nodes);
    if (rc != TARL_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (batch == 0 || n_nodes == 0) {
        return cudaMemsetAsync(grads, 0, sizeof(float) * kGrads, s) == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    }
    if (!node_features || !agent_index || !agent_features || !node_weight || !mean || !v || !gm ||
        !partials || (by_source->n_edges > 0 && (!edge_features || !msg)))
        return TARL_E_BADARG;
    if ((head_g != nullptr) != (head_w != nullptr) || (grad_v == nullptr && head_g == nullptr)) return TARL_E_BADARG;
    const Inputs in = {node_features, nf_batch_stride, nf_row_stride, edge_features, ef_batch_stride,
                       reinterpret_cast<const long long*>(agent_index), agent_features, agent_rows, batch, n_nodes,
                       by_source->n_edges};
    const dim3 grid = tarl::tile_grid(n_nodes, batch);
    const int Bp = tarl::tile_rows_pow2(batch);
    k_value_node_grad<<<grid, tarl::kTileThreads, 0, s>>>(*by_source, batch, Bp, n_nodes, node_weight, mean, v, grad_v,
                                                          gv_batch_stride, gv_node_stride, head_g, head_w, gm, partials);
    if ((rc = drop_smem_ready()) != TARL_OK) return rc;
    const Drop drop = make_drop(keep_bits, keep_batch_stride, seed, p, keep_words);
    const int n_tiles = (int)(grid.x * grid.y);
    if (by_target->n_edges > 0)
        k_value_dz<<<kDzCtas, kThreads, 0, s>>>(*by_target, in, drop, msg, gm, partials + (size_t)n_tiles * kGrads);
    else if (cudaMemsetAsync(partials + (size_t)n_tiles * kGrads, 0, sizeof(float) * kGrads * kDzCtas, s) != cudaSuccess)
        return TARL_E_LAUNCH;
    if (agent_pack != nullptr) {
        if ((reinterpret_cast<uintptr_t>(agent_pack) & 15) != 0) return TARL_E_BADARG;
        k_value_pack_agents<<<(agent_rows + 255) / 256, 256, 0, s>>>(agent_features, agent_rows, reinterpret_cast<float4*>(agent_pack));
    }
    k_value_edge_grad_dropout<<<grid, tarl::kTileThreads, kDropSmemBytes, s>>>(*by_target, in, Bp, drop, msg, partials,
                                                                               reinterpret_cast<const float4*>(agent_pack));
    k_value_finish<<<kGrads, kThreads, 0, s>>>(partials, n_tiles + kDzCtas, grads);
    return launch_status();
}

int32_t tarl_value_head_partial_count(int32_t n_nodes) { return n_nodes > 0 ? (n_nodes + kHeadNodes - 1) / kHeadNodes : 0; }

int tarl_value_head_forward(const float* v, int32_t batch, int32_t n_nodes, const float* head_weight, float* partials,
                            float* out, void* stream) {
    if (batch < 0 || n_nodes < 0) return TARL_E_BADARG;
    if (batch == 0) return TARL_OK;
    if (out == nullptr) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_nodes == 0) return cudaMemsetAsync(out, 0, sizeof(float) * batch, s) == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    if (!v || !head_weight || !partials) return TARL_E_BADARG;
    const int nb = tarl_value_head_partial_count(n_nodes);
    k_value_head<<<nb, kThreads, 0, s>>>(v, batch, n_nodes, head_weight, partials);
    k_value_head_finish<<<blocks_for(batch), kThreads, 0, s>>>(partials, nb, batch, out);
    return launch_status();
}

int tarl_value_head_weight_grad(const float* v, int32_t batch, int32_t n_nodes, const float* grad_out, float* grad_weight,
                                void* stream) {
    if (batch < 0 || n_nodes < 0) return TARL_E_BADARG;
    if (n_nodes == 0) return TARL_OK;
    if (grad_weight == nullptr) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (batch == 0) return cudaMemsetAsync(grad_weight, 0, sizeof(float) * n_nodes, s) == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    if (!v || !grad_out) return TARL_E_BADARG;
    const int64_t warps = n_nodes;
    const int blocks = (int)((warps + kThreads / 32 - 1) / (kThreads / 32));
    k_value_head_wgrad<<<blocks < 148 * 16 ? blocks : 148 * 16, kThreads, 0, s>>>(v, batch, n_nodes, grad_out, grad_weight);
    return launch_status();
}

}  // extern "C"

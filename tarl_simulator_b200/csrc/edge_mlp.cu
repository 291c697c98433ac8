// edge_mlp.cu — MPNNPolicyNet's per-edge MLPs, forward and backward (sm_100a, fp32).
//
// Reference: /root/reference/src/agents/mpnn_agent.py:30-50 (the modules) and :220-231 (their only use, the two
// commented-out bodies of update_edges):
//   x[b,n]      = [node_features[b,n,0:7] ‖ agent_features[agent_index[b,n], 0:9]]              (:163-167, :181-184)
//   edge_mlp    : logit[b,e] = L3(relu(L2(relu(L1([x[b,src e] ‖ x[b,dst e] ‖ edge_attr[b,e]])))))   33 -> 64 -> 32 -> 1
//   edge_mlp_test: logit[b,e] = L2(relu(L1([x[b,src e] ‖ x[b,dst e]])))                              32 -> 16 -> 1
// (edge_mlp_test is written against 1-wide embeddings in the comment, which does not match its 32-wide first layer; the
// reading here feeds it the same 16-wide x rows as edge_mlp.)
//
// Two forward implementations with identical interfaces:
//   * fp32 pipe (this file, every shape): one thread per (row, edge) pair, weights broadcast from shared memory;
//   * tensor cores (edge_mlp_tc.cu, edge_mlp only): both hidden layers as tcgen05.mma with the activations in TMEM.
// Backward (parameter gradients only — the observation is a leaf): per tile of 128 pairs the activations and their
// gradients are recomputed, parked in shared memory and reduced over the tile's pairs as small outer-product sums held in
// registers across a CTA's tiles; per-CTA partial vectors, then a fixed-order finish: run-to-run deterministic.
// Pairs are tiled ROW-major (128 consecutive edges of one batch row): edge endpoints and outputs are contiguous, and
// consecutive edges of a source-sorted edge list share their source row of x.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace tarl {
// csrc/edge_mlp_tc.cu
int edge_mlp_forward_tc(const int32_t* src, const int32_t* dst, int E, int B, const float* x, int64_t x_bs, const float* ea,
                        int64_t ea_bs, const float* const* weights, float* scratch, float* out, int64_t out_bs,
                        int64_t out_es, cudaStream_t s);
}  // namespace tarl

namespace {

constexpr int kX = 16, kAgentDim = 9;
constexpr int kTile = 128;                       // pairs per tile = threads per CTA

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

// ---------------------------------------------------------------------------------------------------- x assembly
__global__ void __launch_bounds__(256) k_edge_mlp_x(const float* __restrict__ nf, int64_t nf_bs, int64_t nf_rs,
                                                    const long long* __restrict__ ai, const float* __restrict__ af,
                                                    int af_rows, int B, int N, float* __restrict__ x,
                                                    int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)B * N) return;
    const int b = (int)(i / N), n = (int)(i - (int64_t)b * N);
    const float* p = nf + b * nf_bs + n * nf_rs;
    long long a = ai[i];
    if (a < 0) a += af_rows;                                      // torch advanced indexing wraps negatives
    if (a < 0 || a >= af_rows) {
        atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_AGENT_RANGE);
        a = 0;
    }
    const float* q = af + a * kAgentDim;
    float4* o = reinterpret_cast<float4*>(x + i * kX);
    o[0] = make_float4(p[0], p[1], p[2], p[3]);
    o[1] = make_float4(p[4], p[5], p[6], q[0]);
    o[2] = make_float4(q[1], q[2], q[3], q[4]);
    o[3] = make_float4(q[5], q[6], q[7], q[8]);
}

// ---------------------------------------------------------------------------------------------------- the MLP itself
// IN inputs (33 or 32), H1 first hidden width, H2 second hidden width or 0 (then the output layer follows H1 directly).
template <int IN, int H1, int H2>
struct Shape {
    static constexpr int kInP = (IN + 3) / 4 * 4;                 // padded row of W1 in shared memory
    static constexpr int kW2P = H2 + 4;                           // row pitch of the transposed W2 copy (see k_edge_mlp_bwd)
    static constexpr int kLast = H2 > 0 ? H2 : H1;                // inputs of the output layer
    static constexpr int kParams = H1 * IN + H1 + (H2 > 0 ? H2 * H1 + H2 : 0) + kLast + 1;
    // offsets into the flat gradient vector: W1, b1, [W2, b2,] w_out, b_out — the order of Sequential.parameters()
    static constexpr int oW1 = 0, oB1 = H1 * IN, oW2 = oB1 + H1, oB2 = oW2 + (H2 > 0 ? H2 * H1 : 0),
                         oWo = oB2 + (H2 > 0 ? H2 : 0), oBo = oWo + kLast;
    // shared-memory copy of the weights (floats)
    static constexpr int sW1 = 0, sB1 = H1 * kInP, sW2 = sB1 + H1, sB2 = sW2 + (H2 > 0 ? kW2P * H1 : 0),
                         sWo = sB2 + (H2 > 0 ? H2 : 0), sBo = sWo + kLast, kWeightFloats = (sBo + 1 + 3) / 4 * 4;
};

struct Weights {
    const float *w1, *b1, *w2, *b2, *wo, *bo;                     // w2 / b2 unused when H2 == 0
};

template <int IN, int H1, int H2>
__device__ __forceinline__ void stage_weights(const Weights& w, float* sm) {
    using S = Shape<IN, H1, H2>;
    for (int i = threadIdx.x; i < H1 * S::kInP; i += blockDim.x) {
        const int r = i / S::kInP, c = i - r * S::kInP;
        sm[S::sW1 + i] = c < IN ? w.w1[r * IN + c] : 0.0f;
    }
    for (int i = threadIdx.x; i < H1; i += blockDim.x) sm[S::sB1 + i] = w.b1[i];
    if (H2 > 0) {
        for (int i = threadIdx.x; i < H2 * H1; i += blockDim.x) {         // transposed: W2t[i][j], the j of one i contiguous
            const int j = i / H1, c = i - j * H1;
            sm[S::sW2 + c * S::kW2P + j] = w.w2[i];
        }
        for (int i = threadIdx.x; i < H2; i += blockDim.x) sm[S::sB2 + i] = w.b2[i];
    }
    for (int i = threadIdx.x; i < S::kLast; i += blockDim.x) sm[S::sWo + i] = w.wo[i];
    if (threadIdx.x == 0) sm[S::sBo] = w.bo[0];
}

// The inputs of pair (b, e): [x[b, src e] ‖ x[b, dst e] ‖ edge_attr[b, e]] padded with zeros to kInP.
template <int IN>
__device__ __forceinline__ void load_inputs(const float* __restrict__ x, int64_t x_bs, const float* __restrict__ ea,
                                            int64_t ea_bs, int b, int e, int s, int d, float (&a)[(IN + 3) / 4 * 4]) {
    const float4* xi = reinterpret_cast<const float4*>(x + b * x_bs + (int64_t)s * kX);
    const float4* xj = reinterpret_cast<const float4*>(x + b * x_bs + (int64_t)d * kX);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 u = xi[q], v = xj[q];
        a[4 * q] = u.x; a[4 * q + 1] = u.y; a[4 * q + 2] = u.z; a[4 * q + 3] = u.w;
        a[kX + 4 * q] = v.x; a[kX + 4 * q + 1] = v.y; a[kX + 4 * q + 2] = v.z; a[kX + 4 * q + 3] = v.w;
    }
    if (IN > 2 * kX) {
        a[2 * kX] = ea[b * ea_bs + e];
#pragma unroll
        for (int c = 2 * kX + 1; c < (IN + 3) / 4 * 4; ++c) a[c] = 0.0f;
    }
}

// One hidden unit of the first layer: relu(W1[i] . a + b1[i]), the weight row broadcast from shared memory (128-bit reads).
template <int IN, int H1, int H2>
__device__ __forceinline__ float hidden1(const float* sm, const float (&a)[Shape<IN, H1, H2>::kInP], int i) {
    using S = Shape<IN, H1, H2>;
    const float4* row = reinterpret_cast<const float4*>(sm + S::sW1 + i * S::kInP);
    float acc = sm[S::sB1 + i];
#pragma unroll
    for (int q = 0; q < S::kInP / 4; ++q) {
        const float4 w = row[q];
        acc += w.x * a[4 * q] + w.y * a[4 * q + 1] + w.z * a[4 * q + 2] + w.w * a[4 * q + 3];
    }
    return fmaxf(acc, 0.0f);
}
// z2[0..H2) += W2[:, i] * h (the column i of W2 is row i of the transposed copy in shared memory)
template <int IN, int H1, int H2>
__device__ __forceinline__ void feed2(const float* sm, int i, float h, float (&z2)[H2 > 0 ? H2 : 1]) {
    using S = Shape<IN, H1, H2>;
    const float4* col = reinterpret_cast<const float4*>(sm + S::sW2 + i * S::kW2P);
#pragma unroll
    for (int q = 0; q < H2 / 4; ++q) {
        const float4 v = col[q];
        z2[4 * q] += v.x * h; z2[4 * q + 1] += v.y * h; z2[4 * q + 2] += v.z * h; z2[4 * q + 3] += v.w * h;
    }
}

// The hidden units of the first layer are produced one at a time and consumed at once (second layer's sums, or the
// output): no 64-wide activation vector lives in registers, nothing is indexed dynamically but shared memory.
template <int IN, int H1, int H2>
__global__ void __launch_bounds__(kTile) k_edge_mlp_fwd(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                        int E, const float* __restrict__ x, int64_t x_bs,
                                                        const float* __restrict__ ea, int64_t ea_bs, Weights w,
                                                        float* __restrict__ out, int64_t out_bs, int64_t out_es) {
    using S = Shape<IN, H1, H2>;
    extern __shared__ float sm[];
    stage_weights<IN, H1, H2>(w, sm);
    __syncthreads();
    const int b = blockIdx.y;
    for (int e = blockIdx.x * kTile + threadIdx.x; e < E; e += gridDim.x * kTile) {
        float a[S::kInP];
        load_inputs<IN>(x, x_bs, ea, ea_bs, b, e, src[e], dst[e], a);
        float o = sm[S::sBo];
        if (H2 > 0) {
            float z2[H2 > 0 ? H2 : 1];
#pragma unroll
            for (int j = 0; j < H2; ++j) z2[j] = sm[S::sB2 + j];
#pragma unroll 2
            for (int i = 0; i < H1; ++i) feed2<IN, H1, H2>(sm, i, hidden1<IN, H1, H2>(sm, a, i), z2);
#pragma unroll
            for (int j = 0; j < H2; ++j) o += sm[S::sWo + j] * fmaxf(z2[j], 0.0f);
        } else {
#pragma unroll 2
            for (int i = 0; i < H1; ++i) o += sm[S::sWo + i] * hidden1<IN, H1, H2>(sm, a, i);
        }
        out[b * out_bs + e * out_es] = o;
    }
}

// ---------------------------------------------------------------------------------------------------- backward
// Shared memory of the backward kernel, per tile: A [128][kInP], H1 [128][H1], G1 [128][H1] and, with a second hidden
// layer, H2 / G2 [128][H2] each, plus g [128]. Row pitches are padded by 4 floats against bank conflicts of the
// row-per-thread writes.
template <int IN, int H1, int H2>
struct BwdSmem {
    using S = Shape<IN, H1, H2>;
    static constexpr int pA = S::kInP + 4, pH1 = H1 + 4, pH2 = (H2 > 0 ? H2 : 0) + 4;
    static constexpr int oA = S::kWeightFloats, oH1 = oA + kTile * pA, oG1 = oH1 + kTile * pH1,
                         oH2 = oG1 + kTile * pH1, oG2 = oH2 + (H2 > 0 ? kTile * pH2 : 0),
                         oG = oG2 + (H2 > 0 ? kTile * pH2 : 0), kFloats = oG + kTile;
};

// acc[r][c..c+3] += sum over the tile's pairs of L[p][r] * R[p][c..c+3]: work items (r, c/4) dealt round-robin to the
// CTA's threads, `kItems` per thread, accumulators in registers across tiles.
template <int ROWS, int COLS4, int kItems, int kThreadsB>
__device__ __forceinline__ void outer_sum(const float* __restrict__ L, int pL, const float* __restrict__ R, int pR,
                                          int live, float4 (&acc)[kItems]) {
    // the pair loop is the OUTER one: a thread's kItems accumulators are independent chains, and the shared-memory reads of
    // one pair are in flight together (item-outer order left one dependent FMA chain per warp: 44 us per tile)
    int offL[kItems], offR[kItems];
    bool on[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int item = threadIdx.x + k * kThreadsB;
        on[k] = item < ROWS * COLS4;
        const int r = on[k] ? item / COLS4 : 0, c4 = on[k] ? item - r * COLS4 : 0;
        offL[k] = r; offR[k] = 4 * c4;
    }
#pragma unroll 2
    for (int p = 0; p < live; ++p) {
        float l[kItems];
        float4 v[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            l[k] = L[p * pL + offL[k]];
            v[k] = *reinterpret_cast<const float4*>(R + p * pR + offR[k]);
        }
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            if (!on[k]) continue;
            acc[k].x += l[k] * v[k].x; acc[k].y += l[k] * v[k].y; acc[k].z += l[k] * v[k].z; acc[k].w += l[k] * v[k].w;
        }
    }
}

// acc[i][j] += sum over the tile's pairs of L[p][4 rb + i] * R[p][4 cb + j]: one 4 x 4 block per thread, two 128-bit
// shared-memory reads per 16 multiply-adds (the 1 x 4 form above reads two words per 4: the reductions of a tile then
// sit on the shared-memory pipe for ~18 us).
__device__ __forceinline__ void outer_block4(const float* __restrict__ L, int pL, const float* __restrict__ R, int pR,
                                             int rb, int cb, int live, float (&acc)[4][4]) {
#pragma unroll 2
    for (int p = 0; p < live; ++p) {
        const float4 l = *reinterpret_cast<const float4*>(L + p * pL + 4 * rb);
        const float4 v = *reinterpret_cast<const float4*>(R + p * pR + 4 * cb);
        const float lv[4] = {l.x, l.y, l.z, l.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += lv[i] * vv[j];
    }
}

// FOUR threads per pair in the recompute phase (sub-thread s owns the hidden units i = 4 ii + s; the second layer's
// sums are completed across the four by shuffles): 512 threads = 16 warps per SM on the one CTA the 145 KB of parked
// activations allow. (One thread per pair: 4 warps per SM, the recompute a bare latency chain — 110 ms at 8 rows x
// 6.0 M edges against 15 ms for the same arithmetic in the forward kernel's 12 CTAs per SM.)
constexpr int kBwdThreads = 4 * kTile;
template <int IN, int H1, int H2>
__global__ void __launch_bounds__(kBwdThreads) k_edge_mlp_bwd(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                              int E, int B, const float* __restrict__ x, int64_t x_bs,
                                                              const float* __restrict__ ea, int64_t ea_bs, Weights w,
                                                              const float* __restrict__ gout, int64_t g_bs, int64_t g_es,
                                                              float* __restrict__ partials) {
    using S = Shape<IN, H1, H2>;
    using M = BwdSmem<IN, H1, H2>;
    extern __shared__ float sm[];
    stage_weights<IN, H1, H2>(w, sm);
    float* sA = sm + M::oA;
    float* sH1 = sm + M::oH1;
    float* sG1 = sm + M::oG1;
    float* sH2 = sm + M::oH2;
    float* sG2 = sm + M::oG2;
    float* sG = sm + M::oG;
    constexpr int kOwn = H1 / 4;                                  // hidden units per sub-thread
    // weight-gradient blocks: thread t < kW1Blocks owns the 4 x 4 block (t / C1, t % C1) of dW1 (padded to kInP columns);
    // threads [kW2First, kW2First + kW2Blocks) own the blocks of dW2 — different warps, so both sets run side by side
    constexpr int C1 = S::kInP / 4, kW1Blocks = (H1 / 4) * C1;
    constexpr int C2 = H1 / 4, kW2Blocks = H2 > 0 ? (H2 / 4) * C2 : 0;
    constexpr int kW2First = (kW1Blocks + 31) / 32 * 32;
    static_assert(kW2First + kW2Blocks <= kBwdThreads, "not enough threads for the weight-gradient blocks");
    float accW[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accW[i][j] = 0.0f;
    const int tid = threadIdx.x;
    const bool ownW1 = tid < kW1Blocks, ownW2 = H2 > 0 && tid >= kW2First && tid < kW2First + kW2Blocks;
    float accB1 = 0.0f, accB2 = 0.0f, accWo = 0.0f, accBo = 0.0f;    // thread t: b1[t], b2[t], w_out[t]; thread 0: b_out
    const int tiles_per_row = (E + kTile - 1) / kTile;
    const long long n_tiles = (long long)tiles_per_row * B;
    const int p = threadIdx.x >> 2, sub = threadIdx.x & 3;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_row);
        const int e0 = (int)(tile - (long long)b * tiles_per_row) * kTile;
        const int live = min(kTile, E - e0);
        __syncthreads();                                          // weights staged / previous tile's reductions done
        const unsigned quad = __ballot_sync(0xffffffffu, p < live);   // the four sub-threads of a pair branch together
        if (p < live) {
            const int e = e0 + p;
            float a[S::kInP];
            load_inputs<IN>(x, x_bs, ea, ea_bs, b, e, src[e], dst[e], a);
            const float g = gout[b * g_bs + e * g_es];
            if (sub == 0) sG[p] = g;
#pragma unroll
            for (int c = 0; c < S::kInP; ++c)
                if ((c & 3) == sub) sA[p * M::pA + c] = (S::kInP > IN && c == IN) ? 1.0f : a[c];   // spare column: carries d b1
            if (H2 > 0) {
                float z2[H2 > 0 ? H2 : 1];
#pragma unroll
                for (int j = 0; j < H2; ++j) z2[j] = 0.0f;
#pragma unroll 2
                for (int ii = 0; ii < kOwn; ++ii) {
                    const int i = 4 * ii + sub;
                    const float h = hidden1<IN, H1, H2>(sm, a, i);
                    sH1[p * M::pH1 + i] = h;
                    feed2<IN, H1, H2>(sm, i, h, z2);
                }
#pragma unroll
                for (int j = 0; j < H2; ++j) {
                    float z = z2[j];
                    z += __shfl_xor_sync(quad, z, 1);
                    z += __shfl_xor_sync(quad, z, 2);
                    z += sm[S::sB2 + j];
                    const float g2 = z > 0.0f ? g * sm[S::sWo + j] : 0.0f;
                    if ((j & 3) == sub) { sH2[p * M::pH2 + j] = fmaxf(z, 0.0f); sG2[p * M::pH2 + j] = g2; }
                    z2[j] = g2;
                }
#pragma unroll 2
                for (int ii = 0; ii < kOwn; ++ii) {
                    const int i = 4 * ii + sub;
                    const float4* col = reinterpret_cast<const float4*>(sm + S::sW2 + i * S::kW2P);
                    float acc = 0.0f;
#pragma unroll
                    for (int q = 0; q < H2 / 4; ++q) {
                        const float4 v = col[q];
                        acc += v.x * z2[4 * q] + v.y * z2[4 * q + 1] + v.z * z2[4 * q + 2] + v.w * z2[4 * q + 3];
                    }
                    sG1[p * M::pH1 + i] = sH1[p * M::pH1 + i] > 0.0f ? acc : 0.0f;     // (its own store: same thread)
                }
            } else {
#pragma unroll
                for (int ii = 0; ii < kOwn; ++ii) {
                    const int i = 4 * ii + sub;
                    const float h = hidden1<IN, H1, H2>(sm, a, i);
                    sH1[p * M::pH1 + i] = h;
                    sG1[p * M::pH1 + i] = h > 0.0f ? g * sm[S::sWo + i] : 0.0f;
                }
            }
        }
        __syncthreads();
        // reductions over the tile's pairs (ascending pair order: fixed)
        if (ownW1) outer_block4(sG1, M::pH1, sA, M::pA, tid / C1, tid % C1, live, accW);
        else if (ownW2) outer_block4(sG2, M::pH2, sH1, M::pH1, (tid - kW2First) / C2, (tid - kW2First) % C2, live, accW);
        const int t = threadIdx.x;
        if (S::kInP == IN && t < H1) {                            // (otherwise the spare column of A carries it)
            float s = 0.0f;
            for (int q = 0; q < live; ++q) s += sG1[q * M::pH1 + t];
            accB1 += s;
        }
        if (H2 > 0) {
            if (t < H2) {
                float s2 = 0.0f, so = 0.0f;
                for (int q = 0; q < live; ++q) { s2 += sG2[q * M::pH2 + t]; so += sG[q] * sH2[q * M::pH2 + t]; }
                accB2 += s2; accWo += so;
            }
        } else if (t < H1) {
            float so = 0.0f;
            for (int q = 0; q < live; ++q) so += sG[q] * sH1[q * M::pH1 + t];
            accWo += so;
        }
        if (t == 0) {
            float s = 0.0f;
            for (int q = 0; q < live; ++q) s += sG[q];
            accBo += s;
        }
    }
    // this CTA's partial gradient vector
    float* out = partials + (size_t)blockIdx.x * S::kParams;
    if (ownW1) {
        const int rb = tid / C1, cb = tid % C1;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * rb + i, c = 4 * cb + j;
                if (c < IN) out[S::oW1 + r * IN + c] = accW[i][j];
                else if (c == IN) out[S::oB1 + r] = accW[i][j];                  // the spare column of A carried d b1
            }
    } else if (ownW2) {
        const int rb = (tid - kW2First) / C2, cb = (tid - kW2First) % C2;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) out[S::oW2 + (4 * rb + i) * H1 + 4 * cb + j] = accW[i][j];
    }
    const int t = threadIdx.x;
    if (S::kInP == IN && t < H1) out[S::oB1 + t] = accB1;
    if (H2 > 0 && t < H2) out[S::oB2 + t] = accB2;
    if (t < S::kLast) out[S::oWo + t] = accWo;
    if (t == 0) out[S::oBo] = accBo;
}

// grads[j] = sum over the CTAs' partial vectors, ascending CTA order
__global__ void __launch_bounds__(256) k_edge_mlp_finish(const float* __restrict__ partials, int n_parts, int n_params,
                                                         float* __restrict__ grads) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_params) return;
    float s = 0.0f;
    for (int i = 0; i < n_parts; ++i) s += partials[(size_t)i * n_params + j];
    grads[j] = s;
}

constexpr int kBwdCtas = 148;          // one CTA per SM: a tile's activations and gradients take 145 KB of shared memory

template <int IN, int H1, int H2>
int run_forward(const int32_t* src, const int32_t* dst, int E, int B, const float* x, int64_t x_bs, const float* ea,
                int64_t ea_bs, const Weights& w, float* out, int64_t out_bs, int64_t out_es, cudaStream_t s) {
    using S = Shape<IN, H1, H2>;
    const int bytes = S::kWeightFloats * (int)sizeof(float);
    const int gx = (E + kTile - 1) / kTile;
    k_edge_mlp_fwd<IN, H1, H2><<<dim3(gx < 4096 ? gx : 4096, B), kTile, bytes, s>>>(src, dst, E, x, x_bs, ea, ea_bs, w, out,
                                                                                  out_bs, out_es);
    return launch_status();
}

template <int IN, int H1, int H2>
int run_backward(const int32_t* src, const int32_t* dst, int E, int B, const float* x, int64_t x_bs, const float* ea,
                 int64_t ea_bs, const Weights& w, const float* gout, int64_t g_bs, int64_t g_es, float* partials,
                 float* grads, cudaStream_t s) {
    using S = Shape<IN, H1, H2>;
    using M = BwdSmem<IN, H1, H2>;
    const int bytes = M::kFloats * (int)sizeof(float);
    static const bool ok =
        cudaFuncSetAttribute(k_edge_mlp_bwd<IN, H1, H2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
    if (!ok) return TARL_E_LAUNCH;
    k_edge_mlp_bwd<IN, H1, H2><<<kBwdCtas, kBwdThreads, bytes, s>>>(src, dst, E, B, x, x_bs, ea, ea_bs, w, gout, g_bs, g_es, partials);
    k_edge_mlp_finish<<<(S::kParams + 255) / 256, 256, 0, s>>>(partials, kBwdCtas, S::kParams, grads);
    return launch_status();
}

bool args_ok(int32_t variant, const int32_t* src, const int32_t* dst, int32_t E, int32_t B, int32_t N, const float* x,
             const float* ea, const float* const* w) {
    if (variant != TARL_EDGE_MLP && variant != TARL_EDGE_MLP_TEST) return false;
    if (E < 0 || B < 0 || N < 0) return false;
    if (E == 0 || B == 0) return true;
    if (!src || !dst || !x || N == 0) return false;
    if (variant == TARL_EDGE_MLP && !ea) return false;
    const int n = variant == TARL_EDGE_MLP ? 6 : 4;
    for (int i = 0; i < n; ++i)
        if (w[i] == nullptr) return false;
    return (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

}  // namespace

extern "C" {

int32_t tarl_edge_mlp_param_count(int32_t variant) {
    return variant == TARL_EDGE_MLP ? Shape<33, 64, 32>::kParams : (variant == TARL_EDGE_MLP_TEST ? Shape<32, 16, 0>::kParams : 0);
}
int32_t tarl_edge_mlp_partial_count(void) { return kBwdCtas; }

int tarl_edge_mlp_inputs(const float* node_features, int64_t nf_batch_stride, int64_t nf_row_stride,
                         const int64_t* agent_index, const float* agent_features, int32_t agent_rows, int32_t batch,
                         int32_t n_nodes, float* x, int32_t* flags, void* stream) {
    if (batch < 0 || n_nodes < 0 || agent_rows < 1) return TARL_E_BADARG;
    if (batch == 0 || n_nodes == 0) return TARL_OK;
    if (!node_features || !agent_index || !agent_features || !x || !flags || (reinterpret_cast<uintptr_t>(x) & 15) != 0)
        return TARL_E_BADARG;
    const int64_t total = (int64_t)batch * n_nodes;
    k_edge_mlp_x<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        node_features, nf_batch_stride, nf_row_stride, reinterpret_cast<const long long*>(agent_index), agent_features,
        agent_rows, batch, n_nodes, x, flags);
    return launch_status();
}

int tarl_edge_mlp_forward(int32_t variant, const int32_t* edge_src, const int32_t* edge_dst, int32_t n_edges,
                          const float* x, int32_t batch, int32_t n_nodes, const float* edge_attr, int64_t ea_batch_stride,
                          const float* const* weights, int32_t use_tensor_cores, float* tc_scratch, float* out,
                          int64_t out_batch_stride, int64_t out_edge_stride, void* stream) {
    if (weights == nullptr || !args_ok(variant, edge_src, edge_dst, n_edges, batch, n_nodes, x, edge_attr, weights))
        return TARL_E_BADARG;
    if (n_edges == 0 || batch == 0) return TARL_OK;
    if (out == nullptr) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t x_bs = (int64_t)n_nodes * kX;
    if (variant == TARL_EDGE_MLP) {
        if (use_tensor_cores != 0 && tarl_edge_mlp_tc_available() != 0)
            return tarl::edge_mlp_forward_tc(edge_src, edge_dst, n_edges, batch, x, x_bs, edge_attr, ea_batch_stride, weights,
                                             tc_scratch, out, out_batch_stride, out_edge_stride, s);
        const Weights w = {weights[0], weights[1], weights[2], weights[3], weights[4], weights[5]};
        return run_forward<33, 64, 32>(edge_src, edge_dst, n_edges, batch, x, x_bs, edge_attr, ea_batch_stride, w, out,
                                       out_batch_stride, out_edge_stride, s);
    }
    const Weights w = {weights[0], weights[1], nullptr, nullptr, weights[2], weights[3]};
    return run_forward<32, 16, 0>(edge_src, edge_dst, n_edges, batch, x, x_bs, nullptr, 0, w, out, out_batch_stride,
                                  out_edge_stride, s);
}

int tarl_edge_mlp_backward(int32_t variant, const int32_t* edge_src, const int32_t* edge_dst, int32_t n_edges,
                           const float* x, int32_t batch, int32_t n_nodes, const float* edge_attr, int64_t ea_batch_stride,
                           const float* const* weights, const float* grad_out, int64_t go_batch_stride,
                           int64_t go_edge_stride, float* partials, float* grads, void* stream) {
    if (weights == nullptr || !args_ok(variant, edge_src, edge_dst, n_edges, batch, n_nodes, x, edge_attr, weights) ||
        grads == nullptr)
        return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_params = tarl_edge_mlp_param_count(variant);
    if (n_edges == 0 || batch == 0)
        return cudaMemsetAsync(grads, 0, sizeof(float) * n_params, s) == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    if (grad_out == nullptr || partials == nullptr) return TARL_E_BADARG;
    const int64_t x_bs = (int64_t)n_nodes * kX;
    if (variant == TARL_EDGE_MLP) {
        const Weights w = {weights[0], weights[1], weights[2], weights[3], weights[4], weights[5]};
        return run_backward<33, 64, 32>(edge_src, edge_dst, n_edges, batch, x, x_bs, edge_attr, ea_batch_stride, w,
                                        grad_out, go_batch_stride, go_edge_stride, partials, grads, s);
    }
    const Weights w = {weights[0], weights[1], nullptr, nullptr, weights[2], weights[3]};
    return run_backward<32, 16, 0>(edge_src, edge_dst, n_edges, batch, x, x_bs, nullptr, 0, w, grad_out, go_batch_stride,
                                   go_edge_stride, partials, grads, s);
}

}  // extern "C"

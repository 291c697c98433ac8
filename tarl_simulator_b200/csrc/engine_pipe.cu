// engine_pipe.cu — the pipelined step kernels of the resident link store (sm_100a). Compiled with -fmad=false.
//
// Same arithmetic, in the same order, as the kernels in engine.cu (results are bit-identical; semantics: SURVEY.md
// Appendix A = /root/reference/src/direction_mpnn.py:44-196 + src/response_mpnn.py:42-127). What changes is how the
// bytes arrive. The per-link kernels were latency-bound (profiles/r01_e, r01_f: ~2.5 TB/s at 40-60 % issue utilisation,
// long-scoreboard and barrier stalls): every CTA first waited for its link records, then for its edge lists, then for
// its gathers. Here a persistent CTA walks tiles of 256 consecutive links and the bulk-copy engine (cp.async.bulk,
// completion on an mbarrier) streams the NEXT tile's contiguous inputs — link records, statics, CSR pointers and the
// tile's edge range, which is contiguous because both CSR orientations are sorted by owner link — into shared memory
// while the current tile is being computed. Only the gathers of neighbouring links' records (L1/L2 hits for any
// sensible link numbering) remain as ordinary loads.
#include "engine_common.cuh"

using namespace tarl;

namespace {

constexpr int kTile = 256;
constexpr int kCapP = 1536;               // staged dual edges per tile (average degree 6); larger tiles read them directly
constexpr int kEdgeSlots = kCapP + 8;     // + alignment slack on both ends
constexpr int kStagesA = 2, kStagesB = 3;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Spin on the barrier's phase; a copy that never completes traps (an error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();
    }
}

struct TileInfo {
    int r, d0;
    bool bulk;   // full tile whose padded pointer copy stays inside the CSR pointer array
};
__device__ __forceinline__ TileInfo tile_info(int T, int tiles_per_rep, int N) {
    TileInfo ti;
    ti.r = T / tiles_per_rep;
    ti.d0 = (T - ti.r * tiles_per_rep) * kTile;
    ti.bulk = ti.d0 + kTile + 4 <= N + 1;
    return ti;
}
// the 16-byte aligned part [a0, a1) of the tile's edge range [e0, e1) that is staged by bulk copy
__device__ __forceinline__ void staged_range(int e0, int e1, int E, int* a0, int* a1) {
    *a0 = e0 & ~3;
    int hi = (e1 + 3) & ~3;
    const int lim = E & ~3;
    if (hi > lim) hi = lim;
    if (e1 - e0 > kCapP || hi < *a0) hi = *a0;
    *a1 = hi;
}

// ------------------------------------------------------------------------------------------------ direction phase
struct __align__(128) StageA {
    float4 hot[2 * kTile];   // the tile's own PRE-step records
    float4 stat[kTile];      // {FFTT, cc, ROAD_INDEX, MAXN}
    int ptr[kTile + 4];      // in_ptr[d0 .. d0+256] (+3 padding)
    int src[kEdgeSlots];     // in_src over the staged range; overwritten in place by the candidates' head ids
    float attr[kEdgeSlots];  // edge_attr over the staged range; overwritten in place by the eligibility product p
};
static_assert(sizeof(StageA) % 128 == 0, "stage size must keep the next stage aligned");

__device__ __forceinline__ void issue_stage_a(StageA& st, uint64_t* bar, const tarl_dual_csr& g, const Store& s,
                                              const float* attr_in, const TileInfo& ti, int e0, int e1) {
    int a0, a1;
    staged_range(e0, e1, g.n_edges, &a0, &a1);
    const uint32_t eb = (uint32_t)(a1 - a0) * 4u;
    mbar_expect_tx(bar, (uint32_t)(sizeof(st.hot) + sizeof(st.stat) + sizeof(st.ptr)) + 2u * eb);
    bulk_g2s(st.hot, s.hot_cur + 2 * ((size_t)ti.r * s.N + ti.d0), sizeof(st.hot), bar);
    bulk_g2s(st.stat, s.stat_a + ti.d0, sizeof(st.stat), bar);
    bulk_g2s(st.ptr, g.in_ptr + ti.d0, sizeof(st.ptr), bar);
    if (eb) {
        bulk_g2s(st.src, g.in_src + a0, eb, bar);
        bulk_g2s(st.attr, attr_in + a0, eb, bar);
    }
}

template <bool kExtNoise>
__global__ void __launch_bounds__(kTile, 4) k_pipe_select_append(
    tarl_dual_csr g, Store s, const float* __restrict__ attr_in, const float* __restrict__ noise, uint32_t seed_lo,
    uint32_t seed_hi, uint32_t step_id, float t, int32_t* __restrict__ flags, int tiles_per_rep, int total_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StageA* stages = reinterpret_cast<StageA*>(smem_raw);
    float* s_u = reinterpret_cast<float*>(smem_raw + kStagesA * sizeof(StageA));   // kExtNoise only
    __shared__ __align__(8) uint64_t full[kStagesA];
    __shared__ float s_room[kTile], s_ridx[kTile];
    __shared__ uint8_t s_free[kTile], s_list[kTile];
    __shared__ uint8_t s_owner[kEdgeSlots];
    __shared__ int s_count;

    const int tid = threadIdx.x;
    const int first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int j = 0; j < kStagesA; ++j) mbar_init(&full[j], 1);
        fence_barrier_init();
    }
    __syncthreads();

    int pe0 = 0, pe1 = 0;     // thread 0: edge range of the next tile to issue, fetched one iteration ahead
    if (tid == 0) {
        for (int j = 0; j < kStagesA - 1; ++j) {
            const int T = first + j * stride;
            if (T < total_tiles) {
                const TileInfo ti = tile_info(T, tiles_per_rep, s.N);
                if (ti.bulk) issue_stage_a(stages[j], &full[j], g, s, attr_in, ti, g.in_ptr[ti.d0], g.in_ptr[ti.d0 + kTile]);
            }
        }
        const int T = first + (kStagesA - 1) * stride;
        if (T < total_tiles) {
            const TileInfo ti = tile_info(T, tiles_per_rep, s.N);
            if (ti.bulk) { pe0 = g.in_ptr[ti.d0]; pe1 = g.in_ptr[ti.d0 + kTile]; }
        }
    }
    uint32_t parity = 0;      // bit j: phase of stage j's barrier

    for (int it = 0;; ++it) {
        const int T = first + it * stride;
        if (T >= total_tiles) break;
        const int stage = it % kStagesA;
        StageA& st = stages[stage];
        const TileInfo ti = tile_info(T, tiles_per_rep, s.N);

        if (tid == 0) {       // keep the pipeline full: the stage consumed in the previous iteration is free again
            const int Tn = first + (it + kStagesA - 1) * stride;
            if (Tn < total_tiles) {
                const TileInfo tn = tile_info(Tn, tiles_per_rep, s.N);
                const int sn = (it + kStagesA - 1) % kStagesA;
                if (tn.bulk) issue_stage_a(stages[sn], &full[sn], g, s, attr_in, tn, pe0, pe1);
            }
            const int Tnn = Tn + stride;
            if (Tnn < total_tiles) {
                const TileInfo tnn = tile_info(Tnn, tiles_per_rep, s.N);
                if (tnn.bulk) { pe0 = g.in_ptr[tnn.d0]; pe1 = g.in_ptr[tnn.d0 + kTile]; }
            }
        }

        const int d = ti.d0 + tid;
        const bool valid = d < s.N;
        const int base = ti.r * s.N;
        const int L = base + d;
        if (ti.bulk) {
            mbar_wait(&full[stage], (parity >> stage) & 1u);
            parity ^= 1u << stage;
        } else {              // ragged last tile of a replica: ordinary loads
            st.ptr[tid] = g.in_ptr[min(d, s.N)];
            if (tid == kTile - 1) st.ptr[kTile] = g.in_ptr[min(d + 1, s.N)];
            if (valid) {
                st.hot[2 * tid] = s.hot_cur[2 * (size_t)L];
                st.hot[2 * tid + 1] = s.hot_cur[2 * (size_t)L + 1];
                st.stat[tid] = s.stat_a[d];
            }
            __syncthreads();
        }

        float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, sa = h0;
        if (valid) { h0 = st.hot[2 * tid]; h1 = st.hot[2 * tid + 1]; sa = st.stat[tid]; }
        const float num = h1.x, maxn = h1.z, fftt = sa.x, ridx_d = sa.z;
        int meta = __float_as_int(h1.w);
        const bool bad = !(num >= 0.0f) || !(num < (float)s.Nmax);
        const bool free_d = num < (maxn - 3.0f);
        const float room_d = maxn - num;
        s_room[tid] = room_d;
        s_ridx[tid] = ridx_d;
        s_free[tid] = free_d ? 1 : 0;
        if (tid == 0) s_count = 0;

        const int e0 = st.ptr[0], e1 = st.ptr[kTile], ne = e1 - e0;
        const int kb = st.ptr[tid], ke = st.ptr[tid + 1];
        int a0, a1;
        staged_range(e0, e1, g.n_edges, &a0, &a1);
        if (!ti.bulk) a1 = a0;
        float best_id = 0.0f, psum = 0.0f;
        bool have = false;

        if (ne <= kCapP) {    // block-uniform
            for (int k = kb; k < ke; ++k) s_owner[k - a0] = (uint8_t)tid;
            __syncthreads();
            for (int k0 = e0 + tid; k0 < e1; k0 += 4 * kTile) {
                int u[4];
                float a[4], un[4];
                float4 U0[4], U1[4];
                float S[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = k0 + q * kTile;
                    u[q] = -1;
                    if (k < e1) {
                        if (k >= a0 && k < a1) {
                            u[q] = st.src[k - a0];
                            a[q] = st.attr[k - a0];
                        } else {
                            u[q] = g.in_src[k];
                            a[q] = attr_in[k];
                        }
                        if (kExtNoise) un[q] = noise[(int64_t)ti.r * g.n_edges + g.in_eid[k]];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (u[q] >= 0) {
                        const int Lu = base + u[q];
                        U0[q] = s.hot_cur[2 * Lu];
                        U1[q] = s.hot_cur[2 * Lu + 1];
                        S[q] = s.sel[Lu];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (u[q] >= 0) {
                        const int i = k0 + q * kTile - a0;
                        const int o = s_owner[i];
                        const bool a1m = (U0[q].z <= t) && (U1[q].x > 0.0f);
                        const bool a2m = ((U0[q].z - t) < -10.0f) && ((U1[q].z - 3.0f) <= U1[q].x);
                        const bool match = (S[q] == s_ridx[o]);
                        const bool m = (a1m && s_free[o] && match) || (a2m && ((U1[q].z - U1[q].x) <= s_room[o]) && match);
                        st.attr[i] = a[q] * (m ? 1.0f : 0.0f);
                        st.src[i] = __float_as_int(U0[q].x);
                        if (kExtNoise) s_u[i] = un[q];
                    }
                }
            }
            __syncthreads();
            for (int k = kb; k < ke; ++k) psum += st.attr[k - a0];
            // The scores only matter where somebody is eligible (src/direction_mpnn.py:142-144): compact those links
            // so that the three logf per candidate run on dense warps.
            if (psum > 0.0f) s_list[atomicAdd(&s_count, 1)] = (uint8_t)tid;
            __syncthreads();
            const int n_list = s_count;
            for (int w = tid; w < n_list; w += kTile) {
                const int o = s_list[w];
                const int ob = st.ptr[o], oe = st.ptr[o + 1];
                const uint32_t Lo = (uint32_t)(base + ti.d0 + o);
                float b = -FLT_MAX, bid = 0.0f;
                bool hv = false;
                float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
                for (int k = ob; k < oe; ++k) {
                    const int j = k - ob;
                    float uu;
                    if (kExtNoise) {
                        uu = s_u[k - a0];
                    } else {
                        if ((j & 3) == 0) philox4x32_10(Lo, 0u, step_id, (uint32_t)(j >> 2), seed_lo, seed_hi, un);
                        const int jj = j & 3;
                        uu = jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
                    }
                    const float sc = logf(st.attr[k - a0] + 1e-12f) + (-logf(-logf(uu)));
                    if (sc > b) { b = sc; bid = __int_as_float(st.src[k - a0]); hv = true; }
                }
                s_room[o] = bid;      // the downstream-side terms are no longer needed: reuse as result slots
                s_free[o] = hv ? 1 : 0;
            }
            __syncthreads();
            if (psum > 0.0f) { best_id = s_room[tid]; have = s_free[tid] != 0; }
        } else if (valid) {   // oversize tile: every link walks its own segment
            float best = -FLT_MAX;
            float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
            for (int k = kb; k < ke; ++k) {
                const int j = k - kb;
                const int Lu = base + g.in_src[k];
                const float4 u0 = s.hot_cur[2 * Lu], u1 = s.hot_cur[2 * Lu + 1];
                const float sel_u = s.sel[Lu];
                const bool a1m = (u0.z <= t) && (u1.x > 0.0f);
                const bool a2m = ((u0.z - t) < -10.0f) && ((u1.z - 3.0f) <= u1.x);
                const bool match = (sel_u == ridx_d);
                const bool m = (a1m && free_d && match) || (a2m && ((u1.z - u1.x) <= room_d) && match);
                const float p = attr_in[k] * (m ? 1.0f : 0.0f);
                psum += p;
                float uu;
                if (kExtNoise) {
                    uu = noise[(int64_t)ti.r * g.n_edges + g.in_eid[k]];
                } else {
                    if ((j & 3) == 0) philox4x32_10((uint32_t)L, 0u, step_id, (uint32_t)(j >> 2), seed_lo, seed_hi, un);
                    const int jj = j & 3;
                    uu = jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
                }
                const float sc = logf(p + 1e-12f) + (-logf(-logf(uu)));
                if (sc > best) { best = sc; best_id = u0.x; have = true; }
            }
        }

        if (valid) {
            float chosen = 0.0f;
            if (psum > 0.0f) {
                if (!have) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_NO_WINNER);
                else chosen = best_id;
            }
            const float dtt = max_propagate_nan((h0.z - h0.y) - fftt, 0.0f);
            float num_post = num, tail_post = h0.w, head_post = h0.x;
            if (bad) {
                atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
            } else {
                const int q = (int)num;
                const float dep_new = t + max_propagate_nan(fftt, sa.y / ((maxn + 10.0f) - num));
                if (q == 0) {                       // the tail slot IS the head slot
                    h0.x = chosen; h0.y = t; h0.z = dep_new;
                    head_post = chosen;
                    tail_post = chosen;
                    meta &= ~kMetaGarbage;
                    if (chosen != 0.0f) { num_post = num + 1.0f; h0.w = chosen; }
                } else if (chosen != 0.0f) {        // a real admission: one ring slot write
                    s.queue[(size_t)L * s.M + ring_pos(meta & kMetaRingMask, q, s.M)] = make_float4(chosen, t, dep_new, 0.0f);
                    num_post = num + 1.0f;
                    h0.w = chosen;
                    tail_post = chosen;
                    meta &= ~kMetaGarbage;
                } else {                            // the reference writes (0, t, t+tt) past the tail: keep it pending
                    meta |= kMetaGarbage;
                    h1.y = dep_new;
                }
                h1.x = num_post;
            }
            h1.w = __int_as_float(meta);
            s.hot_next[2 * (size_t)L] = h0;
            s.hot_next[2 * (size_t)L + 1] = h1;
            s.post[L] = make_float4(num_post, tail_post, head_post, dtt);
        }
        fence_proxy_async();   // this stage is refilled by the copy engine at the top of the next iteration
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ response phase
struct __align__(128) StageB {
    float4 post[kTile];      // the tile's own post-append summaries
    int ptr[kTile + 4];      // out_ptr[u0 .. u0+256] (+3 padding)
    int dst[kEdgeSlots];     // out_dst over the staged range
    int pad[20];
};
static_assert(sizeof(StageB) % 128 == 0, "stage size must keep the next stage aligned");

__device__ __forceinline__ void issue_stage_b(StageB& st, uint64_t* bar, const tarl_dual_csr& g, const Store& s,
                                              const TileInfo& ti, int e0, int e1) {
    int a0, a1;
    staged_range(e0, e1, g.n_edges, &a0, &a1);
    const uint32_t eb = (uint32_t)(a1 - a0) * 4u;
    mbar_expect_tx(bar, (uint32_t)(sizeof(st.post) + sizeof(st.ptr)) + eb);
    bulk_g2s(st.post, s.post + ((size_t)ti.r * s.N + ti.d0), sizeof(st.post), bar);
    bulk_g2s(st.ptr, g.out_ptr + ti.d0, sizeof(st.ptr), bar);
    if (eb) bulk_g2s(st.dst, g.out_dst + a0, eb, bar);
}

__global__ void __launch_bounds__(kTile) k_pipe_respond_pop(tarl_dual_csr g, Store s, float t,
                                                            float* __restrict__ delta_tt, uint8_t* __restrict__ pop,
                                                            int32_t* __restrict__ flags, int tiles_per_rep,
                                                            int total_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StageB* stages = reinterpret_cast<StageB*>(smem_raw);
    __shared__ __align__(8) uint64_t full[kStagesB];
    __shared__ float s_head[kTile], s_dtt[kTile];
    __shared__ uint8_t s_has[kTile], s_acc[kTile];
    __shared__ uint8_t s_owner[kEdgeSlots];

    const int tid = threadIdx.x;
    const int first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
        for (int j = 0; j < kStagesB; ++j) mbar_init(&full[j], 1);
        fence_barrier_init();
    }
    __syncthreads();

    int pe0 = 0, pe1 = 0;
    if (tid == 0) {
        for (int j = 0; j < kStagesB - 1; ++j) {
            const int T = first + j * stride;
            if (T < total_tiles) {
                const TileInfo ti = tile_info(T, tiles_per_rep, s.N);
                if (ti.bulk) issue_stage_b(stages[j], &full[j], g, s, ti, g.out_ptr[ti.d0], g.out_ptr[ti.d0 + kTile]);
            }
        }
        const int T = first + (kStagesB - 1) * stride;
        if (T < total_tiles) {
            const TileInfo ti = tile_info(T, tiles_per_rep, s.N);
            if (ti.bulk) { pe0 = g.out_ptr[ti.d0]; pe1 = g.out_ptr[ti.d0 + kTile]; }
        }
    }
    uint32_t parity = 0;

    for (int it = 0;; ++it) {
        const int T = first + it * stride;
        if (T >= total_tiles) break;
        const int stage = it % kStagesB;
        StageB& st = stages[stage];
        const TileInfo ti = tile_info(T, tiles_per_rep, s.N);

        if (tid == 0) {
            const int Tn = first + (it + kStagesB - 1) * stride;
            if (Tn < total_tiles) {
                const TileInfo tn = tile_info(Tn, tiles_per_rep, s.N);
                const int sn = (it + kStagesB - 1) % kStagesB;
                if (tn.bulk) issue_stage_b(stages[sn], &full[sn], g, s, tn, pe0, pe1);
            }
            const int Tnn = Tn + stride;
            if (Tnn < total_tiles) {
                const TileInfo tnn = tile_info(Tnn, tiles_per_rep, s.N);
                if (tnn.bulk) { pe0 = g.out_ptr[tnn.d0]; pe1 = g.out_ptr[tnn.d0 + kTile]; }
            }
        }

        const int u = ti.d0 + tid;
        const bool valid = u < s.N;
        const int base = ti.r * s.N;
        const int L = base + u;
        if (ti.bulk) {
            mbar_wait(&full[stage], (parity >> stage) & 1u);
            parity ^= 1u << stage;
        } else {
            st.ptr[tid] = g.out_ptr[min(u, s.N)];
            if (tid == kTile - 1) st.ptr[kTile] = g.out_ptr[min(u + 1, s.N)];
            if (valid) st.post[tid] = s.post[L];
            __syncthreads();
        }

        float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) P = st.post[tid];
        const bool has_up = at_least_one(P.x);
        // a link with agents may pop: fetch the second half of its record now, off the critical path of the edge phase
        float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && has_up) h1 = s.hot_next[2 * (size_t)L + 1];
        s_head[tid] = P.z;
        s_dtt[tid] = P.w;
        s_has[tid] = has_up ? 1 : 0;
        s_acc[tid] = 0;

        const int e0 = st.ptr[0], e1 = st.ptr[kTile], ne = e1 - e0;
        const int kb = st.ptr[tid], ke = st.ptr[tid + 1];
        int a0, a1;
        staged_range(e0, e1, g.n_edges, &a0, &a1);
        if (!ti.bulk) a1 = a0;
        float* dtt_out = (delta_tt != nullptr) ? delta_tt + (int64_t)ti.r * g.n_edges : nullptr;
        bool accept = false;

        if (ne <= kCapP) {
            for (int k = kb; k < ke; ++k) s_owner[k - a0] = (uint8_t)tid;
            __syncthreads();
            for (int k0 = e0 + tid; k0 < e1; k0 += 4 * kTile) {
                int dn[4], eid[4];
                float4 D[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = k0 + q * kTile;
                    dn[q] = -1;
                    if (k < e1) {
                        dn[q] = (k >= a0 && k < a1) ? st.dst[k - a0] : g.out_dst[k];
                        eid[q] = (g.out_eid != nullptr) ? g.out_eid[k] : k;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (dn[q] >= 0) D[q] = s.post[base + dn[q]];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (dn[q] >= 0) {
                        const int o = s_owner[k0 + q * kTile - a0];
                        if (dtt_out != nullptr) dtt_out[eid[q]] = s_dtt[o];
                        if (s_has[o] && at_least_one(D[q].x) && same_id(D[q].y, s_head[o])) s_acc[o] = 1;
                    }
                }
            }
            __syncthreads();
            accept = s_acc[tid] != 0;
        } else {
            __syncthreads();
            if (valid) {
                for (int k = kb; k < ke; ++k) {
                    if (dtt_out != nullptr) dtt_out[(g.out_eid != nullptr) ? g.out_eid[k] : k] = P.w;
                    const float4 D = s.post[base + g.out_dst[k]];
                    accept = accept || (has_up && at_least_one(D.x) && same_id(D.y, P.z));
                }
            }
        }
        if (valid) pop[L] = accept ? 1 : 0;
        accept = accept && valid;
        fence_proxy_async();
        if (__syncthreads_or(accept) && tid == 0) flags[TARL_FLAG_ANY_POP] = 1;   // also closes this stage's reads
        if (!accept) continue;

        // src/response_mpnn.py:119-122 as a ring-head increment. h0 need not be read: its tail id is post.y and the
        // head triplet is replaced.
        int meta = __float_as_int(h1.w);
        const int rh = meta & kMetaRingMask;
        const int M = s.M;
        const int q = (int)h1.x;                                    // >= 1 here
        const bool gv = meta & kMetaGarbage;
        const float4 garbage = make_float4(0.0f, t, h1.y, 0.0f);    // pending garbage was (re)written this very step
        float4* Q = s.queue + (size_t)L * M;
        const float4 new_head = (gv && q == 1) ? garbage : Q[rh];
        if (M > 1) {
            const float4 last = (gv && q == M) ? garbage : Q[ring_pos(rh, M, M)];
            Q[rh] = last;                                           // becomes logical slot M after the increment
        } else if (gv && q == 1) {
            Q[rh] = garbage;
        }
        h1.x = h1.x - 1.0f;
        int nrh = rh + 1; if (nrh >= M) nrh = 0;
        meta = (meta & ~kMetaRingMask) | nrh;
        if (gv && q == 1) meta &= ~kMetaGarbage;
        h1.w = __int_as_float(meta);
        s.hot_next[2 * (size_t)L] = make_float4(new_head.x, new_head.y, new_head.z, P.y);
        s.hot_next[2 * (size_t)L + 1] = h1;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename K>
int persistent_grid(K kernel, size_t dyn_smem, int total_tiles, int* grid) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return TARL_E_LAUNCH;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return TARL_E_LAUNCH;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem) != cudaSuccess) return TARL_E_LAUNCH;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTile, dyn_smem) != cudaSuccess || per_sm < 1)
        return TARL_E_LAUNCH;
    const int cap = sms * per_sm;
    *grid = total_tiles < cap ? total_tiles : cap;
    return TARL_OK;
}

}  // namespace

namespace tarl {

bool pipelined_step_supported(const tarl_dual_csr& g, const Store& s, const float* attr_in) {
    return aligned16(s.hot_cur) && aligned16(s.stat_a) && aligned16(s.post) && aligned16(g.in_ptr) &&
           aligned16(g.in_src) && aligned16(attr_in) && aligned16(g.out_ptr) && aligned16(g.out_dst) &&
           (int64_t)s.N * s.R * 2 < INT32_MAX;
}

int launch_pipelined_step(const tarl_dual_csr& g, const Store& s, const float* attr_in, const float* noise, uint64_t seed,
                          uint32_t step_id, float t, float* delta_tt, uint8_t* pop, int32_t* flags, cudaStream_t stream,
                          uint32_t phase_mask) {
    const int tiles_per_rep = (s.N + kTile - 1) / kTile;
    const int64_t total64 = (int64_t)tiles_per_rep * s.R;
    if (total64 >= INT32_MAX) return TARL_E_BADARG;
    const int total = (int)total64;
    int grid = 0, rc;
    if (phase_mask & TARL_PHASE_SELECT_APPEND) {
        if (noise != nullptr) {
            const size_t smem = kStagesA * sizeof(StageA) + kEdgeSlots * sizeof(float);
            if ((rc = persistent_grid(k_pipe_select_append<true>, smem, total, &grid)) != TARL_OK) return rc;
            k_pipe_select_append<true><<<grid, kTile, smem, stream>>>(g, s, attr_in, noise, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), step_id, t, flags,
                                                                      tiles_per_rep, total);
        } else {
            const size_t smem = kStagesA * sizeof(StageA);
            if ((rc = persistent_grid(k_pipe_select_append<false>, smem, total, &grid)) != TARL_OK) return rc;
            k_pipe_select_append<false><<<grid, kTile, smem, stream>>>(g, s, attr_in, noise, (uint32_t)seed,
                                                                       (uint32_t)(seed >> 32), step_id, t, flags,
                                                                       tiles_per_rep, total);
        }
    }
    if (phase_mask & TARL_PHASE_RESPOND_SHIFT) {
        const size_t smem = kStagesB * sizeof(StageB);
        if ((rc = persistent_grid(k_pipe_respond_pop, smem, total, &grid)) != TARL_OK) return rc;
        k_pipe_respond_pop<<<grid, kTile, smem, stream>>>(g, s, t, delta_tt, pop, flags, tiles_per_rep, total);
    }
    return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
}

}  // namespace tarl

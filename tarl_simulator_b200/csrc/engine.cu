// engine.cu — the resident link store: the same per-timestep network step as core_step.cu, on a compact device
// layout that moves ~1x the algorithmic bytes instead of ~4x (sm_100a). Compiled with -fmad=false.
//
// Why: on the reference's 208-byte AoS rows (Nmax=15) the step has to read every row whole, write 4 partial sectors
// per link for the mandatory tail write, and shift 3*(Nmax-1) floats per popped link (profiles/r01_a: 796 MB of DRAM
// traffic per step for 213 MB algorithmic at 1M links). The store keeps, per link,
//
//   hot   32 B  {head id, head arrival, head exit, tail id | NUM, pending-garbage exit time, MAXN, meta}
//               read once and rewritten whole every step (full-sector writes, ping-pong buffers: phase A gathers the
//               PRE-step records of upstream links while owners write POST-step records elsewhere)
//   sel    4 B  SELECTED_ROAD (the per-step routing input, its own array so that choice/actions write it coalesced)
//   queue 16 B x (Nmax-1) ring of {id, arrival, exit} for logical FIFO slots 1..Nmax-1 (slot 0 lives in `hot`),
//               touched only by real admissions (one slot write) and pops (one slot read + one slot copy)
//   post  16 B  {NUM, tail id, head id after the direction phase, delta_travel_time} for the response phase
//
// and reproduces the reference's x EXACTLY on export, including its quirks: the tail triplet (0, t, t+tt) it writes
// past the tail of every link every step is kept as one pending record per link ("garbage at logical slot int(NUM)",
// it is always overwritten in place by the next step or consumed by an export), and its shift-left that duplicates the
// last slot becomes a ring-head increment plus one slot copy. Semantics: SURVEY.md Appendix A.
#include "engine_common.cuh"

using namespace tarl;

namespace {

// ------------------------------------------------------------------------------------------------ import / export
__global__ void __launch_bounds__(kThreads) k_store_import(Store s, const float* __restrict__ x, int64_t row_stride,
                                                           int64_t rep_stride, const float* __restrict__ cc,
                                                           float4* __restrict__ hot, float4* __restrict__ stat_a,
                                                           float4* __restrict__ stat_b, int32_t* __restrict__ flags) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), n = (int)(L % s.N);
    const float* row = x + r * rep_stride + (int64_t)n * row_stride;
    const int Nmax = s.Nmax, c0 = 3 * Nmax;
    const float maxn = row[c0], num = row[c0 + 1], fftt = row[c0 + 2];
    const bool bad = !(num >= 0.0f) || !(num <= (float)Nmax);
    if (bad) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
    const int cnt = bad ? 0 : (int)num;
    hot[2 * L] = make_float4(row[0], row[Nmax], row[2 * Nmax], row[max(cnt - 1, 0)]);
    hot[2 * L + 1] = make_float4(num, 0.0f, maxn, __int_as_float(0));
    s.sel[L] = row[c0 + 5];
    for (int k = 1; k <= s.M; ++k)
        s.queue[L * s.M + (k - 1)] = make_float4(row[k], row[Nmax + k], row[2 * Nmax + k], 0.0f);
    if (r == 0) {
        float ccn;
        if (cc != nullptr) ccn = cc[n];
        else ccn = fftt * ((maxn + 10.0f) - (row[c0 + 4] * fftt) / 3600.0f);   // src/simulation_core_model.py:60-67
        stat_a[n] = make_float4(fftt, ccn, row[c0 + 6], maxn);
        stat_b[n] = make_float4(row[c0 + 3], row[c0 + 4], 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(kThreads) k_store_export(Store s, float* __restrict__ x, int64_t row_stride,
                                                           int64_t rep_stride, float t_garbage) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), n = (int)(L % s.N);
    float* row = x + r * rep_stride + (int64_t)n * row_stride;
    const int Nmax = s.Nmax, c0 = 3 * Nmax;
    const float4 h0 = s.hot_cur[2 * L], h1 = s.hot_cur[2 * L + 1];
    const int meta = __float_as_int(h1.w);
    const int rh = meta & kMetaRingMask;
    const int gslot = (meta & kMetaGarbage) ? (int)h1.x : -1;
    row[0] = h0.x; row[Nmax] = h0.y; row[2 * Nmax] = h0.z;
    for (int k = 1; k <= s.M; ++k) {
        float4 v = s.queue[L * s.M + ring_pos(rh, k, s.M)];
        if (k == gslot) v = make_float4(0.0f, t_garbage, h1.y, 0.0f);
        row[k] = v.x; row[Nmax + k] = v.y; row[2 * Nmax + k] = v.z;
    }
    const float4 a = s.stat_a[n], b = s.stat_b[n];
    row[c0] = h1.z; row[c0 + 1] = h1.x; row[c0 + 2] = a.x; row[c0 + 3] = b.x; row[c0 + 4] = b.y;
    row[c0 + 5] = s.sel[L]; row[c0 + 6] = a.z;
}

// ------------------------------------------------------------------------------------------------ direction phase
// One thread per (replica, downstream link d). Reads its own record and the PRE-step records of its upstream links,
// runs masks + Gumbel arg-max in ascending original edge id (src/direction_mpnn.py:74-99,133-144), then applies the
// tail write to its own record (:171-195) and publishes the post-append summary.
__global__ void __launch_bounds__(kThreads) k_store_select_append(
    tarl_dual_csr g, Store s, const float* __restrict__ attr_in, const float* __restrict__ noise, uint32_t seed_lo,
    uint32_t seed_hi, uint32_t step_id, float t, int32_t* __restrict__ flags) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), d = (int)(L % s.N);
    const int64_t base = (int64_t)r * s.N;
    float4 h0 = s.hot_cur[2 * L], h1 = s.hot_cur[2 * L + 1];
    const float4 st = s.stat_a[d];
    const float num = h1.x, maxn = h1.z, fftt = st.x, ridx_d = st.z;
    int meta = __float_as_int(h1.w);
    const bool bad = !(num >= 0.0f) || !(num < (float)s.Nmax);
    const bool free_d = num < (maxn - 3.0f);
    const float room_d = maxn - num;

    float best = -FLT_MAX, best_id = 0.0f, psum = 0.0f;
    bool have = false;
    const int k0 = g.in_ptr[d], k1 = g.in_ptr[d + 1];
    float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    for (int k = k0; k < k1; ++k) {
        const int j = k - k0;
        const int64_t Lu = base + g.in_src[k];
        const float4 u0 = s.hot_cur[2 * Lu], u1 = s.hot_cur[2 * Lu + 1];
        const float sel_u = s.sel[Lu];
        const bool a1 = (u0.z <= t) && (u1.x > 0.0f);
        const bool a2 = ((u0.z - t) < -10.0f) && ((u1.z - 3.0f) <= u1.x);
        const bool match = (sel_u == ridx_d);
        const bool m = (a1 && free_d && match) || (a2 && ((u1.z - u1.x) <= room_d) && match);
        const float p = attr_in[k] * (m ? 1.0f : 0.0f);
        psum += p;
        float uu;
        if (noise != nullptr) {
            uu = noise[(int64_t)r * g.n_edges + g.in_eid[k]];
        } else {
            if ((j & 3) == 0) philox4x32_10((uint32_t)L, (uint32_t)(L >> 32), step_id, (uint32_t)(j >> 2), seed_lo, seed_hi, un);
            const int jj = j & 3;
            uu = jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
        }
        const float sc = logf(p + 1e-12f) + (-logf(-logf(uu)));
        if (sc > best) { best = sc; best_id = u0.x; have = true; }
    }
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
        const float dep_new = t + max_propagate_nan(fftt, st.y / ((maxn + 10.0f) - num));
        if (q == 0) {                       // the tail slot IS the head slot
            h0.x = chosen; h0.y = t; h0.z = dep_new;
            head_post = chosen;
            tail_post = chosen;
            meta &= ~kMetaGarbage;
            if (chosen != 0.0f) { num_post = num + 1.0f; h0.w = chosen; }
        } else if (chosen != 0.0f) {        // a real admission: one ring slot write
            s.queue[L * s.M + ring_pos(meta & kMetaRingMask, q, s.M)] = make_float4(chosen, t, dep_new, 0.0f);
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
    s.hot_next[2 * L] = h0;
    s.hot_next[2 * L + 1] = h1;
    s.post[L] = make_float4(num_post, tail_post, head_post, dtt);
}

// ------------------------------------------------------------------------------------------------ response phase
// One thread per (replica, upstream link u): OR over out-edges of "tail(d) == head(u)" on the post-append summaries
// (src/response_mpnn.py:66-83), delta_travel_time for its out-edges, and the pop itself (:119-122) as a ring-head
// increment: new head <- logical slot 1, and the slot that becomes logical Nmax-1 <- old logical Nmax-1 (the
// reference's shift leaves the last slot in place, i.e. duplicates it).
__global__ void __launch_bounds__(kThreads) k_store_respond_pop(tarl_dual_csr g, Store s, float t,
                                                                float* __restrict__ delta_tt, uint8_t* __restrict__ pop,
                                                                int32_t* __restrict__ flags) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), u = (int)(L % s.N);
    const int64_t base = (int64_t)r * s.N;
    const float4 P = s.post[L];
    const bool has_up = (long long)P.x > 0;
    const long long head = (long long)P.z;
    bool accept = false;
    const int k1 = g.out_ptr[u + 1];
    for (int k = g.out_ptr[u]; k < k1; ++k) {
        if (delta_tt != nullptr) delta_tt[(int64_t)r * g.n_edges + (g.out_eid != nullptr ? g.out_eid[k] : k)] = P.w;
        const float4 D = s.post[base + g.out_dst[k]];
        accept = accept || (has_up && ((long long)D.x > 0) && ((long long)D.y == head));
    }
    pop[L] = accept ? 1 : 0;
    if (!accept) return;
    flags[TARL_FLAG_ANY_POP] = 1;
    float4 h0 = s.hot_next[2 * L], h1 = s.hot_next[2 * L + 1];
    int meta = __float_as_int(h1.w);
    const int rh = meta & kMetaRingMask;
    const int M = s.M;
    const int q = (int)h1.x;                                    // >= 1 here
    const bool gv = meta & kMetaGarbage;
    const float4 garbage = make_float4(0.0f, t, h1.y, 0.0f);    // pending garbage was (re)written this very step
    float4* Q = s.queue + L * M;
    const float4 new_head = (gv && q == 1) ? garbage : Q[rh];
    if (M > 1) {
        const float4 last = (gv && q == M) ? garbage : Q[ring_pos(rh, M, M)];
        Q[rh] = last;                                           // becomes logical slot M after the increment
    } else if (gv && q == 1) {
        Q[rh] = garbage;                                        // Nmax == 2: slot 1 keeps (a copy of) its value
    }
    h0.x = new_head.x; h0.y = new_head.y; h0.z = new_head.z;
    h1.x = h1.x - 1.0f;
    int nrh = rh + 1; if (nrh >= M) nrh = 0;
    meta = (meta & ~kMetaRingMask) | nrh;
    if (gv && q == 1) meta &= ~kMetaGarbage;                    // the garbage became the head slot
    h1.w = __int_as_float(meta);
    s.hot_next[2 * L] = h0;
    s.hot_next[2 * L + 1] = h1;
}


// ------------------------------------------------------------------------------------------------ tiled variants
// The kernels above walk each link's edge segment with one thread: in_ptr -> in_src -> hot[u] is a chain of dependent
// loads that is repeated once per edge, and the step ends up latency-bound (profiles/r01_c: 2.6 TB/s of DRAM traffic at
// 43 % occupancy). The tiled kernels give one CTA a tile of kTile consecutive links. Because both CSR orientations are
// sorted by their owner link, the tile's edges form ONE contiguous range: the CTA reads it edge-parallel (coalesced
// in_src / attr / out_dst streams, four independent upstream gathers in flight per thread), stages the per-edge result
// in shared memory, and only then runs the per-link scan in ascending edge id out of shared memory. Same arithmetic in
// the same order as the direct kernels, so the results are bit-identical. Tiles with more than kCap edges (average
// degree > 8) take the direct path.
constexpr int kTile = 256;
constexpr int kCap = 2048;

template <bool kExtNoise>
__global__ void __launch_bounds__(kTile, 4) k_tile_select_append(
    tarl_dual_csr g, Store s, const float* __restrict__ attr_in, const float* __restrict__ noise, uint32_t seed_lo,
    uint32_t seed_hi, uint32_t step_id, float t, int32_t* __restrict__ flags) {
    __shared__ int s_ptr[kTile + 1];
    __shared__ float s_room[kTile], s_ridx[kTile];
    __shared__ uint8_t s_free[kTile];
    __shared__ uint8_t s_owner[kCap];
    __shared__ float s_p[kCap], s_id[kCap];
    __shared__ float s_u[kExtNoise ? kCap : 1];
    __shared__ uint8_t s_list[kTile];
    __shared__ int s_count;

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int d = blockIdx.x * kTile + tid;
    if (tid == 0) s_count = 0;
    const bool valid = d < s.N;
    const int base = r * s.N;
    const int L = base + d;

    s_ptr[tid] = g.in_ptr[min(d, s.N)];
    if (tid == kTile - 1) s_ptr[kTile] = g.in_ptr[min(d + 1, s.N)];
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, st = h0;
    if (valid) {
        h0 = s.hot_cur[2 * L];
        h1 = s.hot_cur[2 * L + 1];
        st = s.stat_a[d];
    }
    const float num = h1.x, maxn = h1.z, fftt = st.x, ridx_d = st.z;
    int meta = __float_as_int(h1.w);
    const bool bad = !(num >= 0.0f) || !(num < (float)s.Nmax);
    const bool free_d = num < (maxn - 3.0f);
    const float room_d = maxn - num;
    s_room[tid] = room_d;
    s_ridx[tid] = ridx_d;
    s_free[tid] = free_d ? 1 : 0;
    __syncthreads();

    const int e0 = s_ptr[0], ne = s_ptr[kTile] - e0;
    const int kb = s_ptr[tid], ke = s_ptr[tid + 1];
    float best = -FLT_MAX, best_id = 0.0f, psum = 0.0f;
    bool have = false;

    if (ne <= kCap) {   // block-uniform
        for (int k = kb; k < ke; ++k) s_owner[k - e0] = (uint8_t)tid;
        __syncthreads();
        for (int i0 = tid; i0 < ne; i0 += 4 * kTile) {
            int u[4];
            float a[4], un[4];
            float4 U0[4], U1[4];
            float S[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * kTile;
                u[q] = -1;
                if (i < ne) {
                    u[q] = g.in_src[e0 + i];
                    a[q] = attr_in[e0 + i];
                    if (kExtNoise) un[q] = noise[(int64_t)r * g.n_edges + g.in_eid[e0 + i]];
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
                    const int i = i0 + q * kTile;
                    const int o = s_owner[i];
                    const bool a1 = (U0[q].z <= t) && (U1[q].x > 0.0f);
                    const bool a2 = ((U0[q].z - t) < -10.0f) && ((U1[q].z - 3.0f) <= U1[q].x);
                    const bool match = (S[q] == s_ridx[o]);
                    const bool m = (a1 && s_free[o] && match) || (a2 && ((U1[q].z - U1[q].x) <= s_room[o]) && match);
                    s_p[i] = a[q] * (m ? 1.0f : 0.0f);
                    s_id[i] = U0[q].x;
                    if (kExtNoise) s_u[i] = un[q];
                }
            }
        }
        __syncthreads();
        for (int k = kb; k < ke; ++k) psum += s_p[k - e0];
        // The scores only matter where somebody is eligible (src/direction_mpnn.py:142-144): compact those links so
        // that the three logf per candidate run on dense warps instead of on ~1 lane in 4.
        if (psum > 0.0f) s_list[atomicAdd(&s_count, 1)] = (uint8_t)tid;
        __syncthreads();
        const int n_list = s_count;
        for (int w = tid; w < n_list; w += kTile) {
            const int o = s_list[w];
            const int ob = s_ptr[o], oe = s_ptr[o + 1];
            const uint32_t Lo = (uint32_t)(L - tid + o);
            float b = -FLT_MAX, bid = 0.0f;
            bool hv = false;
            float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
            for (int k = ob; k < oe; ++k) {
                const int j = k - ob;
                float uu;
                if (kExtNoise) {
                    uu = s_u[k - e0];
                } else {
                    if ((j & 3) == 0) philox4x32_10(Lo, 0u, step_id, (uint32_t)(j >> 2), seed_lo, seed_hi, un);
                    const int jj = j & 3;
                    uu = jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
                }
                const float sc = logf(s_p[k - e0] + 1e-12f) + (-logf(-logf(uu)));
                if (sc > b) { b = sc; bid = s_id[k - e0]; hv = true; }
            }
            s_room[o] = bid;             // the downstream-side terms are no longer needed: reuse as the result slots
            s_free[o] = hv ? 1 : 0;
        }
        __syncthreads();
        if (psum > 0.0f) { best_id = s_room[tid]; have = s_free[tid] != 0; }
    } else if (valid) {
        float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
        for (int k = kb; k < ke; ++k) {
            const int j = k - kb;
            const int Lu = base + g.in_src[k];
            const float4 u0 = s.hot_cur[2 * Lu], u1 = s.hot_cur[2 * Lu + 1];
            const float sel_u = s.sel[Lu];
            const bool a1 = (u0.z <= t) && (u1.x > 0.0f);
            const bool a2 = ((u0.z - t) < -10.0f) && ((u1.z - 3.0f) <= u1.x);
            const bool match = (sel_u == ridx_d);
            const bool m = (a1 && free_d && match) || (a2 && ((u1.z - u1.x) <= room_d) && match);
            const float p = attr_in[k] * (m ? 1.0f : 0.0f);
            psum += p;
            float uu;
            if (kExtNoise) {
                uu = noise[(int64_t)r * g.n_edges + g.in_eid[k]];
            } else {
                if ((j & 3) == 0) philox4x32_10((uint32_t)L, 0u, step_id, (uint32_t)(j >> 2), seed_lo, seed_hi, un);
                const int jj = j & 3;
                uu = jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
            }
            const float sc = logf(p + 1e-12f) + (-logf(-logf(uu)));
            if (sc > best) { best = sc; best_id = u0.x; have = true; }
        }
    }
    if (!valid) return;

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
        const float dep_new = t + max_propagate_nan(fftt, st.y / ((maxn + 10.0f) - num));
        if (q == 0) {
            h0.x = chosen; h0.y = t; h0.z = dep_new;
            head_post = chosen;
            tail_post = chosen;
            meta &= ~kMetaGarbage;
            if (chosen != 0.0f) { num_post = num + 1.0f; h0.w = chosen; }
        } else if (chosen != 0.0f) {
            s.queue[(size_t)L * s.M + ring_pos(meta & kMetaRingMask, q, s.M)] = make_float4(chosen, t, dep_new, 0.0f);
            num_post = num + 1.0f;
            h0.w = chosen;
            tail_post = chosen;
            meta &= ~kMetaGarbage;
        } else {
            meta |= kMetaGarbage;
            h1.y = dep_new;
        }
        h1.x = num_post;
    }
    h1.w = __int_as_float(meta);
    s.hot_next[2 * L] = h0;
    s.hot_next[2 * L + 1] = h1;
    s.post[L] = make_float4(num_post, tail_post, head_post, dtt);
}


__global__ void __launch_bounds__(kTile) k_tile_respond_pop(tarl_dual_csr g, Store s, float t,
                                                            float* __restrict__ delta_tt, uint8_t* __restrict__ pop,
                                                            int32_t* __restrict__ flags) {
    __shared__ int s_ptr[kTile + 1];
    __shared__ float s_head[kTile], s_dtt[kTile];
    __shared__ uint8_t s_has[kTile], s_acc[kTile];
    __shared__ uint8_t s_owner[kCap];

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int u = blockIdx.x * kTile + tid;
    const bool valid = u < s.N;
    const int base = r * s.N;
    const int L = base + u;

    s_ptr[tid] = g.out_ptr[min(u, s.N)];
    if (tid == kTile - 1) s_ptr[kTile] = g.out_ptr[min(u + 1, s.N)];
    float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) P = s.post[L];
    const bool has_up = at_least_one(P.x);
    // a link with agents may pop: fetch the second half of its record now, off the critical path of the edge phase
    float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid && has_up) h1 = s.hot_next[2 * L + 1];
    s_head[tid] = P.z;
    s_dtt[tid] = P.w;
    s_has[tid] = has_up ? 1 : 0;
    s_acc[tid] = 0;
    __syncthreads();

    const int e0 = s_ptr[0], ne = s_ptr[kTile] - e0;
    const int kb = s_ptr[tid], ke = s_ptr[tid + 1];
    float* dtt_out = (delta_tt != nullptr) ? delta_tt + (int64_t)r * g.n_edges : nullptr;
    bool accept = false;

    if (ne <= kCap) {
        for (int k = kb; k < ke; ++k) s_owner[k - e0] = (uint8_t)tid;
        __syncthreads();
        for (int i0 = tid; i0 < ne; i0 += 4 * kTile) {
            int dn[4], eid[4];
            float4 D[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * kTile;
                dn[q] = -1;
                if (i < ne) {
                    dn[q] = g.out_dst[e0 + i];
                    eid[q] = (g.out_eid != nullptr) ? g.out_eid[e0 + i] : e0 + i;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (dn[q] >= 0) D[q] = s.post[base + dn[q]];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (dn[q] >= 0) {
                    const int o = s_owner[i0 + q * kTile];
                    if (dtt_out != nullptr) dtt_out[eid[q]] = s_dtt[o];
                    if (s_has[o] && at_least_one(D[q].x) && same_id(D[q].y, s_head[o])) s_acc[o] = 1;
                }
            }
        }
        __syncthreads();
        accept = s_acc[tid] != 0;
    } else if (valid) {
        for (int k = kb; k < ke; ++k) {
            if (dtt_out != nullptr) dtt_out[(g.out_eid != nullptr) ? g.out_eid[k] : k] = P.w;
            const float4 D = s.post[base + g.out_dst[k]];
            accept = accept || (has_up && at_least_one(D.x) && same_id(D.y, P.z));
        }
    }
    if (valid) pop[L] = accept ? 1 : 0;
    accept = accept && valid;
    if (__syncthreads_or(accept) && tid == 0) flags[TARL_FLAG_ANY_POP] = 1;
    if (!accept) return;

    // h1 was fetched up front; h0 need not be read at all: its tail id is post.y, the rest is replaced by the new head
    int meta = __float_as_int(h1.w);
    const int rh = meta & kMetaRingMask;
    const int M = s.M;
    const int q = (int)h1.x;
    const bool gv = meta & kMetaGarbage;
    const float4 garbage = make_float4(0.0f, t, h1.y, 0.0f);
    float4* Q = s.queue + (size_t)L * M;
    const float4 new_head = (gv && q == 1) ? garbage : Q[rh];
    if (M > 1) {
        const float4 last = (gv && q == M) ? garbage : Q[ring_pos(rh, M, M)];
        Q[rh] = last;
    } else if (gv && q == 1) {
        Q[rh] = garbage;
    }
    h1.x = h1.x - 1.0f;
    int nrh = rh + 1; if (nrh >= M) nrh = 0;
    meta = (meta & ~kMetaRingMask) | nrh;
    if (gv && q == 1) meta &= ~kMetaGarbage;
    h1.w = __int_as_float(meta);
    s.hot_next[2 * L] = make_float4(new_head.x, new_head.y, new_head.z, P.y);
    s.hot_next[2 * L + 1] = h1;
}

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

int make_store(const tarl_link_store* p, Store* s) {
    if (p == nullptr || p->n_links < 0 || p->n_replicas < 1 || p->nmax < 2 || p->nmax - 1 > kMetaRingMask) return TARL_E_BADARG;
    if (p->n_links > 0 && (!p->hot_cur || !p->hot_next || !p->sel || !p->stat_a || !p->stat_b || !p->queue || !p->post))
        return TARL_E_BADARG;
    s->N = p->n_links; s->R = p->n_replicas; s->Nmax = p->nmax; s->M = p->nmax - 1;
    s->hot_cur = static_cast<const float4*>(p->hot_cur);
    s->hot_next = static_cast<float4*>(p->hot_next);
    s->sel = static_cast<float*>(p->sel);
    s->stat_a = static_cast<const float4*>(p->stat_a);
    s->stat_b = static_cast<const float4*>(p->stat_b);
    s->queue = static_cast<float4*>(p->queue);
    s->post = static_cast<float4*>(p->post);
    return TARL_OK;
}

}  // namespace

extern "C" {

int tarl_store_import(const tarl_link_store* store, const float* x, int64_t x_row_stride, int64_t x_replica_stride,
                      const float* cc, int32_t* flags, void* stream) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if (s.N == 0) return TARL_OK;
    if (x == nullptr || flags == nullptr) return TARL_E_BADARG;
    k_store_import<<<blocks_for((int64_t)s.N * s.R), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        s, x, x_row_stride, x_replica_stride, cc, static_cast<float4*>(store->hot_cur),
        static_cast<float4*>(store->stat_a), static_cast<float4*>(store->stat_b), flags);
    return launch_status();
}

int tarl_store_export(const tarl_link_store* store, float* x, int64_t x_row_stride, int64_t x_replica_stride,
                      float t_last_step, void* stream) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if (s.N == 0) return TARL_OK;
    if (x == nullptr) return TARL_E_BADARG;
    k_store_export<<<blocks_for((int64_t)s.N * s.R), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        s, x, x_row_stride, x_replica_stride, t_last_step);
    return launch_status();
}

int tarl_store_step(const tarl_dual_csr* g, const tarl_link_store* store, const float* attr_in, const float* noise,
                    uint64_t seed, uint32_t step_id, float t, float* delta_tt, uint8_t* pop, int32_t* flags,
                    void* stream, uint32_t phase_mask) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if (g == nullptr || flags == nullptr || g->n_links != s.N) return TARL_E_BADARG;
    if (s.N == 0) return TARL_OK;
    if (pop == nullptr || (g->n_edges > 0 && attr_in == nullptr)) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (g->n_edges > 0 && (g->in_ptr == nullptr || g->in_src == nullptr || g->out_ptr == nullptr || g->out_dst == nullptr))
        return TARL_E_BADARG;
    if (noise != nullptr && g->n_edges > 0 && g->in_eid == nullptr) return TARL_E_BADARG;
    const uint32_t variant = (phase_mask >> TARL_STEP_VARIANT_SHIFT) & 0xfu;
    if (variant == TARL_STEP_VARIANT_PIPELINED && pipelined_step_supported(*g, s, attr_in))
        return launch_pipelined_step(*g, s, attr_in, noise, seed, step_id, t, delta_tt, pop, flags, cs, phase_mask);
    const bool tiled = variant != TARL_STEP_VARIANT_DIRECT && (int64_t)s.N * s.R * 2 < INT32_MAX && s.R <= 65535;
    if (tiled) {
        const dim3 grid((s.N + kTile - 1) / kTile, s.R);
        if (phase_mask & TARL_PHASE_SELECT_APPEND) {
            if (noise != nullptr)
                k_tile_select_append<true><<<grid, kTile, 0, cs>>>(*g, s, attr_in, noise, (uint32_t)seed,
                                                                   (uint32_t)(seed >> 32), step_id, t, flags);
            else
                k_tile_select_append<false><<<grid, kTile, 0, cs>>>(*g, s, attr_in, noise, (uint32_t)seed,
                                                                    (uint32_t)(seed >> 32), step_id, t, flags);
        }
        if (phase_mask & TARL_PHASE_RESPOND_SHIFT)
            k_tile_respond_pop<<<grid, kTile, 0, cs>>>(*g, s, t, delta_tt, pop, flags);
        return launch_status();
    }
    const int nb = blocks_for((int64_t)s.N * s.R);
    if (phase_mask & TARL_PHASE_SELECT_APPEND)
        k_store_select_append<<<nb, kThreads, 0, cs>>>(*g, s, attr_in, noise, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                       step_id, t, flags);
    if (phase_mask & TARL_PHASE_RESPOND_SHIFT)
        k_store_respond_pop<<<nb, kThreads, 0, cs>>>(*g, s, t, delta_tt, pop, flags);
    return launch_status();
}

}  // extern "C"

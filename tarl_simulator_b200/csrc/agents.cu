// agents.cu — the per-step population operations either side of the core step (sm_100a): insertion of departing
// agents, withdrawal of arrived agents, random route choice, application of an RL action, observation build.
//
// Reference semantics: /root/reference/src/agents/base.py:244-331 (insert_agent_into_network), :334-403
// (withdraw_agent_from_network), :446-494 (choice); /root/reference/src/reinforcement_learning.py:222-231 (action ->
// SELECTED_ROAD), :256-266 (reward), /root/reference/src/transportation_simulator.py:360-366 (state()).
// SURVEY.md Appendix B restates them. Compiled with -fmad=false (the exit-time arithmetic must round like ATen's).
//
// Every kernel exists for both state layouts through one accessor interface:
//   RowAcc    the reference's graph.x rows, in place ([R,] N_tot, 3*Nmax+7)            -> drop-in Agents methods
//   StoreAcc  the resident link store of engine.cu (hot records + ring queues)          -> batched rollouts
//
// Insertion without a sort: the reference sorts the ready agents by target road and admits, per road, the first
// min(count, MAXN-3-NUM) of them (ascending agent id inside a road: declared divergence D3, the reference's argsort is
// unstable). A ready agent's target road is x[ORIGIN(agent), SELECTED_ROAD], so all agents of one origin node target
// the same road; with the population indexed ONCE by origin (ascending agent id inside an origin) the per-road work is
// a k-way merge over the (almost always one) origins that currently select the road:
//   k_insert_offer  thread per (replica, origin with agents): pushes the origin on its road's list (atomicExch)
//   k_insert_admit  thread per (replica, origin at the head of its road's list): merges by smallest agent id until the
//                   room is used up
// The merge result does not depend on the order of the list, so the atomics leave no nondeterminism behind.
#include "engine_common.cuh"
#include "population.cuh"
#include "tile_map.cuh"

using namespace tarl;

namespace {


// ---------------------------------------------------------------------------------------------------------- accessors
struct RowAcc {
    float* x;
    int64_t row_stride, rep_stride;
    int N, Nmax;
    const float* cc;      // congestion_constant [>= N] or nullptr (then the congestion term is 0, base.py:318-319)
    float t_garbage;      // unused
    __device__ float* row(int r, int n) const { return x + r * rep_stride + (int64_t)n * row_stride; }
    __device__ float sel_of(int r, int node) const { return row(r, node)[3 * Nmax + 5]; }
    __device__ void set_sel(int r, int node, float v) const { row(r, node)[3 * Nmax + 5] = v; }
    __device__ float road_index(int r, int n) const { return row(r, n)[3 * Nmax + 6]; }

    struct Link {
        float* row;
        int Nmax;
        float num, maxn, fftt;
    };
    __device__ Link open(int r, int n) const {
        float* p = row(r, n);
        return {p, Nmax, p[3 * Nmax + 1], p[3 * Nmax], p[3 * Nmax + 2]};
    }
    __device__ static float slot_id(const Link& l, int k) { return l.row[k]; }
    __device__ static float slot_dep(const Link& l, int k) { return l.row[2 * l.Nmax + k]; }
    __device__ static void put(Link& l, int k, float id, float arr, float dep) {
        l.row[k] = id; l.row[l.Nmax + k] = arr; l.row[2 * l.Nmax + k] = dep;
    }
    __device__ static void commit_insert(Link& l, int admitted) { l.row[3 * l.Nmax + 1] = l.num + (float)admitted; }
    // src/agents/base.py:377-396: every segment shifts left by c with zero fill, NUM -= c
    __device__ static void withdraw(Link& l, int c) {
        for (int seg = 0; seg < 3; ++seg) {
            float* q = l.row + seg * l.Nmax;
            for (int k = 0; k < l.Nmax; ++k) q[k] = (k + c < l.Nmax) ? q[k + c] : 0.0f;
        }
        l.row[3 * l.Nmax + 1] = l.num - (float)c;
    }
};


__device__ __forceinline__ bool acc_has_cc(const RowAcc& a) { return a.cc != nullptr; }
__device__ __forceinline__ float acc_cc(const RowAcc& a, int n) { return a.cc[n]; }
__device__ __forceinline__ bool acc_has_cc(const StoreAcc&) { return true; }
__device__ __forceinline__ float acc_cc(const StoreAcc& a, int n) { return a.s.stat_a[a.s.slot_of(n)].y; }

// ------------------------------------------------------------------------------------------------------------ insert
__device__ __forceinline__ bool agent_ready(const AgentTable& at, int r, int a, float t) {
    const float* p = at.row(r, a);
    return p[kDepartureTime] <= t && p[kOnWay] == 0.0f && p[kDone] == 0.0f;     // base.py:247-251
}

// number of agents of origin node o whose DEPARTURE_TIME <= t (binary search in the origin's sorted departure times)
__device__ __forceinline__ int departed_by(const tarl_agent_index& ai, int o, float t) {
    int lo = ai.org_ptr[o], hi = ai.org_ptr[o + 1];
    const int k0 = lo;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ai.dep_sorted[mid] <= t) lo = mid + 1; else hi = mid;
    }
    return lo - k0;
}

__global__ void __launch_bounds__(kThreads) k_insert_departed(tarl_agent_index ai, float t, int32_t* __restrict__ departed) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < ai.n_origins) departed[i] = departed_by(ai, ai.origins[i], t);
}

// What one (replica, origin) pair does in the offer phase; returns true when the origin was pushed onto a road's list.
template <class Acc>
__device__ __forceinline__ bool insert_offer_one(const Acc& acc, const tarl_agent_index& ai, const AgentTable& at, float t,
                                                 int32_t* __restrict__ head, int32_t* __restrict__ next,
                                                 int32_t* __restrict__ cursor, int32_t* __restrict__ flags,
                                                 const int32_t* __restrict__ inserted,
                                                 const int32_t* __restrict__ departed, bool mark_unlisted, int r, int i) {
    const int o = ai.origins[i];
    const size_t ri = (size_t)r * ai.n_origins + i;
    if (inserted != nullptr && ai.dep_sorted != nullptr) {
        // How many of this origin's agents have departed by now (static, ascending departure times) against how many
        // this replica has inserted so far: equal = nobody is waiting, and neither the road's list nor the scan over
        // the origin's agent rows (the whole cost of an insertion step in steady state) is needed. The count does not
        // depend on the replica: with a `departed` scratch it was computed once per origin (k_insert_departed).
        const int n_dep = departed != nullptr ? departed[i] : departed_by(ai, o, t);
        // (next[ri] = -2 tells a per-(replica, origin) admit pass "not listed"; with a worklist only listed origins are
        // visited and the 4-byte store per pair and step — 40 MB at 1024 x 10 000 origins — is left out)
        if (n_dep <= inserted[ri]) { if (mark_unlisted) next[ri] = -2; return false; }
    }
    const long long road = (long long)acc.sel_of(r, o);                          // base.py:259
    cursor[ri] = ai.org_ptr[o];
    if (road < 0 || road >= acc.N) {          // not a road: harmless unless one of this origin's agents is ready, in
        bool any = false;                     // which case the reference would index a non-road row (or wrap around)
        for (int k = ai.org_ptr[o]; k < ai.org_ptr[o + 1] && !any; ++k) any = agent_ready(at, r, ai.org_agent[k], t);
        if (any) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_INSERT_TARGET);
        next[ri] = -2;
        return false;
    }
    next[ri] = atomicExch(&head[(size_t)r * acc.N + road], i);
    return true;
}

// worklist (optional): the listed origins of replica r are appended to work[r*n_origins + ..], their number kept in
// work_count[r] (zeroed by the launcher) — in steady state ~2 % of the origins are listed, spread one or two per warp,
// and the admit phase (a chain of dependent gathers per listed origin) runs over the compact list instead of over
// every (replica, origin) pair. One atomic per warp that lists anything (a counter per replica); the order inside the list is arbitrary and does not matter (each
// road is served by the one origin at the head of its list, merging by agent id).
template <class Acc>
__global__ void __launch_bounds__(kThreads) k_insert_offer(Acc acc, tarl_agent_index ai, AgentTable at, float t,
                                                           int32_t* __restrict__ head, int32_t* __restrict__ next,
                                                           int32_t* __restrict__ cursor, int32_t* __restrict__ flags,
                                                           const int32_t* __restrict__ inserted,
                                                           int32_t* __restrict__ work, int32_t* __restrict__ work_count,
                                                           const int32_t* __restrict__ departed) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    bool listed = false;
    if (i < ai.n_origins) listed = insert_offer_one(acc, ai, at, t, head, next, cursor, flags, inserted, departed, work == nullptr, r, i);
    if (work == nullptr) return;
    const unsigned m = __ballot_sync(0xffffffffu, listed);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&work_count[r], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (listed) work[(size_t)r * ai.n_origins + base + __popc(m & ((1u << lane) - 1u))] = i;
}

// What the origin i0 of replica r does in the admit phase: if it ended up at the HEAD of its road's list it serves the
// road. (A thread per road would launch N threads per replica to find the few roads with a list.)
// kDirect: the origin serves its road on its own — no list was built (k_insert_direct: networks in which a road can be
// selected by one origin only), the "list" is the origin itself.
template <bool kDirect, class Acc>
__device__ __forceinline__ void insert_admit_one(const Acc& acc, const tarl_agent_index& ai, const AgentTable& at, float t,
                                                 int32_t* __restrict__ head, const int32_t* __restrict__ next,
                                                 int32_t* __restrict__ cursor, int32_t* __restrict__ counters,
                                                 int32_t* __restrict__ inserted, const int32_t* __restrict__ departed,
                                                 float* __restrict__ num_out, int32_t* __restrict__ occupancy,
                                                 int n_nodes, int r, int i0) {
    if (i0 >= ai.n_origins) return;
    if (!kDirect && next[(size_t)r * ai.n_origins + i0] == -2) return;            // not listed this step
    const long long road = (long long)acc.sel_of(r, ai.origins[i0]);
    if (road < 0 || road >= acc.N) return;
    const int n = (int)road;
    const size_t L = (size_t)r * acc.N + n;
    const int h = kDirect ? i0 : head[L];
    if (h != i0) return;
    if (!kDirect) head[L] = -1;                                                  // ready for the next call
    typename Acc::Link l = acc.open(r, n);
    const long long cap = (long long)((l.maxn - 3.0f) - l.num);                  // base.py:262-267
    if (cap <= 0) return;
    if (!(l.num >= 0.0f)) return;
    const int q0 = (int)l.num;
    float tc = 0.0f;
    if (acc_has_cc(acc)) tc = acc_cc(acc, n) / ((l.maxn + 10.0f) - (float)q0);   // :314-319 (start_counts is a long)
    const float dep = t + max_propagate_nan(l.fftt, tc);                         // :321-325
    const int32_t* nx = next + (size_t)r * ai.n_origins;
    int32_t* cur = cursor + (size_t)r * ai.n_origins;
    int admitted = 0;
    while (admitted < cap && q0 + admitted < acc.Nmax) {
        int best_a = INT32_MAX, best_i = -1;
        for (int i = h; i >= 0; i = kDirect ? -1 : nx[i]) {
            // Nobody (left) waiting at this origin — departed by now = inserted so far, see insert_offer_one: the scan
            // below would walk the rest of the origin's agents, two dependent loads per four of them, to find nothing
            // (after the one agent an origin typically inserts in a step that walk was half of the thread's chain)
            if (inserted != nullptr && ai.dep_sorted != nullptr) {
                const int n_dep = departed != nullptr ? departed[i] : departed_by(ai, ai.origins[i], t);
                if (n_dep <= inserted[(size_t)r * ai.n_origins + i]) continue;
            }
            const int end = ai.org_ptr[ai.origins[i] + 1];
            int k = cur[i];
            // first ready agent at or after k. Four candidates per round: their ids and then their rows are loaded
            // together — two dependent loads per FOUR skipped agents instead of per agent (the walk over agents that
            // are already on their way or done is the longest dependent chain of an insertion step)
            while (k < end) {
                int cand[4];
                bool rdy[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) cand[q] = k + q < end ? ai.org_agent[k + q] : -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) rdy[q] = cand[q] >= 0 && agent_ready(at, r, cand[q], t);
                int first = 4;
#pragma unroll
                for (int q = 3; q >= 0; --q) first = rdy[q] ? q : first;
                k += first;
                if (first < 4) break;
            }
            if (k > end) k = end;
            cur[i] = k;
            if (k < end && ai.org_agent[k] < best_a) { best_a = ai.org_agent[k]; best_i = i; }
        }
        if (best_i < 0) break;
        Acc::put(l, q0 + admitted, (float)best_a, t, dep);                       // :310-325
        at.row(r, best_a)[kOnWay] = 1.0f;                                        // :328
        cur[best_i] += 1;
        if (inserted != nullptr) inserted[(size_t)r * ai.n_origins + best_i] += 1;   // an origin sits in ONE road's list
        ++admitted;
    }
    if (admitted > 0) {
        Acc::commit_insert(l, admitted);                                         // :327
        if (counters != nullptr) atomicAdd(&counters[2 * r], admitted);
        if (num_out != nullptr) {               // occupancy observation left behind by k_withdraw_observe: patch it
            num_out[(size_t)r * n_nodes + n] = l.num + (float)admitted;
            atomicAdd(&occupancy[r], admitted);
        }
    }
}

// One thread per (replica, origin) — or, with a worklist, a few CTAs per replica striding over its LISTED origins (a
// grid sized for every origin spends its time retiring thousands of CTAs that find their index beyond the count).
template <class Acc>
__global__ void __launch_bounds__(kThreads) k_insert_admit(Acc acc, tarl_agent_index ai, AgentTable at, float t,
                                                           int32_t* __restrict__ head, const int32_t* __restrict__ next,
                                                           int32_t* __restrict__ cursor, int32_t* __restrict__ counters,
                                                           int32_t* __restrict__ inserted, const int32_t* __restrict__ work,
                                                           const int32_t* __restrict__ work_count,
                                                           const int32_t* __restrict__ departed,
                                                           float* __restrict__ num_out, int32_t* __restrict__ occupancy,
                                                           int n_nodes) {
    const int r = blockIdx.y;
    if (work == nullptr) {
        insert_admit_one<false>(acc, ai, at, t, head, next, cursor, counters, inserted, departed, num_out, occupancy,
                                n_nodes, r, (int)(blockIdx.x * kThreads + threadIdx.x));
        return;
    }
    const int count = work_count[r];
    for (int w = blockIdx.x * kThreads + threadIdx.x; w < count; w += gridDim.x * kThreads)
        insert_admit_one<false>(acc, ai, at, t, head, next, cursor, counters, inserted, departed, num_out, occupancy,
                                n_nodes, r, work[(size_t)r * ai.n_origins + w]);
}

// Offer and admit phase in ONE pass, for networks in which every road can be selected by a single origin only
// (road_origin[n] = that origin's index, -1 = none: config_network's graphs, where a road leaves one intersection and
// that intersection's SRC node is the only node with an edge into it). No list, no second kernel, no grid-wide
// dependency between the two: one thread per (replica, origin); origins nobody waits at return after two loads, the
// others insert straight away. A SELECTED_ROAD that names a road of another origin is reported (sticky error flag)
// instead of being served: two threads could otherwise append to one queue.
// (Measured and rejected: a fast path that requests the road's record and the origin's first 16 agents with their rows
// all at once — four dependent load levels instead of ~ten — 43 -> 68 us per step of 128 grid100 replicas: the cost of an
// insertion is the NUMBER of scattered agent rows it touches (every one a TLB miss in a 0.5 GB table), not the depth of
// the chain; the chunked scan stops at the first ready agent.)
template <class Acc>
__global__ void __launch_bounds__(kThreads) k_insert_direct(Acc acc, tarl_agent_index ai, AgentTable at, float t,
                                                            const int32_t* __restrict__ road_origin,
                                                            int32_t* __restrict__ cursor, int32_t* __restrict__ counters,
                                                            int32_t* __restrict__ inserted,
                                                            const int32_t* __restrict__ departed, int32_t* __restrict__ flags,
                                                            float* __restrict__ num_out, int32_t* __restrict__ occupancy,
                                                            int n_nodes) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= ai.n_origins) return;
    const int o = ai.origins[i];
    const size_t ri = (size_t)r * ai.n_origins + i;
    const int n_dep = departed != nullptr ? departed[i] : departed_by(ai, o, t);
    if (n_dep <= inserted[ri]) return;                                            // nobody is waiting here
    const long long road = (long long)acc.sel_of(r, o);                          // base.py:259
    cursor[ri] = ai.org_ptr[o];
    if (road < 0 || road >= acc.N || road_origin[road] != i) {
        // not a road, or not this origin's road: harmless unless one of this origin's agents is ready (see insert_offer_one)
        bool any = false;
        for (int k = ai.org_ptr[o]; k < ai.org_ptr[o + 1] && !any; ++k) any = agent_ready(at, r, ai.org_agent[k], t);
        if (any) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_INSERT_TARGET);
        return;
    }
    insert_admit_one<true>(acc, ai, at, t, nullptr, nullptr, cursor, counters, inserted, departed, num_out, occupancy,
                           n_nodes, r, i);
}


template <class Acc>
__global__ void __launch_bounds__(kThreads) k_withdraw(Acc acc, AgentTable at, tarl_csr adj, float t,
                                                       uint8_t* __restrict__ mask, int32_t* __restrict__ counters,
                                                       int32_t* __restrict__ flags) {
    const int n = blockIdx.x * kThreads + threadIdx.x;
    if (n >= acc.N) return;
    const int r = blockIdx.y;
    withdraw_one(acc, at, adj, t, mask, counters, flags, r, n);
}

// The same pass also leaving the occupancy observation behind (rollouts whose nets read NUMBER_OF_AGENT only): this
// kernel has every link's record in registers anyway, so it writes num_out[r, n] = NUM after the withdrawal (0 for the
// non-road nodes: the grid covers n_nodes) and adds the NUMs into occupancy[r] (integer atomics, one per warp); the
// insertion that follows patches the few roads it touches (k_insert_admit). A separate observe pass would read all the
// records once more.
__global__ void __launch_bounds__(kThreads) k_withdraw_observe(StoreAcc acc, AgentTable at, tarl_csr adj, float t,
                                                               uint8_t* __restrict__ mask, int32_t* __restrict__ counters,
                                                               int32_t* __restrict__ flags, float* __restrict__ num_out,
                                                               int32_t* __restrict__ occupancy) {
    const int n = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    float num = 0.0f;
    if (n < acc.N) num = withdraw_one(acc, at, adj, t, mask, counters, flags, r, n);
    if (n < acc.n_nodes) num_out[(size_t)r * acc.n_nodes + n] = num;
    int num_i = (int)num;
    for (int off = 16; off > 0; off >>= 1) num_i += __shfl_xor_sync(0xffffffffu, num_i, off);
    if ((threadIdx.x & 31) == 0 && num_i != 0) atomicAdd(&occupancy[r], num_i);
}

// ------------------------------------------------------------------------------------------------------------ choice
// Every node with out-neighbours (roads -> roads, SRC nodes -> their outgoing roads) draws one of them uniformly
// (base.py:446-494; the reference samples torch.multinomial over dense 0/1 rows). Here: neighbour number
// min(floor(u*deg), deg-1) in ascending road id, with u injected per (replica, chooser) or drawn from Philox.
template <class Acc>
__global__ void __launch_bounds__(kThreads) k_choice(Acc acc, tarl_csr nbr, const int32_t* __restrict__ choosers,
                                                     int n_choosers, const float* __restrict__ uniforms,
                                                     uint32_t seed_lo, uint32_t seed_hi, uint32_t step_id) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_choosers) return;
    const int r = blockIdx.y;
    const int node = choosers[i];
    const int k0 = nbr.ptr[node], deg = nbr.ptr[node + 1] - k0;
    float u;
    if (uniforms != nullptr) {
        u = uniforms[(size_t)r * n_choosers + i];
    } else {
        float un[4];
        philox4x32_10((uint32_t)node, (uint32_t)r, step_id, 0x43484f49u, seed_lo, seed_hi, un);
        u = un[0];
    }
    int k = (int)(u * (float)deg);
    k = min(max(k, 0), deg - 1);
    acc.set_sel(r, node, (float)nbr.idx[k0 + k]);
}

// RL action: x[edge_index[0][e], SELECTED_ROAD] = edge_index[1][e] for every selected edge of the FULL graph
// (reinforcement_learning.py:223-231). action: [R, E_full] one-hot per source group, any strides; the thread order
// follows the action's memory order (replica innermost for an edge-major action) so that its bytes are read coalesced.
template <class Acc>
__global__ void __launch_bounds__(kThreads) k_apply_action(Acc acc, const int32_t* __restrict__ src,
                                                           const int32_t* __restrict__ dst, int E, int R,
                                                           const void* __restrict__ action, int64_t a_sr, int64_t a_se,
                                                           int action_dtype, bool replica_innermost) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)E * R) return;
    int e, r;
    if (replica_innermost) { e = (int)(i / R); r = (int)(i - (int64_t)e * R); }
    else { r = (int)(i / E); e = (int)(i - (int64_t)r * E); }
    const int64_t o = r * a_sr + e * a_se;
    bool on;
    switch (action_dtype) {
        case TARL_ACTION_U8: on = static_cast<const uint8_t*>(action)[o] != 0; break;
        case TARL_ACTION_I64: on = static_cast<const long long*>(action)[o] != 0; break;
        default: on = static_cast<const float*>(action)[o] != 0.0f; break;
    }
    if (on) acc.set_sel(r, src[e], (float)dst[e]);
}

// The same write with the edges grouped by source node (CSR of the full graph by source rank, what GraphDistribution
// samples over). The action is edge-major (replica innermost) while SELECTED_ROAD is node-major inside a replica, so a
// thread per (edge, replica) scatters its writes one sector each; here a CTA owns a tile of (source nodes x replicas)
// (tile_map.cuh): it scans each group's edges with the replica innermost (one 32-byte sector of action bytes per edge
// and 32 replicas), keeps the destination of the LAST selected edge in ascending edge id (what sequential index_put
// leaves behind), and writes with the node innermost. Groups without a selected edge are left untouched.
template <class Acc>
__global__ void __launch_bounds__(tarl::kTileThreads) k_apply_action_groups(Acc acc, tarl_csr grp,
                                                                            const int32_t* __restrict__ group_node,
                                                                            const int32_t* __restrict__ dst, int R, int Bp,
                                                                            const void* __restrict__ action, int64_t a_sr,
                                                                            int64_t a_se, int action_dtype) {
    __shared__ float sm[tarl::kTileSmem];
    const tarl::Tile t = tarl::tile_here(R, Bp);
    tarl::tile_walk_rows(t, [&](int rr, int j) {
        const int g = t.n0 + j, r = t.b0 + rr;
        float chosen = -1.0f;
        if (g < grp.n_rows && rr < t.nrows) {
            const int k1 = grp.ptr[g + 1];
            for (int k = grp.ptr[g]; k < k1; ++k) {
                const int e = grp.eid[k];
                const int64_t o = r * a_sr + e * a_se;
                bool on;
                switch (action_dtype) {
                    case TARL_ACTION_U8: on = static_cast<const uint8_t*>(action)[o] != 0; break;
                    case TARL_ACTION_I64: on = static_cast<const long long*>(action)[o] != 0; break;
                    default: on = static_cast<const float*>(action)[o] != 0.0f; break;
                }
                if (on) chosen = (float)dst[e];
            }
        }
        sm[tarl::tile_slot(t, rr, j)] = chosen;
    });
    __syncthreads();
    tarl::tile_walk_nodes(t, [&](int rr, int j) {
        const int g = t.n0 + j;
        if (g >= grp.n_rows || rr >= t.nrows) return;
        const float v = sm[tarl::tile_slot(t, rr, j)];
        if (v >= 0.0f) acc.set_sel(t.b0 + rr, group_node[g], v);
    });
}

// ------------------------------------------------------------------------------------------------------- observation
// state() of the reference on the store (transportation_simulator.py:360-366): node_features[r, n, 0:7] =
// {MAXN, NUM, FFTT, LENGTH, MAX_FLOW, SELECTED_ROAD, ROAD_INDEX}, agent_index[r, n] = head agent id; non-road nodes
// are {0,0,0,0,0,sel,-1} / 0. occupancy[r] += sum_n NUM (integer atomics: the reward is -occupancy, :266).
__global__ void __launch_bounds__(kThreads) k_observe(StoreAcc acc, float* __restrict__ node_features,
                                                      long long* __restrict__ agent_index,
                                                      int32_t* __restrict__ occupancy, float* __restrict__ num_agents,
                                                      float* __restrict__ selected_road) {
    const int n = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    int num_i = 0;
    if (n < acc.n_nodes) {
        float f[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, -1.0f};
        long long head = 0;
        if (n < acc.N) {
            const int slot = acc.s.slot_of(n);
            const float4* rec = reinterpret_cast<const float4*>(acc.hot) + 2 * ((size_t)r * acc.N + slot);
            const float4 A = rec[0];
            f[0] = A.w; f[1] = A.z;
            if (node_features != nullptr) {
                const float4 sa = acc.s.stat_a[slot], sb = acc.s.stat_b[slot];
                f[2] = sa.x; f[3] = sb.x; f[4] = sb.y; f[6] = sa.z;
            }
            head = (long long)A.x;
            num_i = (int)A.z;
        }
        const size_t o = (size_t)r * acc.n_nodes + n;
        if (node_features != nullptr || selected_road != nullptr) f[5] = acc.sel_of(r, n);
        if (node_features != nullptr)
            for (int c = 0; c < 7; ++c) node_features[o * 7 + c] = f[c];
        if (agent_index != nullptr) agent_index[o] = head;
        if (num_agents != nullptr) num_agents[o] = f[1];
        if (selected_road != nullptr) selected_road[o] = f[5];
    }
    if (occupancy != nullptr) {
        for (int off = 16; off > 0; off >>= 1) num_i += __shfl_xor_sync(0xffffffffu, num_i, off);
        if ((threadIdx.x & 31) == 0 && num_i != 0) atomicAdd(&occupancy[r], num_i);
    }
}

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

int check_state(const tarl_agent_state* st, RowAcc* row, StoreAcc* sto, bool* is_store, int* R) {
    if (st == nullptr) return TARL_E_BADARG;
    if ((st->x != nullptr) == (st->store != nullptr)) return TARL_E_BADARG;      // exactly one layout
    if (st->x != nullptr) {
        if (st->n_links < 0 || st->nmax < 2 || st->n_replicas < 1 || st->n_replicas > 65535) return TARL_E_BADARG;
        *row = RowAcc{st->x, st->x_row_stride, st->x_replica_stride, st->n_links, st->nmax, st->cc, 0.0f};
        *is_store = false;
        *R = st->n_replicas;
        return TARL_OK;
    }
    Store s;
    int rc = make_store(st->store, &s);
    if (rc != TARL_OK) return rc;
    if (st->n_nodes < s.N || (st->n_nodes > s.N && st->src_sel == nullptr)) return TARL_E_BADARG;
    *sto = StoreAcc{s, static_cast<float*>(st->store->hot_cur), st->src_sel, st->n_nodes, st->t_garbage, s.N, s.Nmax};
    *is_store = true;
    *R = s.R;
    return TARL_OK;
}

int check_agents(const tarl_agent_table* t, int R, AgentTable* at) {
    if (t == nullptr || t->agent_features == nullptr || t->n_rows < 1) return TARL_E_BADARG;
    if (R > 1 && t->replica_stride < (int64_t)t->n_rows * 9) return TARL_E_BADARG;
    *at = AgentTable{t->agent_features, R > 1 ? t->replica_stride : 0, t->n_rows};
    return TARL_OK;
}

}  // namespace

extern "C" {

int tarl_agents_insert(const tarl_agent_state* state, const tarl_agent_table* agents, const tarl_agent_index* index,
                       float t, int32_t* head, int32_t* next, int32_t* cursor, int32_t* counters, int32_t* inserted,
                       int32_t* flags, int32_t* worklist, int32_t* work_count, float* num_out, int32_t* occupancy,
                       const int32_t* road_origin, void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R; AgentTable at;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if ((rc = check_agents(agents, R, &at)) != TARL_OK) return rc;
    if (index == nullptr || flags == nullptr || index->n_origins < 0) return TARL_E_BADARG;
    const int N = is_store ? sto.N : row.N;
    if (index->n_origins == 0 || N == 0) return TARL_OK;
    if (!index->org_ptr || !index->org_agent || !index->origins || !head || !next || !cursor) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const dim3 g1(blocks_for(index->n_origins), R);
    const dim3 g2(worklist != nullptr ? (g1.x < 4 ? g1.x : 4) : g1.x, R);      // worklist: CTAs stride over the listed origins
    if ((worklist != nullptr) != (work_count != nullptr)) return TARL_E_BADARG;
    if ((num_out != nullptr) != (occupancy != nullptr) || (num_out != nullptr && !is_store)) return TARL_E_BADARG;
    if (worklist != nullptr && cudaMemsetAsync(work_count, 0, sizeof(int32_t) * R, cs) != cudaSuccess) return TARL_E_LAUNCH;
    // work_count = [R] counters followed by n_origins words of scratch: the departed agents per origin at time t,
    // computed once per origin instead of once per (replica, origin)
    int32_t* dep_scratch = nullptr;
    if (work_count != nullptr && inserted != nullptr && index->dep_sorted != nullptr && R > 1) {
        dep_scratch = work_count + R;
        k_insert_departed<<<blocks_for(index->n_origins), kThreads, 0, cs>>>(*index, t, dep_scratch);
    }
    if (road_origin != nullptr && inserted != nullptr && index->dep_sorted != nullptr) {
        if (is_store)
            k_insert_direct<<<g1, kThreads, 0, cs>>>(sto, *index, at, t, road_origin, cursor, counters, inserted, dep_scratch,
                                                     flags, num_out, occupancy, sto.n_nodes);
        else
            k_insert_direct<<<g1, kThreads, 0, cs>>>(row, *index, at, t, road_origin, cursor, counters, inserted, dep_scratch,
                                                     flags, nullptr, nullptr, 0);
        return launch_status();
    }
    if (is_store) {
        k_insert_offer<<<g1, kThreads, 0, cs>>>(sto, *index, at, t, head, next, cursor, flags, inserted, worklist, work_count, dep_scratch);
        k_insert_admit<<<g2, kThreads, 0, cs>>>(sto, *index, at, t, head, next, cursor, counters, inserted, worklist, work_count,
                                                dep_scratch, num_out, occupancy, sto.n_nodes);
    } else {
        k_insert_offer<<<g1, kThreads, 0, cs>>>(row, *index, at, t, head, next, cursor, flags, inserted, worklist, work_count, dep_scratch);
        k_insert_admit<<<g2, kThreads, 0, cs>>>(row, *index, at, t, head, next, cursor, counters, inserted, worklist, work_count,
                                                dep_scratch, nullptr, nullptr, 0);
    }
    return launch_status();
}

int tarl_agents_withdraw(const tarl_agent_state* state, const tarl_agent_table* agents, const tarl_csr* adjacency,
                         float t, uint8_t* mask, int32_t* counters, int32_t* flags, float* num_out, int32_t* occupancy,
                         void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R; AgentTable at;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if ((rc = check_agents(agents, R, &at)) != TARL_OK) return rc;
    const int N = is_store ? sto.N : row.N;
    if (N == 0) return TARL_OK;
    if (adjacency == nullptr || flags == nullptr || adjacency->n_rows < 0) return TARL_E_BADARG;
    if (adjacency->n_rows > 0 && (!adjacency->ptr || (adjacency->n_edges > 0 && !adjacency->idx))) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const dim3 grid(blocks_for(N), R);
    if ((num_out != nullptr) != (occupancy != nullptr) || (num_out != nullptr && !is_store)) return TARL_E_BADARG;
    if (num_out != nullptr) {
        if (cudaMemsetAsync(occupancy, 0, sizeof(int32_t) * (size_t)R, cs) != cudaSuccess) return TARL_E_LAUNCH;
        const dim3 grid_nodes(blocks_for(sto.n_nodes), R);
        k_withdraw_observe<<<grid_nodes, kThreads, 0, cs>>>(sto, at, *adjacency, t, mask, counters, flags, num_out, occupancy);
        return launch_status();
    }
    if (is_store) k_withdraw<<<grid, kThreads, 0, cs>>>(sto, at, *adjacency, t, mask, counters, flags);
    else k_withdraw<<<grid, kThreads, 0, cs>>>(row, at, *adjacency, t, mask, counters, flags);
    return launch_status();
}

int tarl_agents_choice(const tarl_agent_state* state, const tarl_csr* neighbours, const int32_t* choosers,
                       int32_t n_choosers, const float* uniforms, uint64_t seed, uint32_t step_id, void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if (n_choosers < 0 || neighbours == nullptr) return TARL_E_BADARG;
    if (n_choosers == 0) return TARL_OK;
    if (!choosers || !neighbours->ptr || !neighbours->idx) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const dim3 grid(blocks_for(n_choosers), R);
    if (is_store)
        k_choice<<<grid, kThreads, 0, cs>>>(sto, *neighbours, choosers, n_choosers, uniforms, (uint32_t)seed,
                                            (uint32_t)(seed >> 32), step_id);
    else
        k_choice<<<grid, kThreads, 0, cs>>>(row, *neighbours, choosers, n_choosers, uniforms, (uint32_t)seed,
                                            (uint32_t)(seed >> 32), step_id);
    return launch_status();
}

int tarl_agents_apply_action(const tarl_agent_state* state, const int32_t* edge_src, const int32_t* edge_dst,
                             int32_t n_edges, const tarl_rows* action, int32_t action_dtype, void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if (n_edges < 0 || action_dtype < 0 || action_dtype > 2) return TARL_E_BADARG;
    if (n_edges == 0) return TARL_OK;
    if (!edge_src || !edge_dst || !action || !action->data) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const int grid = blocks_for((int64_t)n_edges * R);
    const bool rin = R > 1 && action->row_stride == 1;
    if (is_store)
        k_apply_action<<<grid, kThreads, 0, cs>>>(sto, edge_src, edge_dst, n_edges, R, action->data, action->row_stride,
                                                  action->col_stride, action_dtype, rin);
    else
        k_apply_action<<<grid, kThreads, 0, cs>>>(row, edge_src, edge_dst, n_edges, R, action->data, action->row_stride,
                                                  action->col_stride, action_dtype, rin);
    return launch_status();
}

int tarl_agents_apply_action_groups(const tarl_agent_state* state, const tarl_csr* groups, const int32_t* group_node,
                                    const int32_t* edge_dst, const tarl_rows* action, int32_t action_dtype, void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if (groups == nullptr || groups->n_rows < 0 || groups->n_edges < 0 || action_dtype < 0 || action_dtype > 2)
        return TARL_E_BADARG;
    if (groups->n_rows == 0 || groups->n_edges == 0) return TARL_OK;
    if (!groups->ptr || !groups->eid || !group_node || !edge_dst || !action || !action->data) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const dim3 grid = tarl::tile_grid(groups->n_rows, R);
    const int Bp = tarl::tile_rows_pow2(R);
    if (is_store)
        k_apply_action_groups<<<grid, tarl::kTileThreads, 0, cs>>>(sto, *groups, group_node, edge_dst, R, Bp, action->data,
                                                                   action->row_stride, action->col_stride, action_dtype);
    else
        k_apply_action_groups<<<grid, tarl::kTileThreads, 0, cs>>>(row, *groups, group_node, edge_dst, R, Bp, action->data,
                                                                   action->row_stride, action->col_stride, action_dtype);
    return launch_status();
}

int tarl_store_observe(const tarl_agent_state* state, float* node_features, int64_t* agent_index, int32_t* occupancy,
                       float* num_agents, float* selected_road, void* stream) {
    RowAcc row; StoreAcc sto; bool is_store; int R;
    int rc = check_state(state, &row, &sto, &is_store, &R);
    if (rc != TARL_OK) return rc;
    if (!is_store) return TARL_E_BADARG;
    if (sto.n_nodes == 0) return TARL_OK;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (occupancy != nullptr && cudaMemsetAsync(occupancy, 0, sizeof(int32_t) * (size_t)R, cs) != cudaSuccess)
        return TARL_E_LAUNCH;
    const dim3 grid(blocks_for(sto.n_nodes), R);
    k_observe<<<grid, kThreads, 0, cs>>>(sto, node_features, reinterpret_cast<long long*>(agent_index), occupancy,
                                         num_agents, selected_road);
    return launch_status();
}

}  // extern "C"

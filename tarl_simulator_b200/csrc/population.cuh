// population.cuh — what both the population kernels (agents.cu) and the fused response + withdrawal kernel of the link
// store (engine.cu) need: the agent table view, the link-store accessor (logical FIFO slots over the hot records and
// the ring queues) and the withdrawal of one link (src/agents/base.py:334-403).
#pragma once
#include "engine_common.cuh"

namespace tarl {

constexpr int kDestination = 1, kDepartureTime = 2, kArrivalTime = 3, kOnWay = 7, kDone = 8;  // AgentFeatureHelpers

struct AgentTable {       // agent_features [R?, A+1, 9] fp32
    float* af;
    int64_t rep_stride;   // 0 = one table shared by... (only legal when R == 1)
    int32_t n_rows;
    __device__ float* row(int r, long long a) const { return af + r * rep_stride + a * 9; }
};

struct StoreAcc {
    Store s;
    float* hot;           // == s.hot_cur, writable: these operations run between two steps, on the current records
    float* src_sel;       // [R, n_nodes - N] SELECTED_ROAD of the non-road nodes (SRC nodes use it)
    int n_nodes;
    float t_garbage;      // arrival time of a pending tail-garbage record = time of the latest core step
    int N, Nmax;
    // links are addressed by their id; the store may hold them in its own slot order (Store::slot_of)
    __device__ float sel_of(int r, int node) const {
        return node < N ? s.sel[(size_t)r * N + s.slot_of(node)] : src_sel[(size_t)r * (n_nodes - N) + (node - N)];
    }
    __device__ void set_sel(int r, int node, float v) const {
        if (node < N) s.sel[(size_t)r * N + s.slot_of(node)] = v;
        else src_sel[(size_t)r * (n_nodes - N) + (node - N)] = v;
    }
    __device__ float road_index(int, int n) const { return s.stat_a[s.slot_of(n)].z; }

    struct Link {
        float4* rec;      // the two halves of the hot record
        float4* ring;
        float4 A, B;
        int M, rh;
        bool gv;
        float num, maxn, fftt, t_garbage;
    };
    __device__ Link open(int r, int n) const {
        const int slot = s.slot_of(n);
        const size_t L = (size_t)r * N + slot;
        float4* rec = reinterpret_cast<float4*>(hot) + 2 * L;
        Link l;
        l.rec = rec; l.ring = s.queue + L * s.M; l.A = rec[0]; l.B = rec[1]; l.M = s.M;
        const int meta = __float_as_int(l.B.w);
        l.rh = meta & kMetaRingMask; l.gv = (meta & kMetaGarbage) != 0;
        l.num = l.A.z; l.maxn = l.A.w; l.fftt = s.stat_a[slot].x; l.t_garbage = t_garbage;
        return l;
    }
    __device__ static float4 slot(const Link& l, int k) {      // logical FIFO slot k as {id, arrival, exit}
        if (k == 0) return make_float4(l.A.x, l.B.x, l.A.y, 0.0f);
        if (l.gv && k == (int)l.num) return make_float4(0.0f, l.t_garbage, l.B.z, 0.0f);
        return l.ring[ring_pos(l.rh, k, l.M)];
    }
    __device__ static float slot_id(const Link& l, int k) { return slot(l, k).x; }
    __device__ static float slot_dep(const Link& l, int k) { return slot(l, k).z; }
    __device__ static void put(Link& l, int k, float id, float arr, float dep) {
        if (k == 0) { l.A.x = id; l.B.x = arr; l.A.y = dep; }
        else l.ring[ring_pos(l.rh, k, l.M)] = make_float4(id, arr, dep, 0.0f);
        l.B.y = id;         // the latest append is the tail
        l.gv = false;       // appends start at slot int(NUM): the pending garbage record is overwritten
    }
    __device__ static void store_back(Link& l) {
        l.B.w = __int_as_float(l.rh | (l.gv ? kMetaGarbage : 0));
        l.rec[0] = l.A; l.rec[1] = l.B;
    }
    __device__ static void commit_insert(Link& l, int admitted) {
        l.A.z = l.num + (float)admitted;
        store_back(l);
    }
    __device__ static void withdraw(Link& l, int c) {
        const int g = (int)l.num;                     // slot of the pending garbage record, if any
        const float4 head = slot(l, c);               // c <= NUM < Nmax, so slot c exists
        for (int k = 1; k <= c; ++k) l.ring[ring_pos(l.rh, k, l.M)] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero fill
        int rh = l.rh + c; if (rh >= l.M) rh -= l.M;
        l.rh = rh;
        l.A.x = head.x; l.B.x = head.y; l.A.y = head.z;
        if (l.gv && g == c) l.gv = false;             // the garbage record became the head slot
        l.A.z = l.num - (float)c;
        store_back(l);
    }
};

// ---------------------------------------------------------------------------------------------------------- withdraw
// The maximal PREFIX of a link's queue whose agents are due (exit time <= t) and whose DESTINATION node is adjacent to
// this link leaves the network (base.py:355-400). adj = CSR of the FULL edge_index by source node: the sparse form of
// the reference's dense adj_matrix[ROAD_INDEX, DESTINATION] lookup.
__device__ __forceinline__ bool adjacent(const tarl_csr& adj, long long row, long long dest) {
    if (row < 0 || row >= adj.n_rows) return false;
    const int k1 = adj.ptr[row + 1];
    for (int k = adj.ptr[row]; k < k1; ++k)
        if (adj.idx[k] == dest) return true;
    return false;
}

// one link of one replica; returns its NUM after the withdrawal
template <class Acc>
__device__ __forceinline__ float withdraw_one(const Acc& acc, const AgentTable& at, const tarl_csr& adj, float t,
                                              uint8_t* __restrict__ mask, int32_t* __restrict__ counters,
                                              int32_t* __restrict__ flags, int r, int n) {
    typename Acc::Link l = acc.open(r, n);
    int c = 0;
    long long ridx = -1;
    while (c < acc.Nmax && (float)c < l.num) {                                   // active_slots, :363-366
        if (!(Acc::slot_dep(l, c) <= t)) break;                                  // depart_ok, :362
        const long long a = (long long)Acc::slot_id(l, c);
        if (a < 0 || a >= at.n_rows) { atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_AGENT_RANGE); break; }
        if (c == 0) ridx = (long long)acc.road_index(r, n);
        if (!adjacent(adj, ridx, (long long)at.row(r, a)[kDestination])) break;  // connectivity, :361
        ++c;
    }
    if (mask != nullptr) mask[(size_t)r * acc.N + n] = c > 0 ? 1 : 0;
    if (c == 0) return l.num;
    for (int k = 0; k < c; ++k) {                                                // :398-400
        float* p = at.row(r, (long long)Acc::slot_id(l, k));
        p[kDone] = 1.0f; p[kOnWay] = 0.0f; p[kArrivalTime] = t;
    }
    Acc::withdraw(l, c);
    if (counters != nullptr) atomicAdd(&counters[2 * r + 1], c);
    return l.num - (float)c;
}

}  // namespace tarl

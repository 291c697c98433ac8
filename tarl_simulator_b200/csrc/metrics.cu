// metrics.cu — on-device side channels of the simulation loop (sm_100a): hourly traffic counters per link and the
// per-link road-optimality aggregate, accumulated where the step's masks are produced instead of keeping one bool[N]
// (and one fp32[E] host copy) per timestep as the reference does.
//
// Reference semantics:
//   * ResponseMPNN.update_history / Agents.withdraw_history hold (time, bool[N]) per step
//     (/root/reference/src/response_mpnn.py:125, /root/reference/src/agents/base.py:402);
//     TransportationSimulator.compute_node_metrics sums them per hour = time // 3600
//     (/root/reference/src/transportation_simulator.py:584-610): counts[n, h] = number of recorded steps of hour h on
//     which link n popped its head (hand-off) or had agents withdrawn.
//   * road_optimality_values holds delta_travel_time[E] per step (:351); plot_road_optimality aggregates it by the
//     edge's SOURCE link with scatter_add over edge_index_routes[0] (:486-488): agg[n] = sum over out-edges of n.
// Integer counters: one thread owns one (replica, link) cell, so there are no atomics and the result is exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_metrics_accumulate(
    const uint8_t* __restrict__ pop, const uint8_t* __restrict__ withdrawn, int N, int R, const float* __restrict__ delta_tt,
    int E, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_eid, int hour, int H,
    int32_t* __restrict__ counts, float* __restrict__ optimality_sum, float* __restrict__ optimality_now) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (int64_t)N * R) return;
    const int r = (int)(i / N), n = (int)(i - (int64_t)r * N);
    const int64_t cell = ((int64_t)r * H + hour) * N + n;
    if (counts != nullptr) {
        int c = 0;
        if (pop != nullptr) c += pop[i] != 0;
        if (withdrawn != nullptr) c += withdrawn[i] != 0;
        if (c != 0) counts[cell] += c;
    }
    if (delta_tt != nullptr) {
        // ascending original edge id inside the segment: the order scatter_add visits this link's out-edges in
        const float* d = delta_tt + (int64_t)r * E;
        float acc = 0.0f;
        const int k1 = out_ptr[n + 1];
        for (int k = out_ptr[n]; k < k1; ++k) acc += d[out_eid != nullptr ? out_eid[k] : k];
        if (optimality_now != nullptr) optimality_now[i] = acc;
        if (optimality_sum != nullptr) optimality_sum[cell] += acc;
    }
}

}  // namespace

extern "C" int tarl_metrics_accumulate(const tarl_dual_csr* g, int32_t n_replicas, const uint8_t* pop,
                                       const uint8_t* withdrawn, const float* delta_tt, int32_t hour, int32_t n_hours,
                                       int32_t* counts, float* optimality_sum, float* optimality_now, void* stream) {
    if (g == nullptr || g->n_links < 0 || n_replicas < 1 || hour < 0 || hour >= n_hours) return TARL_E_BADARG;
    if (g->n_links == 0) return TARL_OK;
    if (delta_tt != nullptr && g->out_ptr == nullptr) return TARL_E_BADARG;
    if (counts == nullptr && (delta_tt == nullptr || (optimality_sum == nullptr && optimality_now == nullptr))) return TARL_OK;
    const int64_t cells = (int64_t)g->n_links * n_replicas;
    k_metrics_accumulate<<<(unsigned)((cells + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        pop, withdrawn, g->n_links, n_replicas, delta_tt, g->n_edges, g->out_ptr, g->out_eid, hour, n_hours, counts,
        optimality_sum, optimality_now);
    return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
}

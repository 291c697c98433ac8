/*
 * tarl_b200.h — C ABI of libtarl_b200.so, the sm_100a (B200) implementation of TARL-simulator's data-parallel hot
 * path: the per-timestep network step and the learned-MPNN forward/backward.
 *
 * The reference (OliBus801/TARL-simulator) is pure Python/PyTorch and has NO native/FFI interface of its own; the
 * boundary a maintainer binds is therefore defined here and cited, entry by entry, against the reference Python
 * function it replaces (paths relative to the reference repo root). INTEGRATION.md shows the ctypes stub.
 *
 * Conventions (every entry point):
 *   - plain C: device pointers + sizes only, no torch types; the CALLER owns every buffer (no allocation inside);
 *   - asynchronous on `stream` (a cudaStream_t passed as void*; 0 = legacy default stream);
 *   - returns 0 on success or a negative TARL_E_* code for argument/launch errors (see tarl_error_string);
 *   - data-dependent faults (queue overflow, Gumbel arg-max without a winner) are reported through a sticky device
 *     word `flags[TARL_FLAG_ERROR]` that the host may read whenever it chooses to synchronise;
 *   - re-entrant, no global state; fp32 state, int32 topology.
 */
#ifndef TARL_B200_H
#define TARL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TARL_ABI_VERSION 1

/* return codes */
#define TARL_OK 0
#define TARL_E_BADARG (-1)
#define TARL_E_WORKSPACE (-2)
#define TARL_E_LAUNCH (-3)

/* flags[] words written by the kernels (int32 device array of TARL_FLAG_COUNT words, zeroed by the caller) */
#define TARL_FLAG_ANY_POP 0 /* != 0 iff at least one link popped its FIFO head in the last response phase  */
#define TARL_FLAG_ERROR 1   /* sticky OR of TARL_ERR_* bits                                              */
#define TARL_FLAG_COUNT 4

#define TARL_ERR_QUEUE_RANGE 1 /* NUMBER_OF_AGENT of some link is <0, NaN or >= Nmax: the reference's tail write
                                  (src/direction_mpnn.py:175-191) would alias other columns / raise IndexError */
#define TARL_ERR_NO_WINNER 2   /* a link had positive total probability but no finite Gumbel score (u == 0 or NaN):
                                  the reference raises IndexError at src/direction_mpnn.py:144 */

/* Static topology of the dual graph (`edge_index_routes` of the reference, src/transportation_simulator.py:150-171)
 * in both CSR orientations. Original edge ids are kept because the Gumbel arg-max breaks ties towards the lowest
 * edge id (torch-scatter CPU semantics) and delta_travel_time is reported in original edge order. Within every CSR
 * segment edges MUST be in ascending original id (stable sort). All pointers are device pointers. */
typedef struct tarl_dual_csr {
    int32_t n_links;        /* N = graph.num_roads                                  */
    int32_t n_edges;        /* E = edge_index_routes.size(1)                        */
    const int32_t* in_ptr;  /* [N+1] in-edges of downstream link d: [in_ptr[d], in_ptr[d+1])  */
    const int32_t* in_src;  /* [E]   upstream link of the k-th in-edge               */
    const int32_t* in_eid;  /* [E]   its original edge id                            */
    const int32_t* out_ptr; /* [N+1] out-edges of upstream link u                    */
    const int32_t* out_dst; /* [E]   downstream link of the k-th out-edge            */
    const int32_t* out_eid; /* [E]   its original edge id                            */
} tarl_dual_csr;

int tarl_abi_version(void);
const char* tarl_error_string(int code);

/* ---------------------------------------------------------------------------------------------------------------
 * Per-timestep network step on the reference's own row layout, IN PLACE.
 *
 * x: road rows of `graph.x` (src/feature_helpers.py:38-54): fp32 [N, 3*nmax+7], element (n, c) at
 *    x[n*x_row_stride + c]; columns [0,nmax) FIFO agent ids (head = col 0), [nmax,2nmax) arrival times,
 *    [2nmax,3nmax) scheduled exit times, then MAXN, NUM, FFTT, LENGTH, MAX_FLOW, SELECTED_ROAD, ROAD_INDEX.
 * edge_attr: edge_attr_routes, [E] in original edge order.   cc: congestion_constant[:N] or NULL (then the formula
 *    of src/simulation_core_model.py:58-67 is evaluated in-kernel).   noise: the E uniforms the reference draws with
 *    torch.rand_like at src/direction_mpnn.py:137, original edge order.   sel: NULL, or [N] SELECTED_ROAD values for
 *    this step, written into x[:, SELECTED_ROAD] before anything reads it — the fused form of the caller's
 *    "apply action / choice, then core" sequence (src/reinforcement_learning.py:231-237,
 *    src/transportation_simulator.py:316-322).   t: the simulation time baked in by set_time
 *    (src/simulation_core_model.py:85-88), as fp32.
 * workspace: tarl_core_workspace_bytes(N) bytes of device scratch, 16-byte aligned.
 * ------------------------------------------------------------------------------------------------------------- */
size_t tarl_core_workspace_bytes(int32_t n_links);

/* Replaces DirectionMPNN.forward = message + aggregate + update (src/direction_mpnn.py:44-196, 199-236):
 * eligibility masks, per-downstream-link Gumbel-max pick of one upstream head, tail append on EVERY link.
 * delta_tt: [E] out, road_optimality_data["delta_travel_time"] in original edge order (may be NULL). */
int tarl_direction_forward(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax,
                           const float* edge_attr, const float* cc, const float* noise, const float* sel, float t,
                           float* delta_tt,
                           int32_t* flags, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces ResponseMPNN.forward = message + max-aggregate + update (src/response_mpnn.py:27-127): an upstream link
 * pops its FIFO head iff the tail of one of its downstream links now equals that head; 3 queue segments shift left.
 * pop: [N] out (uint8 0/1), the mask the reference appends to update_history; flags[TARL_FLAG_ANY_POP] tells whether
 * the reference would have appended at all (src/response_mpnn.py:106-107,125). */
int tarl_response_forward(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax, uint8_t* pop,
                          int32_t* flags, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces SimulationCoreModel.forward (src/simulation_core_model.py:41-83): direction then response, sharing the
 * per-link summaries so that x is read once, not gathered four times per edge. */
int tarl_core_step(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax, const float* edge_attr,
                   const float* cc, const float* noise, const float* sel, float t, float* delta_tt, uint8_t* pop,
                   int32_t* flags, void* workspace, size_t workspace_bytes, void* stream);

/* The same step with its three kernels individually selectable (profiling and per-kernel timing only; a partial
 * mask leaves x mid-step). Phases must be issued in order on one stream. */
#define TARL_PHASE_OFFER 1u         /* per-link summaries of the pre-step rows                      */
#define TARL_PHASE_SELECT_APPEND 2u /* masks + Gumbel arg-max per downstream link, tail append      */
#define TARL_PHASE_RESPOND_SHIFT 4u /* acknowledgement, delta_tt, FIFO shift of popping links       */
#define TARL_PHASE_ALL 7u
int tarl_core_step_phases(const tarl_dual_csr* g, float* x, int64_t x_row_stride, int32_t nmax,
                          const float* edge_attr, const float* cc, const float* noise, const float* sel, float t,
                          float* delta_tt,
                          uint8_t* pop, int32_t* flags, void* workspace, size_t workspace_bytes, void* stream,
                          uint32_t phase_mask);

#ifdef __cplusplus
}
#endif
#endif /* TARL_B200_H */

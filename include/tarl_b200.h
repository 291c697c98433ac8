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

#define TARL_ABI_VERSION 28

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
#define TARL_ERR_EMBED_RANGE 4 /* an embedding index fell outside nodes_embedding (nn.Embedding raises IndexError,
                                  src/agents/mpnn_agent.py:216) */
#define TARL_ERR_INSERT_TARGET 8 /* SELECTED_ROAD of an origin node with agents is not a road id: the reference would
                                    index a non-road row at src/agents/base.py:259-266 */
#define TARL_ERR_AGENT_RANGE 16  /* a queued agent id is outside agent_features (IndexError at src/agents/base.py:358) */

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
    const int32_t* out_eid; /* [E]   its original edge id; NULL = the edge list is already sorted by source,
                                      i.e. out_eid[k] == k (config_network's order) and need not be read    */
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

/* ---------------------------------------------------------------------------------------------------------------
 * Resident link store: the same step on a compact device layout (csrc/engine.cu), for loops that keep the state on
 * the device between steps (rollouts, benchmarks). tarl_store_import / tarl_store_export convert from / to the
 * reference's row layout exactly (every cell of x, including what the reference leaves past the queue tails).
 * All buffers are caller-owned device memory; R replicas of the same network are stepped by one launch.
 * ------------------------------------------------------------------------------------------------------------- */
#define TARL_STORE_UNIFORM_WEIGHTS 1
typedef struct tarl_link_store {
    int32_t n_links;    /* N                                                                             */
    int32_t n_replicas; /* R >= 1: independent copies of the network state (PPO rollout environments)    */
    int32_t nmax;       /* Nmax (FIFO slots per link)                                                    */
    int32_t hints;      /* 0, or TARL_STORE_UNIFORM_WEIGHTS: the caller filled stat_a[.].w (see there)    */
    void* hot_cur;      /* [R*N] 32-byte records holding the CURRENT state: {head id, head exit, NUM, MAXN |
                           head arrival, tail id, pending tail-garbage exit time, meta}                   */
    void* hot_next;     /* [R*N] 32-byte records written by the step; the caller swaps the two afterwards */
    void* sel;          /* [R*N] fp32 SELECTED_ROAD                                                      */
    void* stat_a;       /* [N] 16-byte {FFTT, congestion_constant, ROAD_INDEX, w}. w belongs to the caller:
                           tarl_store_import leaves NaN there. With hints = TARL_STORE_UNIFORM_WEIGHTS it holds, per link
                           (store slot), the edge_attr value shared by ALL in-edges of the link in the topology the
                           step calls are given, or NaN where they differ: the ELL direction kernel then takes the
                           weight from the 16 bytes it loads anyway and reads its edge-weight columns (16 of ~100
                           bytes per link) only for the NaN links                                         */
    void* stat_b;       /* [N] 16-byte {LENGTH, MAX_FLOW, 0, 0} (export only)                            */
    void* queue;        /* [R*N*(nmax-1)] 16-byte ring slots {agent id, arrival, exit, pad}              */
    void* post;         /* [R*N] 8-byte scratch {NUM, tail id} handed from the direction to the response phase */
    void* pop_hint;     /* unused since ABI 23 (may be NULL): the "pop hint" bytes of earlier versions cost the direction
                           phase more than they returned to the response phase                             */
    const int32_t* slot_link; /* [N] link id held by store slot s, or NULL = identity. The store may keep the links
                                 in a locality order of its own (tarl_cluster_links): everything indexed [R*N] above
                                 and the topology handed to tarl_store_step are then in SLOT order; import / export
                                 and the population entry points translate through these two arrays.             */
    const int32_t* link_slot; /* [N] inverse of slot_link, or NULL                                               */
} tarl_link_store;

/* x -> store. x element (r, n, c) at x[r*x_replica_stride + n*x_row_stride + c]; cc = congestion_constant[:N] or
 * NULL (formula). Fills hot_cur, sel, queue and (from replica 0) stat_a / stat_b. */
int tarl_store_import(const tarl_link_store* store, const float* x, int64_t x_row_stride, int64_t x_replica_stride,
                      const float* cc, int32_t* flags, void* stream);

/* store (hot_cur) -> x, bit-exact with what the reference's x would hold. t_last_step: the time of the most recent
 * tarl_store_step (the arrival time the reference wrote past the tails on that step). */
int tarl_store_export(const tarl_link_store* store, float* x, int64_t x_row_stride, int64_t x_replica_stride,
                      float t_last_step, void* stream);

/* Optional ELLPACK copy of the first `width` edges of every link, column-major: entry j of link n at [j*pitch + n].
 * -1 = no such edge. A link with MORE than `width` in-edges (out-edges) carries -2 in column width-1 of in_src
 * (out_dst) and is served from its CSR segment instead; so does — in in_src — a link with an in-edge whose
 * edge_attr_routes is not >= 1e-3 (below that bound an INELIGIBLE edge can win the Gumbel arg-max of
 * src/direction_mpnn.py:136-139, so the link needs the literal scan over every in-edge; the streaming kernel never
 * decides such a link). Edge order inside a link is the CSR's (ascending edge id). */
typedef struct tarl_dual_ell {
    int32_t width;          /* 4 or 8                                                        */
    int32_t pitch;          /* elements between consecutive columns, >= n_links               */
    const int32_t* in_src;  /* [width*pitch] upstream link of the j-th in-edge                */
    const float* in_attr;   /* [width*pitch] edge_attr_routes of that edge                    */
    const int32_t* out_dst; /* [width*pitch] downstream link of the j-th out-edge             */
} tarl_dual_ell;

/* Inputs and outputs of one store step (all device pointers, caller-owned).
 * noise: [R*E] uniforms in original edge order per replica, or NULL to draw them in-kernel: Philox4x32-10 keyed by
 *   `seed`, counter = (replica*N + link, 0, step_id, in-edge rank / 4), uniform number (in-edge rank % 4) of the block,
 *   in-edge rank = position of the edge among the link's in-edges in ascending edge id — a documented stream of its
 *   own, not torch's (the reference draws torch.rand_like at src/direction_mpnn.py:137; declared divergence D4).
 *   tarl_store_noise writes that stream out in the [R*E] form, so that the CPU oracle can replay a step exactly.
 * delta_tt_link: NULL or [R*N]: delta_travel_time (src/direction_mpnn.py:94-96) is a function of the UPSTREAM link of an
 *   edge only, so the step emits one value per link; tarl_expand_delta_tt turns it into the reference's [E] vector.
 * pop: [R*N] bytes (the mask of update_history). pop_bits: NULL or [R * ceil(N/32)] words, bit (n % 32) of word
 *   r*ceil(N/32) + n/32 = pop[r, n] — the form that travels to the host. flags: TARL_FLAG_* words. */
typedef struct tarl_step_io {
    const float* noise;
    uint64_t seed;
    const uint64_t* seed_dev; /* NULL, or a device word holding the key instead of `seed`: a step captured in a CUDA graph
                                 then draws new noise on every replay once the caller has changed the word */
    uint32_t step_id;
    float t;
    float* delta_tt_link;
    uint8_t* pop;
    uint32_t* pop_bits;
    int32_t* flags;
} tarl_step_io;

/* SimulationCoreModel.forward on the store (src/simulation_core_model.py:41-83): reads hot_cur, writes hot_next.
 * ell: NULL (every link walks its CSR segment) or the ELLPACK copy above (same results, shorter dependent-load chain).
 * attr_in: edge_attr_routes permuted into g->in_* order. phase_mask: TARL_PHASE_SELECT_APPEND |
 * TARL_PHASE_RESPOND_SHIFT (both for a full step). Two launches per step. */
int tarl_store_step(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                    const float* attr_in, const tarl_step_io* io, void* stream, uint32_t phase_mask);

/* The uniforms tarl_store_step(noise = NULL) uses for (seed, step_id), written out as [R*E] in original edge order:
 * a step replayed with noise = that array is bit-identical to the step with the in-kernel stream, and the CPU oracle
 * can be fed the same numbers (tests/test_store_replay_gpu.py). */
int tarl_store_noise(const tarl_dual_csr* g, int32_t n_replicas, uint64_t seed, uint32_t step_id, float* noise,
                     void* stream);

/* road_optimality_data["delta_travel_time"] (src/direction_mpnn.py:94-99): out[r, e] = delta_tt_link[r, source link of
 * e], original edge order. edge_src: [E] int32 = edge_index_routes[0]. */
int tarl_expand_delta_tt(const int32_t* edge_src, int32_t n_edges, int32_t n_links, int32_t n_replicas,
                         const float* delta_tt_link, float* delta_tt, void* stream);

/* HOST-side helper (plain C++, host pointers): a locality order of the links for the store — clusters of `cluster`
 * links grown breadth-first over the dual graph (adj = CSR listing each link's in- and out-neighbours), emitted one
 * after the other: order[slot] = link id. A CTA tile of `cluster` consecutive slots then gathers mostly from itself. */
int tarl_cluster_links(int32_t n_links, const int32_t* adj_ptr, const int32_t* adj_idx, int32_t cluster,
                       int32_t* order);

/* n_steps consecutive steps enqueued by one call (times t0, t0+dt, ...; in-kernel noise; step ids first_step_id, +1,
 * ...): the loop SimulatorEnv.rollout / TransportationSimulator.run drive from Python, without a host round trip per
 * step. io: first step's time (io->t) and step id, in-kernel noise only (io->noise must be NULL). sel_bank: NULL/0
 * (SELECTED_ROAD stays as it is) or n_bank device arrays [R*N] cycled through as the routing
 * decisions of successive steps. hot_cur / hot_next alternate internally: after an odd n_steps the caller's two
 * buffers have swapped roles. delta_tt_link / pop / pop_bits hold the LAST step's outputs. */
int tarl_store_run(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                   const float* attr_in, const tarl_step_io* io, float dt, int32_t n_steps,
                   const float* const* sel_bank, int32_t n_bank, void* stream);

/* One step whose inputs and outputs live in HOST memory (what SimulationCoreModel.forward(graph, selected_road=<CPU
 * tensor>, host_out=...) runs): the routing decisions come from sel_host (pinned, [R*N] fp32; NULL = keep the store's),
 * delta_travel_time per upstream link and the pop bits go to dtt_host / pop_bits_host (pinned; NULL = not copied) —
 * the copies of src/transportation_simulator.py:351 (.cpu() of the per-step outputs) and of the routing decisions an
 * external controller hands in. The copies run on two streams of the pipe's own, ordered against the kernels by
 * events, so that with two slots used alternately (slot = step & 1) the upload of step k+1 and the download of step
 * k-1 overlap the kernels of step k while the host merely enqueues: ONE library call per step.
 * Per slot the caller owns: sel_stage (device [R*N] fp32, receives sel_host and becomes this step's SELECTED_ROAD —
 * pass it as store->sel), io->delta_tt_link and io->pop_bits (device outputs of this step; the next use of the slot
 * waits for their download). tarl_host_pipe_join makes `stream` wait for every copy issued so far (timing, or
 * before the host reads the buffers after synchronising `stream`). */
typedef struct tarl_host_pipe tarl_host_pipe;
int tarl_host_pipe_create(tarl_host_pipe** pipe);
int tarl_host_pipe_destroy(tarl_host_pipe* pipe);
int tarl_host_pipe_join(tarl_host_pipe* pipe, void* stream);
int tarl_store_step_host(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                         const float* attr_in, const tarl_step_io* io, tarl_host_pipe* pipe, int32_t slot,
                         const float* sel_host, float* sel_stage, float* dtt_host, uint32_t* pop_bits_host,
                         void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Learned-MPNN path (fp32, tolerance 1e-5 relative against the reference).
 * ------------------------------------------------------------------------------------------------------------- */

/* One CSR orientation of an edge list: row r owns [ptr[r], ptr[r+1]); eid[k] = original edge id of the k-th entry,
 * ascending inside every row; idx[k] = the other endpoint (may be NULL where unused). */
typedef struct tarl_csr {
    int32_t n_rows;
    int32_t n_edges;
    const int32_t* ptr;
    const int32_t* idx;
    const int32_t* eid;
} tarl_csr;

/* A [B, E]-shaped tensor with arbitrary element strides: element (b, e) at data[b*row_stride + e*col_stride].
 * The policy kernels emit EDGE-major tensors (row_stride 1, col_stride B): the B rows of one edge are contiguous. */
typedef struct tarl_rows {
    void* data;
    int64_t row_stride;
    int64_t col_stride;
} tarl_rows;

/* Replaces MPNNPolicyNet.forward's active path (src/agents/mpnn_agent.py:117-192; update_edges :215-217):
 * logits[b,e] = nodes_embedding[ idx(b, edge_index[1][e]) ],  idx(b,n) = ROAD_INDEX(b,n) if >= 0 else n  (declared
 * divergence D2: the literal code raises on ROAD_INDEX = -1 rows). node_features: [B,N,*] fp32 with the given
 * element strides. Outputs (caller-owned), batch row innermost: node_emb and node_idx (what backward needs) hold
 * element (b, n) at n*B + b; logits holds element (b, e) at e*B + b. */
int tarl_policy_embed_forward(const float* emb_weight, int32_t emb_rows, const float* node_features,
                              int64_t nf_batch_stride, int64_t nf_row_stride, int32_t road_index_col, int32_t batch,
                              int32_t n_nodes, const int32_t* edge_dst, int32_t n_edges, float* node_emb,
                              int32_t* node_idx, float* logits, int32_t* flags, void* stream);

/* Gradient of the above w.r.t. nodes_embedding.weight (what embedding_dense_backward computes in the reference):
 * by_target = CSR of the full edge_index by target node. grad_logits: [B,E] with any strides. node_grad: [N*B]
 * scratch. grad_weight [emb_rows] is overwritten. Fixed summation order (no float atomics when idx is injective and
 * batch-invariant). */
int tarl_policy_embed_backward(const tarl_csr* by_target, const tarl_rows* grad_logits, const int32_t* node_idx,
                               int32_t batch, float* node_grad, float* grad_weight, int32_t emb_rows, void* stream);

/* GraphDistribution (src/reinforcement_learning.py:15-96): categorical over the out-edges of every source node.
 * groups = CSR of edge_index by RANK of the source id (declared divergence D1). logits [B,E], any strides.
 * Any of proba [B,E], mode [B,E] (one-hot of the per-group arg-max, lowest edge id on ties; the CALLER zeroes it),
 * entropy [B], log_prob [B] may be NULL. action: [B,E] one-hot in the dtype named by action_dtype (required for
 * log_prob; a row without exactly one selected edge per group gets -inf).
 * partials: 3*B*tarl_graphdist_partial_count(K, B) floats. */
#define TARL_ACTION_U8 0
#define TARL_ACTION_I64 1
#define TARL_ACTION_F32 2
int32_t tarl_graphdist_partial_count(int32_t n_groups, int32_t batch);
int tarl_graphdist_forward(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                           const tarl_rows* action, int32_t action_dtype, const tarl_rows* proba, const tarl_rows* mode,
                           float* entropy, float* log_prob, float* partials, void* stream);

/* d(sum_b grad_log_prob[b]*log_prob[b] + grad_entropy[b]*entropy[b]) / d logits  -> grad_logits [B,E]. Either
 * upstream gradient may be NULL. log_prob (forward output, may be NULL) marks -inf rows, which get no gradient. */
int tarl_graphdist_backward(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                            const tarl_rows* action, int32_t action_dtype, const float* grad_log_prob,
                            const float* grad_entropy, const float* log_prob, const tarl_rows* grad_logits, void* stream);

/* GraphDistribution.sample (:57-80): uniforms [B,K] (any strides; group-major memory reads fastest), one per (row,
 * group) in ascending source id; onehot [B,E] out in
 * int64 (TARL_ACTION_I64, the reference's dtype) or uint8 / bool (TARL_ACTION_U8), every entry written. Inside a
 * group edges are walked in ascending edge id (D3); batched rows are independent (D7).
 * log_prob: NULL, or [B] out = log_prob of the sampled action, accumulated in the same pass (needs `partials` as for
 * tarl_graphdist_forward). Only with edge-major fp32 logits (or ONE logits row broadcast over the batch: row stride
 * 0), an edge-major uint8 one-hot and B in {4, 8, 16, 32k}
 * (the layout MPNNPolicyNet emits); TARL_E_BADARG otherwise — callers then use tarl_graphdist_forward. */
int tarl_graphdist_sample(const tarl_csr* groups, const tarl_rows* logits, float temperature, int32_t batch,
                          const tarl_rows* uniforms, const tarl_rows* onehot, int32_t onehot_dtype, float* log_prob,
                          float* partials, void* stream);

/* Rollout form of the two calls a policy step makes — GraphDistribution.sample (src/reinforcement_learning.py:57-80)
 * and the action write of SimulatorEnv._step, x[edge_index[0][mask], SELECTED_ROAD] = edge_index[1][mask] (:223-231) —
 * for `batch` replicas that share ONE logits row (logits_row [E]; the active policy path does not read the dynamic
 * observation). Draws one edge per (replica, source group) from uniforms [batch, K], writes the byte one-hot
 * EDGE-major (element (b, e) at e*batch + b; batch % 4 == 0, 4-byte aligned), optionally the log-probability of the
 * draw (log_prob [batch], partials as for tarl_graphdist_sample), and stores the target node of the drawn edge into
 * SELECTED_ROAD of the group's source node: sel_links [batch, n_links] for road links (node id < n_links),
 * sel_sources [batch, n_nodes - n_links] for the other nodes; group_node [K] = source node of each group, edge_dst
 * [E] = edge_index[1] as int32. A group without a hit (uniform >= the rounded cumulative sum) leaves SELECTED_ROAD
 * untouched and makes the row's log-probability -inf, exactly as the two separate calls do. uniforms == NULL: the
 * uniforms are drawn in the kernel (Philox4x32-10 keyed by `seed`, counter (group, 4-row chunk)) instead of being
 * produced by torch.rand and read back — the reference draws them from torch's global generator (:62), so which
 * stream they come from is declared divergence D4 either way. The counter is (group, (row_offset + row) / 4, draw_id):
 * row_offset = global index of this call's first row (replicas sharded over ranks draw different uniforms from one
 * seed), draw_id = which draw of the seed this is (the step of a rollout); seed_dev: NULL, or a device word holding the
 * key instead of `seed` (a call captured in a CUDA graph then draws differently on every replay). row_offset % 4 == 0.
 * prev_links / prev_sources: NULL, or the SELECTED_ROAD arrays of the previous step when sel_links / sel_sources are a
 * SECOND pair of buffers (a rollout alternates two pairs so that the draw of step t+1 can run while step t still reads
 * its decisions): a group without a hit then copies its previous value instead of leaving the entry untouched. */
int tarl_graphdist_sample_apply(const tarl_csr* groups, const float* logits_row, float temperature, int32_t batch,
                                const tarl_rows* uniforms, uint8_t* onehot, float* log_prob, float* partials,
                                const int32_t* group_node, const int32_t* edge_dst, float* sel_links, float* sel_sources,
                                const float* prev_links, const float* prev_sources,
                                int32_t n_links, int32_t n_nodes, uint64_t seed, const uint64_t* seed_dev,
                                uint32_t draw_id, int32_t row_offset, void* stream);

/* MPNNValueNet's propagate (src/agents/mpnn_agent.py:300-402, dropout off): per node x = [node_features(7) ‖
 * agent_features[agent_index](9)]; per edge e of the FULL graph msg = tanh(w·[x[edge_index[1][e]] ‖ edge_features[e]]
 * + w0); mean over the edges sharing edge_index[0]; v = tanh(a*mean + c). by_source / by_target: CSR of edge_index by
 * source / target NODE (n_rows = n_nodes), idx = the other endpoint, eid ascending inside a row.
 * node_features [B,N,>=7] (element strides given), edge_features [B,E] with batch stride ef_batch_stride (0 = one
 * row shared by the whole batch), agent_index [B,N] int64, agent_features [agent_rows, 9]. msg_weight [17]
 * (= message_mlp.1.weight), msg_bias [1], node_weight [1], node_bias [1] are device pointers. Outputs proj, mean, v
 * are NODE-major with the batch row innermost: element (b, n) at n*B + b (proj and mean are what backward needs).
 * agent_proj: [agent_rows] scratch (the agent part of the projection, msg_weight[7:16] . agent row, computed once per
 * agent row and gathered per node). */
int tarl_value_mp_forward(const tarl_csr* by_source, const float* node_features, int64_t nf_batch_stride,
                          int64_t nf_row_stride, const float* edge_features, int64_t ef_batch_stride,
                          const int64_t* agent_index, const float* agent_features, int32_t agent_rows,
                          const float* msg_weight, const float* msg_bias, const float* node_weight,
                          const float* node_bias, int32_t batch, int32_t n_nodes, float* agent_proj, float* proj,
                          float* mean, float* v, int32_t* flags, void* stream);

/* Gradient of sum(grad_v * v) w.r.t. the four parameter tensors: grads[0:17] = d msg_weight, [17] = d msg_bias,
 * [18] = d node_weight, [19] = d node_bias. grad_v: element (b, n) at b*gv_batch_stride + n*gv_node_stride.
 * gm: [N*B] scratch; partials: 20*tarl_value_mp_partial_count(N,B) floats. Fixed summation order (deterministic). */
int32_t tarl_value_mp_partial_count(int32_t n_nodes, int32_t batch);
int tarl_value_mp_backward(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                           int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                           int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                           int32_t agent_rows, const float* msg_weight, const float* msg_bias, const float* node_weight,
                           int32_t batch, int32_t n_nodes, const float* proj, const float* mean, const float* v,
                           const float* grad_v, int64_t gv_batch_stride, int64_t gv_node_stride, const float* head_g,
                           const float* head_w, float* gm,
                           float* partials, float* grads, void* stream);

/* The node part of MPNNValueNet's head (src/agents/mpnn_agent.py:359-361: final_mlp over [v ‖ time_net(t)]):
 * out[b] = sum_n v[n, b] * head_weight[n] on the node-major v the propagate returns, fixed summation order; partials:
 * tarl_value_head_partial_count(n_nodes) * batch floats. Its backward never materialises grad_v: both backward entry
 * points take head_g [B] (= d loss / d out) and head_w [N] in place of grad_v (which may then be NULL) and use
 * grad_v[b, n] = head_g[b] * head_w[n]; tarl_value_head_weight_grad gives d head_weight[n] = sum_b head_g[b] * v[n, b]. */
int32_t tarl_value_head_partial_count(int32_t n_nodes);
int tarl_value_head_forward(const float* v, int32_t batch, int32_t n_nodes, const float* head_weight, float* partials,
                            float* out, void* stream);
int tarl_value_head_weight_grad(const float* v, int32_t batch, int32_t n_nodes, const float* grad_out, float* grad_weight,
                                void* stream);

/* The same propagate in TRAIN mode: nn.Dropout(p) on the [B*E, 17] message input (src/agents/mpnn_agent.py:278,
 * 385-386; ATen computes x * (mask / (1 - p))). Every (row, edge) pair has a 17-bit keep word (bit k = input k
 * survives; k < 16 the target node's inputs, k = 16 the edge feature): either injected — keep_bits [B, E] with batch
 * stride keep_batch_stride, e.g. the mask the reference itself drew — or, with keep_bits == NULL, drawn in the kernel
 * from Philox4x32-10 keyed by `seed` (tarl_value_mp_dropout_bits writes the words of that stream: what the kernels
 * will use for the same seed / p). keep_words: NULL, or [E*B] scratch (element (b, e) at e*B + b): with keep_bits ==
 * NULL the forward pass stores the words it draws there and the backward pass reads them back instead of drawing them
 * again. msg: [E*B] output (the tanh messages; backward reads them back and overwrites them). Both are kept in
 * BY-TARGET order: element (b, e) at pos(e)*B + b, pos(e) = position of edge e in by_target (a target node's in-edges are
 * contiguous: the backward pass finds them without loading an edge id first). source_pos: [E] int32, entry j = pos of
 * the j-th edge of by_source (static, built once per graph). mean, v as in tarl_value_mp_forward.
 * agent_pack: NULL, or 16-byte aligned scratch of agent_rows * 12 floats: both passes re-lay agent_features there as
 * 48-byte rows (three 128-bit loads per (node, row) pair instead of nine scalar gathers).
 * In-kernel stream for p <= 1/16: ONE Philox block per (row, edge) — 17 top nibbles decide 15 of 16 inputs, the rare
 * zero nibbles take a low byte from the block's 7 spare bytes; larger p: two blocks, 17 twelve-bit fields. */
int tarl_value_mp_dropout_bits(uint64_t seed, float p, int32_t batch, int32_t n_edges, uint32_t* keep_bits, void* stream);
int tarl_value_mp_forward_dropout(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                                  int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                                  int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                                  int32_t agent_rows, const float* msg_weight, const float* msg_bias,
                                  const float* node_weight, const float* node_bias, int32_t batch, int32_t n_nodes,
                                  const uint32_t* keep_bits, int64_t keep_batch_stride, uint64_t seed, float p,
                                  const int32_t* source_pos, uint32_t* keep_words, float* agent_pack, float* msg,
                                  float* mean, float* v, int32_t* flags, void* stream);
/* grads / gm / partials as in tarl_value_mp_backward; keep_bits / seed / p must be the forward call's. msg is
 * OVERWRITTEN (it becomes d z, the gradient at the tanh's argument). */
int tarl_value_mp_backward_dropout(const tarl_csr* by_source, const tarl_csr* by_target, const float* node_features,
                                   int64_t nf_batch_stride, int64_t nf_row_stride, const float* edge_features,
                                   int64_t ef_batch_stride, const int64_t* agent_index, const float* agent_features,
                                   int32_t agent_rows, const float* node_weight, int32_t batch, int32_t n_nodes,
                                   const uint32_t* keep_bits, int64_t keep_batch_stride, uint64_t seed, float p,
                                   const uint32_t* keep_words, float* agent_pack, float* msg, const float* mean,
                                   const float* v, const float* grad_v,
                                   int64_t gv_batch_stride, int64_t gv_node_stride, const float* head_g,
                                   const float* head_w, float* gm, float* partials,
                                   float* grads, void* stream);

/* MPNNPolicyNet's per-edge MLPs (src/agents/mpnn_agent.py:30-50; their only use in the reference are the two
 * commented-out bodies of update_edges, :220-231): with x[b,n] = [node_features[b,n,0:7] ‖ agent_features[agent_index
 * [b,n], 0:9]] (:163-167),
 *   TARL_EDGE_MLP       logit[b,e] = L3(relu(L2(relu(L1([x[b,src e] ‖ x[b,dst e] ‖ edge_attr[b,e]])))))   33 -> 64 -> 32 -> 1
 *   TARL_EDGE_MLP_TEST  logit[b,e] = L2(relu(L1([x[b,src e] ‖ x[b,dst e]])))                              32 -> 16 -> 1
 * tarl_edge_mlp_inputs assembles x ([B, N, 16] fp32, contiguous, 16-byte aligned) from the observation. weights: the
 * module's parameters in Sequential order, row-major as nn.Linear stores them — {W1, b1, W2, b2, W3, b3} (6 pointers)
 * or {W1, b1, W2, b2} (4). out / grad_out: element (b, e) at b*batch_stride + e*edge_stride. edge_attr: [B, E] with
 * the given batch stride (0 = one row shared by all b); unused by TARL_EDGE_MLP_TEST.
 * Forward: with tarl_edge_mlp_tc_available() != 0 and variant TARL_EDGE_MLP both hidden layers run on the tensor cores
 * (tcgen05.mma kind::tf32, 3xTF32, activations in TMEM; csrc/edge_mlp_tc.cu; needs tc_scratch) unless
 * use_tensor_cores == 0; otherwise on the fp32 pipe. Backward: parameter gradients only (the observation is a leaf), flat in Sequential.parameters()
 * order — tarl_edge_mlp_param_count(variant) floats; partials: scratch of tarl_edge_mlp_partial_count() x that many
 * floats; fixed summation order (deterministic). */
#define TARL_EDGE_MLP 0
#define TARL_EDGE_MLP_TEST 1
int32_t tarl_edge_mlp_param_count(int32_t variant);
int32_t tarl_edge_mlp_partial_count(void);
int32_t tarl_edge_mlp_tc_available(void);
int32_t tarl_edge_mlp_tc_scratch_floats(void);   /* tc_scratch of the forward call: that many floats, 16-byte aligned */
int tarl_edge_mlp_inputs(const float* node_features, int64_t nf_batch_stride, int64_t nf_row_stride,
                         const int64_t* agent_index, const float* agent_features, int32_t agent_rows, int32_t batch,
                         int32_t n_nodes, float* x, int32_t* flags, void* stream);
int tarl_edge_mlp_forward(int32_t variant, const int32_t* edge_src, const int32_t* edge_dst, int32_t n_edges,
                          const float* x, int32_t batch, int32_t n_nodes, const float* edge_attr, int64_t ea_batch_stride,
                          const float* const* weights, int32_t use_tensor_cores, float* tc_scratch, float* out,
                          int64_t out_batch_stride, int64_t out_edge_stride, void* stream);
int tarl_edge_mlp_backward(int32_t variant, const int32_t* edge_src, const int32_t* edge_dst, int32_t n_edges,
                           const float* x, int32_t batch, int32_t n_nodes, const float* edge_attr, int64_t ea_batch_stride,
                           const float* const* weights, const float* grad_out, int64_t go_batch_stride,
                           int64_t go_edge_stride, float* partials, float* grads, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Population operations either side of the core step (csrc/agents.cu). Each works on either state layout.
 * ------------------------------------------------------------------------------------------------------------- */

/* Which road state to operate on: EITHER the reference rows (x != NULL: graph.x, element (r, n, c) at
 * x[r*x_replica_stride + n*x_row_stride + c], ALL n_nodes rows because SRC nodes carry SELECTED_ROAD too) OR a link
 * store (store != NULL; then the SELECTED_ROAD of the non-road nodes lives in src_sel [R, n_nodes - n_links]). */
typedef struct tarl_agent_state {
    float* x;
    int64_t x_row_stride;
    int64_t x_replica_stride;
    int32_t n_links;    /* N = graph.num_roads (rows layout; taken from the store otherwise)          */
    int32_t nmax;
    int32_t n_replicas;
    int32_t n_nodes;    /* N_tot = graph.x.size(0)                                                    */
    const float* cc;    /* rows layout: congestion_constant[:N] or NULL (hasattr test, base.py:314)    */
    const tarl_link_store* store;
    float* src_sel;
    float t_garbage;    /* store layout: time of the latest tarl_store_step (see tarl_store_export)    */
    int32_t reserved;
} tarl_agent_state;

/* agent_features (src/feature_helpers.py:56-71): fp32 [R, n_rows, 9] row-major, replica r at +r*replica_stride
 * (ignored when there is one replica). Columns: ORIGIN, DESTINATION, DEPARTURE_TIME, ARRIVAL_TIME, AGE, SEX,
 * EMPLOYMENT_STATUS, ON_WAY, DONE. Row 0 is the dummy agent. */
typedef struct tarl_agent_table {
    float* agent_features;
    int64_t replica_stride;
    int32_t n_rows;
    int32_t reserved;
} tarl_agent_table;

/* The population indexed by ORIGIN node, built once per population by the host (ORIGIN is static, and identical in
 * every replica): agents of node o = org_agent[org_ptr[o] .. org_ptr[o+1]) in ascending agent id; origins = the nodes
 * that own at least one agent. */
typedef struct tarl_agent_index {
    int32_t n_nodes;
    int32_t n_origins;
    const int32_t* org_ptr;   /* [n_nodes+1] */
    const int32_t* org_agent; /* [n_rows]    */
    const int32_t* origins;   /* [n_origins] */
    const float* dep_sorted;  /* [n_rows] or NULL: DEPARTURE_TIME of the agents of node o in ASCENDING order at
                                 [org_ptr[o] .. org_ptr[o+1]) (static and identical in every replica); lets
                                 tarl_agents_insert count the departed agents of an origin without reading their rows */
} tarl_agent_index;

/* Replaces Agents.insert_agent_into_network (src/agents/base.py:244-331): every agent with DEPARTURE_TIME <= t,
 * ON_WAY == 0 and DONE == 0 targets road x[ORIGIN, SELECTED_ROAD]; per road the first min(count, MAXN-3-NUM) of them
 * in ascending agent id (declared divergence D3) are appended at the tail with arrival t and exit time
 * t + max(FFTT, cc/(MAXN+10-NUM_before)); NUM += admitted; ON_WAY = 1.
 * Scratch (caller-owned int32): head [R*N] initialised to -1 ONCE by the caller (the kernels leave it at -1),
 * next [R*n_origins], cursor [R*n_origins]. counters: NULL or [R*2] {inserted, withdrawn} running totals.
 * inserted: NULL, or [R*n_origins] agents inserted so far per (replica, origin) — maintained here, zeroed by the caller
 * whenever it resets ON_WAY / DONE, valid only while nothing else edits those columns; with index->dep_sorted it lets
 * origins without a waiting agent be skipped without touching agent_features (identical results).
 * worklist / work_count: both NULL, or scratch [R*n_origins] / [R + n_origins]: the origins that queue for a road this step are
 * compacted per replica and the admission runs over that list only (identical results; the listed origins are a few
 * per cent of all (replica, origin) pairs and each carries a chain of dependent gathers).
 * num_out / occupancy: both NULL, or (link store only) the occupancy observation tarl_agents_withdraw left behind in
 * this very step — num_out [R, n_nodes], occupancy [R] — which the roads that admit agents patch in place.
 * road_origin: NULL, or [n_links] int32 for networks in which every road can be selected by ONE origin only
 * (road_origin[n] = index into index->origins of that origin, -1 = none; config_network's graphs: a road leaves one
 * intersection, whose SRC node alone has an edge into it). With it (and inserted + index->dep_sorted) the two phases
 * run as ONE kernel without the per-road lists — identical results; an origin whose SELECTED_ROAD names another
 * origin's road while it has a ready agent raises TARL_ERR_INSERT_TARGET instead of being served. */
int tarl_agents_insert(const tarl_agent_state* state, const tarl_agent_table* agents, const tarl_agent_index* index,
                       float t, int32_t* head, int32_t* next, int32_t* cursor, int32_t* counters, int32_t* inserted,
                       int32_t* flags, int32_t* worklist, int32_t* work_count, float* num_out, int32_t* occupancy,
                       const int32_t* road_origin, void* stream);

/* Replaces Agents.withdraw_agent_from_network (src/agents/base.py:334-403): per link the maximal prefix of queue
 * slots k < NUM whose exit time <= t and whose agent's DESTINATION node is adjacent to the link — adjacency = CSR of
 * the FULL edge_index by source node (idx = targets), the sparse form of adj_matrix[ROAD_INDEX, DESTINATION] — is
 * removed; the three queue segments shift left by that count with zero fill; the agents get DONE = 1, ON_WAY = 0,
 * ARRIVAL_TIME = t. mask: NULL or [R*N] (the entry of withdraw_history).
 * num_out / occupancy: both NULL, or (link store only) num_out [R, n_nodes] fp32 receives NUMBER_OF_AGENT of every node
 * after the withdrawal (0 for the non-road nodes) and occupancy [R] their sum — the observation / reward of
 * SimulatorEnv._step (src/reinforcement_learning.py:266) for nets that read the occupancy only, produced by the pass
 * that already holds every record; pass the same two pointers to the tarl_agents_insert call that follows. */
int tarl_agents_withdraw(const tarl_agent_state* state, const tarl_agent_table* agents, const tarl_csr* adjacency,
                         float t, uint8_t* mask, int32_t* counters, int32_t* flags, float* num_out, int32_t* occupancy,
                         void* stream);

/* tarl_store_step followed by tarl_agents_withdraw(num_out, occupancy) at the same t as ONE pass over the records: the
 * response phase of the core step ends with every link's final record in registers, so the same thread withdraws
 * (src/agents/base.py:334-403) where the head is due, writes NUMBER_OF_AGENT into num_out [R, n_nodes] (0 for the
 * non-road nodes; needs n_nodes - n_links <= n_links) and adds it into occupancy [R] (zeroed here). Results are those
 * of the two separate calls. ELLPACK topology and a store in link-id order only (TARL_E_BADARG otherwise: make the two
 * calls). withdrawn: [R*N] mask, counters: NULL or [R*2]. Follow with tarl_agents_insert(num_out, occupancy). */
int tarl_store_step_withdraw(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                             const float* attr_in, const tarl_step_io* io, const tarl_agent_table* agents,
                             const tarl_csr* adjacency, int32_t n_nodes, uint8_t* withdrawn, int32_t* counters,
                             float* num_out, int32_t* occupancy, void* stream);

/* Replaces Agents.choice (src/agents/base.py:446-494): every node listed in choosers (roads with a downstream road,
 * SRC nodes with an outgoing road) draws one of its neighbours[node] (ascending road id) uniformly into
 * SELECTED_ROAD: neighbour number min(floor(u*deg), deg-1). uniforms: [R*n_choosers] or NULL for the in-kernel Philox
 * stream (seed, step_id) — the reference draws torch.multinomial from the global generator (declared divergence D4). */
int tarl_agents_choice(const tarl_agent_state* state, const tarl_csr* neighbours, const int32_t* choosers,
                       int32_t n_choosers, const float* uniforms, uint64_t seed, uint32_t step_id, void* stream);

/* Replaces the action write of SimulatorEnv._step (src/reinforcement_learning.py:223-231):
 * SELECTED_ROAD[edge_src[e]] = edge_dst[e] for every e of the FULL graph with action[r, e] != 0. action: [R, E_full]
 * with any strides (what GraphDistribution.sample returns is edge-major). */
int tarl_agents_apply_action(const tarl_agent_state* state, const int32_t* edge_src, const int32_t* edge_dst,
                             int32_t n_edges, const tarl_rows* action, int32_t action_dtype, void* stream);

/* The same write with the edges grouped by source node: groups = CSR of the FULL edge_index by source RANK (ptr
 * [K+1], eid [E_full] ascending inside a group; idx unused), group_node [K] = node id of each group. Tiled so that
 * both the action reads (replica innermost) and the SELECTED_ROAD writes (node innermost) are coalesced; a group with
 * several selected edges keeps the last one in ascending edge id, one without leaves SELECTED_ROAD untouched. */
int tarl_agents_apply_action_groups(const tarl_agent_state* state, const tarl_csr* groups, const int32_t* group_node,
                                    const int32_t* edge_dst, const tarl_rows* action, int32_t action_dtype, void* stream);

/* state() (src/transportation_simulator.py:360-366) and the reward term of SimulatorEnv._step
 * (src/reinforcement_learning.py:266) from a link store: node_features [R, n_nodes, 7] = {MAXN, NUM, FFTT, LENGTH,
 * MAX_FLOW, SELECTED_ROAD, ROAD_INDEX}, agent_index [R, n_nodes] int64 = head agent ids, occupancy [R] int32 =
 * sum of NUM over the links (integer atomics); num_agents / selected_road [R, n_nodes] = the NUMBER_OF_AGENT and
 * SELECTED_ROAD columns alone (the compact observation a rollout stores per frame). Any output may be NULL. */
int tarl_store_observe(const tarl_agent_state* state, float* node_features, int64_t* agent_index, int32_t* occupancy,
                       float* num_agents, float* selected_road, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Metrics side channels, accumulated on the device (csrc/metrics.cu).
 *
 * Replaces the per-step histories behind TransportationSimulator.compute_node_metrics / plot_daily_counts /
 * plot_road_optimality (src/transportation_simulator.py:351,453-510,563-669): the reference keeps one bool[N] per
 * step in ResponseMPNN.update_history (src/response_mpnn.py:125) and Agents.withdraw_history
 * (src/agents/base.py:402) plus a host copy of delta_travel_time[E] per step, and reduces them afterwards
 * (hour = time // 3600; per-link aggregate = scatter_add over edge_index_routes[0]).
 *   counts[r, hour, n]         += (pop[r, n] != 0) + (withdrawn[r, n] != 0)          int32 [R, n_hours, N]
 *   optimality_now[r, n]        = sum of delta_tt[r, e] over the out-edges e of n, ascending edge id   [R, N]
 *   optimality_sum[r, hour, n] += that sum                                           fp32 [R, n_hours, N]
 * pop / withdrawn: uint8 [R, N] or NULL; delta_tt: fp32 [R, E] in original edge order or NULL; any output may be
 * NULL. g supplies N, E and the out-edge segments (out_ptr, out_eid). 0 <= hour < n_hours. */
int tarl_metrics_accumulate(const tarl_dual_csr* g, int32_t n_replicas, const uint8_t* pop, const uint8_t* withdrawn,
                            const float* delta_tt, int32_t hour, int32_t n_hours, int32_t* counts,
                            float* optimality_sum, float* optimality_now, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * The PPO update's non-network arithmetic on the device (csrc/optim.cu).
 * ------------------------------------------------------------------------------------------------------------- */

/* Generalized advantage estimation over a [T, R] trajectory — what torchrl's GAE(gamma, lmbda, average_gae=True) does at
 * src/rl/ppo_trainer.py:21-27,132 of the reference: delta_t = r_t + gamma V(s_t+1) (1 - terminated_t) - V(s_t);
 * A_t = delta_t + gamma lmbda (1 - done_t) A_t+1; value_target = A + V. value / next_value: element (t, r) at
 * [t*value_step_stride + r] (a [T+1, R] array of V over the frames serves both: next_value = value + stride).
 * reward [T, R] fp32, done / terminated [T, R] bytes. partials: 2*tarl_gae_partial_count(R) doubles = per-CTA {sum A,
 * sum A^2} in a fixed order (the caller adds them up, all-reduces over the ranks and calls tarl_standardise). */
int32_t tarl_gae_partial_count(int32_t n_replicas);
int tarl_gae(const float* value, const float* next_value, int64_t value_step_stride, const float* reward,
             const uint8_t* done, const uint8_t* terminated, int32_t n_steps, int32_t n_replicas, float gamma, float lmbda,
             float* advantage, float* value_target, double* partials, void* stream);

/* advantage <- (advantage - mean) / max(std, 1e-4) with the unbiased std, from stats = {count, sum, sum of squares}
 * (device doubles: no host round trip between the reduction and its use). */
int tarl_standardise(float* advantage, int64_t n, const double* stats, void* stream);

/* ClipPPOLoss as the reference wires it (src/rl/ppo_trainer.py:36: torchrl 0.5.0 ClipPPOLoss, clip_epsilon 0.2,
 * entropy_coef 0.01, critic_coef 1.0, loss_critic_type "smooth_l1", normalize_advantage False) for the n frames of one
 * minibatch, forward and backward in one launch — what the update loop evaluates at :135-139 (loss.backward() included):
 *   ratio = exp(log_prob - sample_log_prob)
 *   loss_objective = -mean(min(ratio A, clamp(ratio, 1 - clip, 1 + clip) A))
 *   loss_entropy = -entropy_coef mean(entropy);   loss_critic = critic_coef mean(smooth_l1(value - value_target))
 * out[7] = {loss_objective, loss_entropy, loss_critic, approx_kl = mean(sample_log_prob - log_prob),
 *           clip_fraction = mean(|ratio - 1| > clip), entropy = mean(entropy), number of impossible frames}.
 * An "impossible" frame is one whose log_prob AND sample_log_prob are both -inf — GraphDistribution.log_prob's marker
 * for an action that selects no edge in some group (src/reinforcement_learning.py:82-93), which sample() (:57-80)
 * produces about once in 10^8 draws. Its ratio exp(-inf - -inf) is NaN in the torch formula and poisons every
 * parameter at the next optimiser step; here it contributes nothing to the objective, approx_kl and clip_fraction
 * (zero gradient w.r.t. its log_prob), still counts in the critic and entropy terms and in every denominator (declared
 * divergence D8). Any other non-finite input propagates as in torch.
 * grad_log_prob / grad_entropy / grad_value [n]: gradient of loss_objective + loss_critic + loss_entropy with respect
 * to the three differentiable inputs (torch's subgradient conventions: minimum splits ties, clamp passes the gradient
 * on its closed interval). All arrays contiguous fp32 on the device; n >= 1. */
int tarl_ppo_clip_loss(const float* log_prob, const float* sample_log_prob, const float* advantage, const float* entropy,
                       const float* value, const float* value_target, int32_t n, double clip_epsilon, float entropy_coef,
                       float critic_coef, float* out, float* grad_log_prob, float* grad_entropy, float* grad_value,
                       void* stream);

/* One torch.optim.Adam step (src/rl/ppo_trainer.py:39,142: lr 1e-3, betas 0.9 / 0.999, eps 1e-8, no weight decay) over
 * a flat fp32 bucket of n parameters — the bucket the gradient all-reduce works on. step: 1 for the first update.
 * grad_scale multiplies every gradient first (1 / world size after a summing all-reduce). grad_norm: NULL, or a device
 * float receiving the global L2 norm of the scaled gradient (the value the reference logs at :141); then
 * norm_partials must hold tarl_adam_partial_count(n) doubles. */
int32_t tarl_adam_partial_count(int64_t n);
int tarl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, int32_t step, float grad_scale, double* norm_partials, float* grad_norm,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MPNNValueNetSimple forward (tensor cores) and backward (csrc/value_mlp.cu: tcgen05.mma kind::tf32 with 3xTF32 error
 * compensation, operands staged by TMA, A split into TMEM, fp32 accumulation in TMEM).
 *
 * Replaces MPNNValueNetSimple.forward (src/agents/mpnn_agent.py:428-450) for n_rows observations at once:
 *   out[m] = w3 . relu(W2 relu(W1 [occupancy[m, 0:n_nodes] ‖ time[m]] + b1) + b2) + b3
 * occupancy: fp32, element (m, n) at occupancy[m*occ_row_stride + n] = NUMBER_OF_AGENT of node n in observation m
 *   (16-byte aligned base, occ_row_stride a multiple of 4: TMA addressing; anything else is TARL_E_BADARG and the
 *   caller uses its library GEMM). time: element m at time[m*time_stride].
 * w1 [64, n_nodes+1], b1 [64], w2 [64, 64], b2 [64], w3 [64], b3 [1]: final_mlp.{0,2,4}.{weight,bias}, row-major.
 * weights_changed: 0 = w1 is what the previous call on this workspace (same n_rows, n_nodes) was given, so its TF32
 *   hi/lo split kept in the workspace is reused; anything else re-splits.
 * workspace: tarl_value_mlp_workspace_bytes(n_rows, n_nodes) bytes, 1024-byte aligned. out: [n_rows].
 * save_z1 / save_z2: NULL (inference), or [n_rows, 64] each: the pre-activations of the two hidden layers, which
 * tarl_value_mlp_backward needs (training-mode forward of the PPO update, src/rl/ppo_trainer.py:132-145). */
size_t tarl_value_mlp_workspace_bytes(int32_t n_rows, int32_t n_nodes);
int tarl_value_mlp_forward(const float* occupancy, int64_t occ_row_stride, const float* time, int64_t time_stride,
                           int32_t n_rows, int32_t n_nodes, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int32_t weights_changed, void* workspace,
                           size_t workspace_bytes, float* out, float* save_z1, float* save_z2, void* stream);

/* Gradient of sum_m grad_out[m] * out[m] w.r.t. final_mlp.{0,2,4}.{weight,bias} (what autograd computes through the three
 * nn.Linear of src/agents/mpnn_agent.py:428-450; the observation is a leaf, so there is no input gradient):
 * grad_w1 [64, n_nodes+1] (last column = the time input), grad_b1 [64], grad_w2 [64, 64], grad_b2 [64], grad_w3 [64],
 * grad_b3 [1], all overwritten. z1 / z2: what the forward call saved. scratch: (2*n_rows + 1) * 64 floats.
 * dW1 = g_z1^T A reads the occupancy matrix once and writes the gradient once (16 flop per byte, reduction depth
 * n_rows: HBM-bound on the fp32 pipe, exact fp32 sums in ascending row order); any row pitch >= n_nodes. */
int tarl_value_mlp_backward(const float* occupancy, int64_t occ_row_stride, const float* time, int64_t time_stride,
                            int32_t n_rows, int32_t n_nodes, const float* w2, const float* w3, const float* z1,
                            const float* z2, const float* grad_out, float* scratch, float* grad_w1, float* grad_b1,
                            float* grad_w2, float* grad_b2, float* grad_w3, float* grad_b3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TARL_B200_H */

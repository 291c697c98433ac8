// engine.cu — the resident link store: the per-timestep network step of core_step.cu on a compact device layout that
// moves ~1x the algorithmic bytes instead of ~4x (sm_100a). Compiled with -fmad=false.
//
// Why: on the reference's 208-byte AoS rows (Nmax=15) the step has to read every row whole, write 4 partial sectors
// per link for the mandatory tail write, and shift 3*(Nmax-1) floats per popped link (profiles/r01_a: 796 MB of DRAM
// traffic per step for 213 MB algorithmic at 1M links). The store keeps, per (replica, link),
//
//   hot   32 B  A = {head id, head exit, NUM, MAXN} | B = {head arrival, tail id, pending-garbage exit time, meta}
//               read once and rewritten whole every step (full-sector writes, ping-pong buffers: the direction phase
//               gathers the PRE-step A halves of upstream links while owners write POST-step records elsewhere)
//   sel    4 B  SELECTED_ROAD (the per-step routing input, its own array so that choice/actions write it coalesced)
//   queue 16 B x (Nmax-1) ring of {id, arrival, exit} for logical FIFO slots 1..Nmax-1 (slot 0 lives in `hot`),
//               touched only by real admissions (one slot write) and pops (one slot read + one slot copy)
//   post  16 B  {NUM, tail id, head id after the direction phase, delta_travel_time} for the response phase
//
// and reproduces the reference's x EXACTLY on export, including its quirks: the tail triplet (0, t, t+tt) it writes
// past the tail of every link every step is kept as one pending record per link ("garbage at logical slot int(NUM)",
// always overwritten in place by the next step or consumed by an export), and its shift-left that duplicates the
// last slot becomes a ring-head increment plus one slot copy. Semantics: SURVEY.md Appendix A.
//
// Two kernel families, bit-identical results:
//   *_csr  one thread per link walking its edge segment of the CSR (in_ptr -> in_src -> record: three dependent loads
//          per edge; any degree);
//   *_ell  the same thread reads its first W edges from an ELLPACK copy of the topology (column j of link d at
//          [j*pitch + d]): all W edge slots load coalesced and INDEPENDENTLY of any pointer array, so the dependent
//          chain is two levels deep (edge slots -> neighbour records) with 2W gathers in flight per thread. Road
//          networks have near-uniform small degree, which is exactly the case ELL is made for; links with more than W
//          edges fall back to their CSR segment.
#include <stdlib.h>

#include "engine_common.cuh"
#include "population.cuh"

using namespace tarl;


namespace {

// ------------------------------------------------------------------------------------------------ import / export
__global__ void __launch_bounds__(kThreads) k_store_import(Store s, const float* __restrict__ x, int64_t row_stride,
                                                           int64_t rep_stride, const float* __restrict__ cc,
                                                           float4* __restrict__ hot, float4* __restrict__ stat_a,
                                                           float4* __restrict__ stat_b, int32_t* __restrict__ flags) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), n = (int)(L % s.N);      // n = store slot
    const int link = s.link_of(n);
    const float* row = x + r * rep_stride + (int64_t)link * row_stride;
    const int Nmax = s.Nmax, c0 = 3 * Nmax;
    const float maxn = row[c0], num = row[c0 + 1], fftt = row[c0 + 2];
    const bool bad = !(num >= 0.0f) || !(num <= (float)Nmax);
    if (bad) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
    const int cnt = bad ? 0 : (int)num;
    hot[2 * L] = make_float4(row[0], row[2 * Nmax], num, maxn);
    hot[2 * L + 1] = make_float4(row[Nmax], row[max(cnt - 1, 0)], 0.0f, __int_as_float(0));
    s.sel[L] = row[c0 + 5];
    for (int k = 1; k <= s.M; ++k)
        s.queue[L * s.M + (k - 1)] = make_float4(row[k], row[Nmax + k], row[2 * Nmax + k], 0.0f);
    if (r == 0) {
        float ccn;
        if (cc != nullptr) ccn = cc[link];
        else ccn = fftt * ((maxn + 10.0f) - (row[c0 + 4] * fftt) / 3600.0f);   // src/simulation_core_model.py:60-67
        stat_a[n] = make_float4(fftt, ccn, row[c0 + 6], __int_as_float(0x7fc00000));    // .w: the caller's hint, NaN = none
        stat_b[n] = make_float4(row[c0 + 3], row[c0 + 4], 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(kThreads) k_store_export(Store s, float* __restrict__ x, int64_t row_stride,
                                                           int64_t rep_stride, float t_garbage) {
    const int64_t L = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (int64_t)s.N * s.R) return;
    const int r = (int)(L / s.N), n = (int)(L % s.N);      // n = store slot
    float* row = x + r * rep_stride + (int64_t)s.link_of(n) * row_stride;
    const int Nmax = s.Nmax, c0 = 3 * Nmax;
    const float4 hA = s.hot_cur[2 * L], hB = s.hot_cur[2 * L + 1];
    const int meta = __float_as_int(hB.w);
    const int rh = meta & kMetaRingMask;
    const int gslot = (meta & kMetaGarbage) ? (int)hA.z : -1;
    row[0] = hA.x; row[Nmax] = hB.x; row[2 * Nmax] = hA.y;
    for (int k = 1; k <= s.M; ++k) {
        float4 v = s.queue[L * s.M + ring_pos(rh, k, s.M)];
        if (k == gslot) v = make_float4(0.0f, t_garbage, hB.z, 0.0f);
        row[k] = v.x; row[Nmax + k] = v.y; row[2 * Nmax + k] = v.z;
    }
    const float4 a = s.stat_a[n], b = s.stat_b[n];
    row[c0] = hA.w; row[c0 + 1] = hA.z; row[c0 + 2] = a.x; row[c0 + 3] = b.x; row[c0 + 4] = b.y;
    row[c0 + 5] = s.sel[L]; row[c0 + 6] = a.z;
}

// ------------------------------------------------------------------------------------------------ direction phase
struct Noise {
    const float* ext;        // [R*E] uniforms in original edge order, or nullptr for the in-kernel Philox stream
    uint32_t seed_lo, seed_hi, step_id;
    const unsigned long long* seed_dev;   // nullptr, or where the Philox key lives on the device (replaces seed_lo / hi:
                                          // a launch captured in a CUDA graph can then draw differently on every replay)
};

__device__ __forceinline__ void philox_key(const Noise& nz, uint32_t& lo, uint32_t& hi) {
    lo = nz.seed_lo; hi = nz.seed_hi;
    if (nz.seed_dev != nullptr) {
        const unsigned long long k = *nz.seed_dev;
        lo = (uint32_t)k; hi = (uint32_t)(k >> 32);
    }
}

// p_e of src/direction_mpnn.py:81-91 for the edge u -> d. U = A half of the upstream record (pre-step).
__device__ __forceinline__ float edge_prob(const float4 U, float sel_u, float t, bool free_d, float room_d, float ridx_d,
                                           float attr) {
    const bool a1 = (U.y <= t) && (U.z > 0.0f);
    const bool a2 = ((U.y - t) < -10.0f) && ((U.w - 3.0f) <= U.z);
    const bool match = (sel_u == ridx_d);
    const bool m = (a1 && free_d && match) || (a2 && ((U.w - U.z) <= room_d) && match);
    return attr * (m ? 1.0f : 0.0f);
}

__device__ __forceinline__ float gumbel_score(float p, float u) {     // src/direction_mpnn.py:137-138
    return logf(p + 1e-12f) + (-logf(-logf(u)));
}

struct Pick {
    float psum, id;
    bool have;
};

// Uniforms inside [kSafeULo, kSafeUHi] give Gumbel noise in [-2.82, 16.64]; then an edge with p = 0 (score
// log(1e-12) + g <= -10.98) can never beat an edge with p >= kSafeAttr (score >= log(1e-3) - 2.82 = -9.73), and any
// such score beats the lowest() start value: the arg-max of src/direction_mpnn.py:136-139 over ALL in-edges equals the
// arg-max over the ELIGIBLE ones (strict '>' in ascending edge id either way), so ineligible edges need no logf and a
// lone eligible edge needs no noise at all. Outside these bounds the literal scan over every in-edge runs.
// The in-kernel Philox stream only produces uniforms inside the interval.
constexpr float kSafeULo = 5.9604645e-08f, kSafeUHi = 0.99999994f, kSafeAttr = 1e-3f;
constexpr float kClearGap = 1e-4f;     // see the uniform-weight race in k_ell_select_append

__device__ __forceinline__ float philox_uniform(const Noise& nz, int L, int j, float (&un)[4], int& have_group) {
    if (have_group != (j >> 2)) {
        uint32_t klo, khi;
        philox_key(nz, klo, khi);
        philox4x32_10((uint32_t)L, 0u, nz.step_id, (uint32_t)(j >> 2), klo, khi, un);
        have_group = j >> 2;
    }
    const int jj = j & 3;
    return jj == 0 ? un[0] : (jj == 1 ? un[1] : (jj == 2 ? un[2] : un[3]));
}

// The whole in-edge scan of link d out of the CSR, in ascending original edge id: probability sum, then — only where it
// is positive (src/direction_mpnn.py:142-144) — the Gumbel arg-max with strict '>' (lowest edge id wins ties).
template <bool kExtNoise>
__device__ __noinline__ Pick scan_in_edges_csr(const tarl_dual_csr& g, const Store& s, const float* __restrict__ attr_in,
                                               const Noise& nz, int r, int d, int L, float t, bool free_d, float room_d,
                                               float ridx_d) {
    const int base = L - d;
    const int k0 = g.in_ptr[d], k1 = g.in_ptr[d + 1];
    Pick out = {0.0f, 0.0f, false};
    int n_elig = 0, lone = -1;
    bool safe = true;
    for (int k = k0; k < k1; ++k) {
        const int Lu = base + g.in_src[k];
        const float a = attr_in[k];
        const float p = edge_prob(s.hot_cur[2 * Lu], s.sel[Lu], t, free_d, room_d, ridx_d, a);
        out.psum += p;
        if (p > 0.0f) { ++n_elig; lone = k; safe = safe && (a >= kSafeAttr); }
        if (kExtNoise) {
            const float uu = nz.ext[(int64_t)r * g.n_edges + g.in_eid[k]];
            safe = safe && (uu >= kSafeULo) && (uu <= kSafeUHi);
        }
    }
    if (!(out.psum > 0.0f)) return out;
    if (safe && n_elig == 1) {
        const int Lu = base + g.in_src[lone];
        out.id = s.hot_cur[2 * Lu].x; out.have = true;
        return out;
    }
    float best = -FLT_MAX;
    float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    int grp = -1;
    for (int k = k0; k < k1; ++k) {
        const int Lu = base + g.in_src[k];
        const float4 U = s.hot_cur[2 * Lu];
        const float p = edge_prob(U, s.sel[Lu], t, free_d, room_d, ridx_d, attr_in[k]);
        if (safe && !(p > 0.0f)) continue;
        const float uu = kExtNoise ? nz.ext[(int64_t)r * g.n_edges + g.in_eid[k]] : philox_uniform(nz, L, k - k0, un, grp);
        const float sc = gumbel_score(p, uu);
        if (sc > best) { best = sc; out.id = U.x; out.have = true; }
    }
    return out;
}

// Tail append on the link's own record (src/direction_mpnn.py:171-195, on EVERY link), the {NUM, tail id} summary the
// response phase gathers, delta_travel_time of the link's head (:94-96, from the PRE-step state: ONE value per upstream
// link — every out-edge of the link carries the same number, tarl_expand_delta_tt materialises the [E] form), and the
// (Until ABI 22 the link also set a "pop hint" byte on the upstream link whose head it admitted, so that the response
// phase could request that link's ring slots one load level earlier: the scattered byte store cost the direction
// kernel 1.9 us per step of one million links and returned 0.3 us to the response kernel — removed.)
template <bool kWide>
__device__ __forceinline__ void append_and_publish(const Store& s, int L, float4 hA, float4 hB, const float4 st,
                                                   const Pick pk, float t, float* __restrict__ dtt_link,
                                                   int32_t* __restrict__ flags) {
    const float num = hA.z, maxn = hA.w, fftt = st.x;
    int meta = __float_as_int(hB.w);
    const bool bad = !(num >= 0.0f) || !(num < (float)s.Nmax);
    float chosen = 0.0f;
    if (pk.psum > 0.0f) {
        if (!pk.have) atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_NO_WINNER);
        else chosen = pk.id;
    }
    if (dtt_link != nullptr) dtt_link[L] = max_propagate_nan((hA.y - hB.x) - fftt, 0.0f);
    float num_post = num, tail_post = hB.y;
    if (bad) {
        atomicOr(&flags[TARL_FLAG_ERROR], TARL_ERR_QUEUE_RANGE);
    } else {
        const int q = (int)num;
        const float dep_new = t + max_propagate_nan(fftt, st.y / ((maxn + 10.0f) - num));
        if (q == 0) {                       // the tail slot IS the head slot
            hA.x = chosen; hB.x = t; hA.y = dep_new;
            tail_post = chosen;
            meta &= ~kMetaGarbage;
            if (chosen != 0.0f) { num_post = num + 1.0f; hB.y = chosen; }
        } else if (chosen != 0.0f) {        // a real admission: one ring slot write
            s.queue[(size_t)L * s.M + ring_pos(meta & kMetaRingMask, q, s.M)] = make_float4(chosen, t, dep_new, 0.0f);
            num_post = num + 1.0f;
            hB.y = chosen;
            tail_post = chosen;
            meta &= ~kMetaGarbage;
        } else {                            // the reference writes (0, t, t+tt) past the tail: keep it pending
            meta |= kMetaGarbage;
            hB.z = dep_new;
        }
        hA.z = num_post;
    }
    hB.w = __int_as_float(meta);
    const uint64_t keep = l2_policy(s.pol_state);
    if (kWide) st_record(&s.hot_next[2 * (size_t)L], hA, hB, s.pol_state == kPolKeep);
    else st_record_narrow(&s.hot_next[2 * (size_t)L], hA, hB, keep);     // inside a __noinline__ function
#ifndef TARL_ABLATE_POST        // tuning only (wrong results)
    st_hint(&s.post[L], make_float2(num_post, tail_post), keep);
#endif
}

// The general form of the direction phase for one link: in-edge scan out of the CSR (any degree, any edge weight, any
// noise), arg-max, append. The CSR kernels run it on every link; the ELL kernels on the links their columns cannot describe.
template <bool kExtNoise>
__device__ __noinline__ void select_append_general(const tarl_dual_csr& g, const Store& s,
                                                   const float* __restrict__ attr_in, const Noise& nz, float t,
                                                   float* __restrict__ dtt_link, int32_t* __restrict__ flags, int r,
                                                   int d, int L) {
    const float4 hA = s.hot_cur[2 * (size_t)L], hB = s.hot_cur[2 * (size_t)L + 1];
    const float4 st = s.stat_a[d];
    const Pick pk = scan_in_edges_csr<kExtNoise>(g, s, attr_in, nz, r, d, L, t, hA.z < (hA.w - 3.0f), hA.w - hA.z, st.z);
    append_and_publish<false>(s, L, hA, hB, st, pk, t, dtt_link, flags);
}

template <bool kExtNoise>
__global__ void __launch_bounds__(kThreads) k_csr_select_append(const __grid_constant__ tarl_dual_csr g,
                                                                const __grid_constant__ Store s,
                                                                const float* __restrict__ attr_in,
                                                                const __grid_constant__ Noise nz, float t,
                                                                float* __restrict__ dtt_link,
                                                                int32_t* __restrict__ flags) {
    pdl_trigger();
    pdl_wait();
    const int d = blockIdx.x * kThreads + threadIdx.x;
    if (d >= s.N) return;
    const int r = blockIdx.y;
    select_append_general<kExtNoise>(g, s, attr_in, nz, t, dtt_link, flags, r, d, r * s.N + d);
}

// The streaming form: one thread per link, the first W in-edges out of the ELLPACK columns.
//   level 1 (addressed by the link id alone): statics, W upstream ids and edge weights before the dependency wait, own
//            record after it;
//   level 2: W gathers of upstream A-halves + SELECTED_ROAD, all in flight together.
// Eligibility p_e (src/direction_mpnn.py:81-91), probability sum in ascending edge id, and — only where the sum is
// positive — the pick. With uniforms in [2^-24, 1-2^-24] and eligible weights >= 1e-3 an ineligible edge can never beat
// an eligible one (see kSafe* above): a lone eligible edge wins without noise or logf, several eligible edges draw Philox
// uniforms and compare Gumbel scores among themselves. Links whose in-edges do not fit the W columns, or carry a weight
// outside the safe bounds, have -2 in column W-1 (topology.py) and walk their CSR segment (the literal scan).
// (Measured and rejected in round 2: listing the contested links — 10 % of the links at bench load, present in 97 % of the
// warps — for a second, dense kernel. The streaming kernel loses a third of its instructions and 16 registers but not a
// microsecond — it is bound by its two dependent load levels, not by issue slots or occupancy — and the second kernel
// adds a third dependent launch to the step: 61 -> 68 us per step.)
// The tile that is launched `ahead` CTAs after this one (x fastest, then the replica) and how many links it holds.
struct Ahead { int r, d0, count; };
__device__ __forceinline__ Ahead tile_ahead(int ahead, int N) {
    const long long tile = (long long)blockIdx.y * gridDim.x + blockIdx.x + ahead;
    Ahead a = {0, 0, 0};
    if (tile < (long long)gridDim.x * gridDim.y) {
        a.r = (int)(tile / gridDim.x);
        a.d0 = (int)(tile - (long long)a.r * gridDim.x) * kThreads;
        a.count = min(kThreads, N - a.d0);
    }
    return a;
}

#ifndef TARL_SELECT_MINBLOCKS
#define TARL_SELECT_MINBLOCKS 9      // resident CTAs per SM the W = 4 kernel is compiled for (56 registers, no spills)
#endif
template <int W, bool kExtNoise>
__global__ void __launch_bounds__(kThreads, W == 4 ? TARL_SELECT_MINBLOCKS : 1) k_ell_select_append(const __grid_constant__ tarl_dual_csr g,
                                                                tarl_dual_ell ell, const __grid_constant__ Store s,
                                                                const float* __restrict__ attr_in,
                                                                const __grid_constant__ Noise nz, float t,
                                                                float* __restrict__ dtt_link,
                                                                int32_t* __restrict__ flags, int ahead) {
    const int d = blockIdx.x * kThreads + threadIdx.x;
    if (d >= s.N) return;
    const int r = blockIdx.y;
    const int base = r * s.N;
    const int L = base + d;
    const uint64_t keep = l2_policy(s.pol_state), strm = l2_policy(s.pol_static);
    // level 0: what the tile `ahead` launches further on will load at its level 1, requested from the L2 now
    if (ahead > 0 && threadIdx.x < 3 + 2 * W) {
        const Ahead tl = tile_ahead(ahead, s.N);
        if (tl.count > 0) {
            const int i = (int)threadIdx.x - 3;
            if (i == -3) l2_prefetch_span(s.hot_cur, ((size_t)tl.r * s.N + tl.d0) * 32, tl.count * 32);
            else if (i == -2) l2_prefetch_span(s.stat_a, (size_t)tl.d0 * 16, tl.count * 16);
            else if (i == -1) l2_prefetch_span(s.sel, ((size_t)tl.r * s.N + tl.d0) * 4, tl.count * 4);
            else if (i < W || !s.uni_hint)       // (the edge-weight columns are read by a few links only under the hint)
                l2_prefetch_span(i < W ? (const void*)ell.in_src : (const void*)ell.in_attr,
                                 ((size_t)(i < W ? i : i - W) * ell.pitch + tl.d0) * 4, tl.count * 4);
        }
    }
    // level 1: everything addressed by the link id alone — the static part before the dependency wait
    pdl_trigger();
    const float4 st = ld_static(&s.stat_a[d], strm);
    int u[W];
    float a[W];
    // Edge weights. On the networks this path is measured on (edge_attr = 1 / out-degree on grids and ring-radials) the
    // in-edges of nearly every link carry ONE weight: with the store's hint (stat_a.w, NaN where they differ) that is the
    // float already loaded above, and the W-column read (16 of the kernel's ~100 bytes per link) is left to the few
    // links whose weights differ — for them one load level later. Without the hint the columns are level-1 loads.
#pragma unroll
    for (int j = 0; j < W; ++j) {
        u[j] = ld_static(&ell.in_src[(size_t)j * ell.pitch + d], strm);
        a[j] = st.w;
        if (!s.uni_hint) a[j] = ld_static(&ell.in_attr[(size_t)j * ell.pitch + d], strm);
    }
    pdl_wait();
    if (u[W - 1] == -2) {     // more than W in-edges or an unsafe weight: this link walks its CSR segment instead
        select_append_general<kExtNoise>(g, s, attr_in, nz, t, dtt_link, flags, r, d, L);
        return;
    }
    if (s.uni_hint && st.w != st.w) {
#pragma unroll
        for (int j = 0; j < W; ++j) a[j] = ld_static(&ell.in_attr[(size_t)j * ell.pitch + d], strm);
    }
    float4 hA, hB;
    ld_record(&s.hot_cur[2 * (size_t)L], hA, hB, s.pol_state == kPolKeep);
    const bool free_d = hA.z < (hA.w - 3.0f);
    const float room_d = hA.w - hA.z, ridx_d = st.z;
    Pick pk = {0.0f, 0.0f, false};
    // level 2: the neighbours' records, all gathers in flight together
    float4 U[W];
    float S[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (u[j] >= 0) {
#ifdef TARL_ABLATE_GATHER       // tuning only (wrong results): what would the kernel cost without its second load level?
            U[j] = hA; S[j] = st.z + (float)j;
#else
            U[j] = ld_hint(&s.hot_cur[2 * (size_t)(base + u[j])], keep);
            S[j] = s.sel[base + u[j]];
#endif
        }
    }
    float p[W];
    int n_elig = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        p[j] = 0.0f;
        if (u[j] >= 0) {
            p[j] = edge_prob(U[j], S[j], t, free_d, room_d, ridx_d, a[j]);
            pk.psum += p[j];
            if (p[j] > 0.0f) { ++n_elig; pk.id = U[j].x; }
        }
    }
    if (pk.psum > 0.0f) {
        bool safe = true;
        float uu[W];
        if (kExtNoise) {          // injected uniforms may lie outside the safe interval
            const int k0 = g.in_ptr[d];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                uu[j] = 0.5f;
                if (u[j] >= 0) {
                    uu[j] = nz.ext[(int64_t)r * g.n_edges + g.in_eid[k0 + j]];
                    safe = safe && (uu[j] >= kSafeULo) && (uu[j] <= kSafeUHi);
                }
            }
        }
#ifdef TARL_ABLATE_CONTEST      // tuning only (wrong results): what does the contested path cost?
        if (true) {
#else
        if (safe && n_elig == 1) {
#endif
            pk.have = true;       // a lone eligible edge wins without noise
        } else {
            float best = -FLT_MAX;
            float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
            uint32_t klo = 0u, khi = 0u;
            if (!kExtNoise) philox_key(nz, klo, khi);
            // One weight w on every in-edge (the store's hint) and only eligible edges in the race: the scores are
            // log(w + 1e-12) + g(u_j) with g(u) = -log(-log u) increasing, so the winner is the edge with the LARGEST
            // UNIFORM — no logarithm needed, where the literal form spends three per edge and nearly every warp has a
            // contested lane (the kernel had become issue-bound: 70 % issue utilisation, ~700 instructions per warp,
            // of which ~300 were these logarithms). fp32 rounding can reorder scores only when they are closer than
            // their evaluation error, < 5e-6 (1 ulp of each logf and of the sum, |score| < 32); g'(u) = 1 / (u (-ln u))
            // >= e everywhere, so uniforms kClearGap = 1e-4 apart give scores >= 2.7e-4 apart: the order of the
            // uniforms IS the order of the reference's fp32 scores. Closer than that (~1e-3 of the contested links, exact
            // ties of the uniforms included) the literal scores below decide, strict '>' in ascending edge id.
            bool decided = false;
            if (safe && s.uni_hint && st.w == st.w) {
                float m1 = -1.0f, m2 = -1.0f, id1 = 0.0f;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    if (!kExtNoise && (j & 3) == 0)
                        philox4x32_10((uint32_t)L, 0u, nz.step_id, (uint32_t)(j >> 2), klo, khi, un);
                    if (u[j] >= 0 && p[j] > 0.0f) {
                        const float uj = kExtNoise ? uu[j] : un[j & 3];
                        if (uj > m1) { m2 = m1; m1 = uj; id1 = U[j].x; }
                        else if (uj > m2) m2 = uj;
                    }
                }
                if (m1 - m2 >= kClearGap) { pk.id = id1; pk.have = true; decided = true; }
            }
            if (!decided) {
#pragma unroll
            for (int j = 0; j < W; ++j) {
#ifdef TARL_ABLATE_PHILOX       // tuning only (wrong results): what does the Philox draw cost?
                if (!kExtNoise && (j & 3) == 0) { un[0] = 0.3f + 1e-9f * (float)L; un[1] = 0.5f; un[2] = 0.7f; un[3] = 0.9f; }
#else
                if (!kExtNoise && (j & 3) == 0)
                    philox4x32_10((uint32_t)L, 0u, nz.step_id, (uint32_t)(j >> 2), klo, khi, un);
#endif
                if (u[j] >= 0 && !(safe && !(p[j] > 0.0f))) {
                    const float sc = gumbel_score(p[j], kExtNoise ? uu[j] : un[j & 3]);
                    if (sc > best) { best = sc; pk.id = U[j].x; pk.have = true; }
                }
            }
            }
        }
    }
    append_and_publish<true>(s, L, hA, hB, st, pk, t, dtt_link, flags);
}

// ------------------------------------------------------------------------------------------------ response phase
// The pop itself (src/response_mpnn.py:119-122) as a ring-head increment: new head <- logical slot 1, and the slot that
// becomes logical Nmax-1 <- old logical Nmax-1 (the reference's shift leaves the last slot in place, i.e. duplicates it).
__device__ __forceinline__ float pop_head(const Store& s, int L, float4 hA, float4 hB, float t) {
    int meta = __float_as_int(hB.w);
    const int rh = meta & kMetaRingMask;
    const int M = s.M;
    const int q = (int)hA.z;                                    // >= 1 here
    const bool gv = meta & kMetaGarbage;
    const float4 garbage = make_float4(0.0f, t, hB.z, 0.0f);    // pending garbage was (re)written this very step
    float4* Q = s.queue + (size_t)L * M;
    const float4 q_head = Q[rh];
    float4 q_last = q_head;
    if (M > 1) q_last = Q[ring_pos(rh, M, M)];
    const float4 new_head = (gv && q == 1) ? garbage : q_head;
    if (M > 1) {
        Q[rh] = (gv && q == M) ? garbage : q_last;              // becomes logical slot M after the increment
    } else if (gv && q == 1) {
        Q[rh] = garbage;                                        // Nmax == 2: slot 1 keeps (a copy of) its value
    }
    int nrh = rh + 1; if (nrh >= M) nrh = 0;
    meta = (meta & ~kMetaRingMask) | nrh;
    if (gv && q == 1) meta &= ~kMetaGarbage;                    // the garbage became the head slot
    st_record(&s.hot_next[2 * (size_t)L], make_float4(new_head.x, new_head.z, hA.z - 1.0f, hA.w),
              make_float4(new_head.y, hB.y, hB.z, __int_as_float(meta)), s.pol_state == kPolKeep);
    return new_head.z;                                          // exit time of the new head
}

// src/response_mpnn.py:66-83: NUM_up > 0, NUM_dn > 0, tail(dn) == head(up). A = own post-append record, D = {NUM, tail}
__device__ __forceinline__ bool accepts(const float4 A, const float2 D) {
    return at_least_one(A.z) && at_least_one(D.x) && same_id(D.y, A.x);
}

__device__ __noinline__ bool scan_out_edges_csr(const tarl_dual_csr& g, const Store& s, int base, int u, const float4 A) {
    bool accept = false;
    const int k1 = g.out_ptr[u + 1];
    for (int k = g.out_ptr[u]; k < k1; ++k) accept = accept || accepts(A, s.post[base + g.out_dst[k]]);
    return accept;
}

// bit (u % 32) of word r*ceil(N/32) + u/32 = pop[r, u]: one ballot and one store per warp (kThreads % 32 == 0, so a
// warp never straddles two words)
__device__ __forceinline__ void publish_pop_bits(uint32_t* __restrict__ pop_bits, bool accept, int r, int u, int N) {
    const unsigned m = __ballot_sync(0xffffffffu, accept);
    if (pop_bits != nullptr && (threadIdx.x & 31) == 0 && u < N) pop_bits[r * ((N + 31) >> 5) + (u >> 5)] = m;
}

__global__ void __launch_bounds__(kThreads) k_csr_respond_pop(tarl_dual_csr g, Store s, float t,
                                                              uint8_t* __restrict__ pop, uint32_t* __restrict__ pop_bits,
                                                              int32_t* __restrict__ flags) {
    pdl_trigger();
    pdl_wait();
    const int u = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    const int base = r * s.N;
    const int L = base + u;
    bool accept = false;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    if (u < s.N) {
        A = s.hot_next[2 * (size_t)L]; B = s.hot_next[2 * (size_t)L + 1];
        accept = scan_out_edges_csr(g, s, base, u, A);
        pop[L] = accept ? 1 : 0;
    }
    publish_pop_bits(pop_bits, accept, r, u, s.N);
    if (__syncthreads_or(accept) && threadIdx.x == 0) flags[TARL_FLAG_ANY_POP] = 1;
    if (accept) pop_head(s, L, A, B, t);
}

#ifndef TARL_RESPOND_MINBLOCKS
#define TARL_RESPOND_MINBLOCKS 12     // 40 registers: at 16 CTAs (32 registers) the policy descriptors spill
#endif
template <int W>
__global__ void __launch_bounds__(kThreads, W == 4 ? TARL_RESPOND_MINBLOCKS : 1) k_ell_respond_pop(tarl_dual_csr g, tarl_dual_ell ell, Store s, float t,
                                                              uint8_t* __restrict__ pop, uint32_t* __restrict__ pop_bits,
                                                              int32_t* __restrict__ flags, int ahead) {
    const int u = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    const int base = r * s.N;
    const int L = base + u;
    bool accept = false;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    const uint64_t keep = l2_policy(s.pol_state), strm = l2_policy(s.pol_static);
    if (ahead > 0 && threadIdx.x < 2 + W) {     // level 0: the level-1 data of the tile `ahead` launches further on
        const Ahead tl = tile_ahead(ahead, s.N);
        if (tl.count > 0) {
            const int i = (int)threadIdx.x - 2;
            const size_t L0 = (size_t)tl.r * s.N + tl.d0;
            if (i == -2) l2_prefetch_span(s.hot_next, L0 * 32, tl.count * 32);
            else if (i == -1) l2_prefetch_span(s.post, L0 * 8, tl.count * 8);
            else l2_prefetch_span(ell.out_dst, ((size_t)i * ell.pitch + tl.d0) * 4, tl.count * 4);
        }
    }
    pdl_trigger();
    int dn[W];
#pragma unroll
    for (int j = 0; j < W; ++j) dn[j] = (u < s.N) ? ld_static(&ell.out_dst[(size_t)j * ell.pitch + u], strm) : -1;   // static: before the wait
    pdl_wait();
    if (u < s.N) {
        // level 1: own post-append record and whether a downstream link admitted this link's head
        ld_record(&s.hot_next[2 * (size_t)L], A, B, s.pol_state == kPolKeep);
        // level 2: the neighbours' summaries (the ring slots of a pop are a third level, on the ~25 % of links that pop)
        if (dn[W - 1] == -2) {
            accept = scan_out_edges_csr(g, s, base, u, A);
        } else {
            float2 D[W];
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (dn[j] >= 0) D[j] = ld_hint(&s.post[base + dn[j]], keep);
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (dn[j] >= 0) accept = accept || accepts(A, D[j]);
        }
        pop[L] = accept ? 1 : 0;
    }
    if (__syncthreads_or(accept) && threadIdx.x == 0) flags[TARL_FLAG_ANY_POP] = 1;
    if (accept) pop_head(s, L, A, B, t);
    publish_pop_bits(pop_bits, accept, r, u, s.N);      // last: nothing else is live any more
}

// ---------------------------------------------------------------- response phase + withdrawal + occupancy observation
// One environment step of the RL loop is core step -> withdrawal -> insertion -> observation / reward
// (src/reinforcement_learning.py:237-266). The withdrawal of a link touches that link's queue and its agents' rows only,
// and the response phase ends with the link's final record of the core step in registers — so the thread that popped
// (or not) also withdraws: only links whose head is due (NUM > 0, exit time <= t: a few per cent) take the rare path
// (withdraw_one of population.cuh on the records just written), every thread leaves NUM in the trajectory frame and the
// warps add it into the replica's occupancy. A separate withdrawal pass would read every record once more (1.3 GB per
// step of 1024 grid100 replicas). Used for rollouts whose nets read the occupancy only; links in link-id order.
// (Compacting the due links into a per-replica list served by a second kernel was measured: the response kernel drops
// from 56 to 40 registers and 588 -> 524 us, but a grid that covers the worst case costs 166 us to find its few items.)
struct WithdrawArgs {
    AgentTable at;
    tarl_csr adj;
    uint8_t* mask;        // [R*N] withdrawn mask (the entry of withdraw_history)
    int32_t* counters;    // [R*2] running totals {inserted, withdrawn} or nullptr
    float* num_out;       // [R, n_nodes]
    int32_t* occupancy;   // [R]
    float* src_sel;
    int n_nodes;
};

__device__ __noinline__ float withdraw_due_link(const Store& s, const WithdrawArgs& wa, float t, int32_t* flags, int r, int u) {
    const StoreAcc acc = {s, reinterpret_cast<float*>(s.hot_next), wa.src_sel, wa.n_nodes, t, s.N, s.Nmax};
    return withdraw_one(acc, wa.at, wa.adj, t, wa.mask, wa.counters, flags, r, u);
}

template <int W>
__global__ void __launch_bounds__(kThreads) k_ell_respond_pop_withdraw(tarl_dual_csr g, tarl_dual_ell ell,
                                                                       const __grid_constant__ Store s, float t,
                                                                       uint8_t* __restrict__ pop,
                                                                       uint32_t* __restrict__ pop_bits,
                                                                       int32_t* __restrict__ flags,
                                                                       const __grid_constant__ WithdrawArgs wa, int ahead) {
    // (__grid_constant__: the rare path takes `s` and `wa` by reference; without it every thread would first copy both
    // structs from the parameter space into local memory — 200 bytes per thread, measured 1.2 ms instead of 0.4 ms)
    const int u = blockIdx.x * kThreads + threadIdx.x;
    const int r = blockIdx.y;
    const int base = r * s.N;
    const int L = base + u;
    bool accept = false;
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    const uint64_t keep = l2_policy(s.pol_state), strm = l2_policy(s.pol_static);
    if (ahead > 0 && threadIdx.x < 2 + W) {     // level 0: the level-1 data of the tile `ahead` launches further on
        const Ahead tl = tile_ahead(ahead, s.N);
        if (tl.count > 0) {
            const int i = (int)threadIdx.x - 2;
            const size_t L0 = (size_t)tl.r * s.N + tl.d0;
            if (i == -2) l2_prefetch_span(s.hot_next, L0 * 32, tl.count * 32);
            else if (i == -1) l2_prefetch_span(s.post, L0 * 8, tl.count * 8);
            else l2_prefetch_span(ell.out_dst, ((size_t)i * ell.pitch + tl.d0) * 4, tl.count * 4);
        }
    }
    pdl_trigger();
    int dn[W];
#pragma unroll
    for (int j = 0; j < W; ++j) dn[j] = (u < s.N) ? ld_static(&ell.out_dst[(size_t)j * ell.pitch + u], strm) : -1;   // static: before the wait
    pdl_wait();
    if (u < s.N) {
        ld_record(&s.hot_next[2 * (size_t)L], A, B, s.pol_state == kPolKeep);
        if (dn[W - 1] == -2) {
            accept = scan_out_edges_csr(g, s, base, u, A);
        } else {
            float2 D[W];
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (dn[j] >= 0) D[j] = ld_hint(&s.post[base + dn[j]], keep);
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (dn[j] >= 0) accept = accept || accepts(A, D[j]);
        }
        pop[L] = accept ? 1 : 0;
    }
    publish_pop_bits(pop_bits, accept, r, u, s.N);
    if (__syncthreads_or(accept) && threadIdx.x == 0) flags[TARL_FLAG_ANY_POP] = 1;
    float head_exit = A.y, num = A.z;
    if (accept) { head_exit = pop_head(s, L, A, B, t); num = A.z - 1.0f; }
    if (u < s.N) {
        if (0.0f < num && head_exit <= t) num = withdraw_due_link(s, wa, t, flags, r, u);     // slot 0 is due: :362-366
        else wa.mask[L] = 0;
        wa.num_out[(size_t)r * wa.n_nodes + u] = num;
    } else {
        num = 0.0f;
    }
    const int extra = wa.n_nodes - s.N;                           // the non-road nodes of the frame hold no agents
    if (u < extra) wa.num_out[(size_t)r * wa.n_nodes + s.N + u] = 0.0f;
    int num_i = (int)num;
    for (int off = 16; off > 0; off >>= 1) num_i += __shfl_xor_sync(0xffffffffu, num_i, off);
#ifndef TARL_ABLATE_OCC      // tuning only: what do the per-warp atomics on one word per replica cost?
    if ((threadIdx.x & 31) == 0 && num_i != 0) atomicAdd(&wa.occupancy[r], num_i);
#endif
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in engine_common.cuh).
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, cudaStream_t cs, Args... args) {
    static const bool off = getenv("TARL_NO_PDL") != nullptr;      // tuning / A-B measurements only
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = cs;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// How many CTA launches ahead a tile's level-1 data is requested from the L2 (0 = off). CTAs are handed out in launch
// order, so the distance is counted in resident CTAs: a little more than one full wave of the kernel (148 SMs x 9 CTAs
// for the direction phase, x 16 for the response phase). TARL_AHEAD_SELECT / TARL_AHEAD_RESPOND override (tuning).
inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v != nullptr ? atoi(v) : dflt;
}
inline int ahead_select() { static const int v = env_int("TARL_AHEAD_SELECT", 700); return v; }
inline int ahead_respond() { static const int v = env_int("TARL_AHEAD_RESPOND", 1200); return v; }

inline int blocks_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

}  // namespace

namespace tarl {

int make_store(const tarl_link_store* p, Store* s) {
    if (p == nullptr || p->n_links < 0 || p->n_replicas < 1 || p->nmax < 2 || p->nmax - 1 > kMetaRingMask) return TARL_E_BADARG;
    if (p->n_replicas > 65535 || (int64_t)p->n_links * p->n_replicas * 2 >= INT32_MAX) return TARL_E_BADARG;
    if (p->n_links > 0 && (!p->hot_cur || !p->hot_next || !p->sel || !p->stat_a || !p->stat_b || !p->queue || !p->post))
        return TARL_E_BADARG;
    s->N = p->n_links; s->R = p->n_replicas; s->Nmax = p->nmax; s->M = p->nmax - 1;
    s->hot_cur = static_cast<const float4*>(p->hot_cur);
    s->hot_next = static_cast<float4*>(p->hot_next);
    s->sel = static_cast<float*>(p->sel);
    s->stat_a = static_cast<const float4*>(p->stat_a);
    s->stat_b = static_cast<const float4*>(p->stat_b);
    s->queue = static_cast<float4*>(p->queue);
    s->post = static_cast<float2*>(p->post);
    s->uni_hint = (p->hints & TARL_STORE_UNIFORM_WEIGHTS) != 0;
    s->slot_link = p->slot_link;
    s->link_slot = p->link_slot;
    if ((p->slot_link == nullptr) != (p->link_slot == nullptr)) return TARL_E_BADARG;
    // records move as one 256-bit access
    if (((reinterpret_cast<uintptr_t>(p->hot_cur) | reinterpret_cast<uintptr_t>(p->hot_next)) & 31) != 0) return TARL_E_BADARG;
    // L2 residency (engine_common.cuh): does the state two consecutive kernels share fit the 126 MB L2?
    const bool fits = (int64_t)p->n_links * p->n_replicas * 80 <= (int64_t)96 << 20;
    s->pol_state = fits ? kPolKeep : kPolDefault;
    s->pol_static = fits ? (p->n_replicas == 1 ? kPolStream : kPolDefault) : kPolKeep;
    static const char* const tune = getenv("TARL_L2_POLICY");        // tuning only: "sk" = state, statics as digits 0/1/2
    if (tune != nullptr && tune[0] >= '0' && tune[0] <= '2' && tune[1] >= '0' && tune[1] <= '2') {
        s->pol_state = tune[0] - '0'; s->pol_static = tune[1] - '0';
    }
    return TARL_OK;
}

}  // namespace tarl

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

namespace {

int check_step(const tarl_dual_csr* g, const tarl_dual_ell* ell, const Store& s, const float* attr_in,
               const tarl_step_io* io) {
    if (g == nullptr || io == nullptr || io->flags == nullptr || g->n_links != s.N) return TARL_E_BADARG;
    if (s.N == 0) return TARL_OK;
    if (io->pop == nullptr || g->in_ptr == nullptr || g->out_ptr == nullptr) return TARL_E_BADARG;
    if (g->n_edges > 0 && (attr_in == nullptr || g->in_src == nullptr || g->out_dst == nullptr)) return TARL_E_BADARG;
    if (io->noise != nullptr && g->n_edges > 0 && g->in_eid == nullptr) return TARL_E_BADARG;
    if (ell != nullptr) {
        if ((ell->width != 4 && ell->width != 8) || ell->pitch < s.N) return TARL_E_BADARG;
        if (!ell->in_src || !ell->in_attr || !ell->out_dst) return TARL_E_BADARG;
    }
    return TARL_OK;
}

void launch_step(const tarl_dual_csr* g, const tarl_dual_ell* ell, const Store& s, const float* attr_in,
                 const tarl_step_io& io, cudaStream_t cs, uint32_t phase_mask) {
    const dim3 grid(blocks_for(s.N), s.R);
    const Noise nz = {io.noise, (uint32_t)io.seed, (uint32_t)(io.seed >> 32), io.step_id,
                      reinterpret_cast<const unsigned long long*>(io.seed_dev)};
    const bool ext = nz.ext != nullptr;
    const float t = io.t;
    float* dtt = io.delta_tt_link;
    int32_t* flags = io.flags;
    const int pf = s.pol_state == kPolKeep ? 1 : 0;       // prefetch ahead only where the state lives in the L2
    if (phase_mask & TARL_PHASE_SELECT_APPEND) {
        if (ell == nullptr) {
            if (ext) launch_pdl(k_csr_select_append<true>, grid, cs, *g, s, attr_in, nz, t, dtt, flags);
            else launch_pdl(k_csr_select_append<false>, grid, cs, *g, s, attr_in, nz, t, dtt, flags);
        } else if (ell->width == 4) {
            if (ext) launch_pdl(k_ell_select_append<4, true>, grid, cs, *g, *ell, s, attr_in, nz, t, dtt, flags, pf * ahead_select());
            else launch_pdl(k_ell_select_append<4, false>, grid, cs, *g, *ell, s, attr_in, nz, t, dtt, flags, pf * ahead_select());
        } else {
            if (ext) launch_pdl(k_ell_select_append<8, true>, grid, cs, *g, *ell, s, attr_in, nz, t, dtt, flags, pf * ahead_select());
            else launch_pdl(k_ell_select_append<8, false>, grid, cs, *g, *ell, s, attr_in, nz, t, dtt, flags, pf * ahead_select());
        }
    }
    if (phase_mask & TARL_PHASE_RESPOND_SHIFT) {
        if (ell == nullptr) launch_pdl(k_csr_respond_pop, grid, cs, *g, s, t, io.pop, io.pop_bits, flags);
        else if (ell->width == 4) launch_pdl(k_ell_respond_pop<4>, grid, cs, *g, *ell, s, t, io.pop, io.pop_bits, flags, pf * ahead_respond());
        else launch_pdl(k_ell_respond_pop<8>, grid, cs, *g, *ell, s, t, io.pop, io.pop_bits, flags, pf * ahead_respond());
    }
}

// ---------------------------------------------------------------------------------------------- noise / delta_tt forms
__global__ void __launch_bounds__(kThreads) k_store_noise(tarl_dual_csr g, int R, uint32_t seed_lo, uint32_t seed_hi,
                                                          uint32_t step_id, float* __restrict__ out) {
    const int d = blockIdx.x * kThreads + threadIdx.x;
    if (d >= g.n_links) return;
    const int r = blockIdx.y;
    const int L = r * g.n_links + d;
    const Noise nz = {nullptr, seed_lo, seed_hi, step_id, nullptr};
    float un[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    int grp = -1;
    const int k0 = g.in_ptr[d], k1 = g.in_ptr[d + 1];
    for (int k = k0; k < k1; ++k)
        out[(int64_t)r * g.n_edges + g.in_eid[k]] = philox_uniform(nz, L, k - k0, un, grp);
}

__global__ void __launch_bounds__(256) k_expand_delta_tt(const int32_t* __restrict__ edge_src, int E, int N,
                                                         const float* __restrict__ dtt_link, float* __restrict__ out) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= E) return;
    const int r = blockIdx.y;
    out[(int64_t)r * E + e] = dtt_link[(int64_t)r * N + edge_src[e]];
}

}  // namespace

int tarl_store_step(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                    const float* attr_in, const tarl_step_io* io, void* stream, uint32_t phase_mask) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if ((rc = check_step(g, ell, s, attr_in, io)) != TARL_OK) return rc;
    if (s.N == 0) return TARL_OK;
    launch_step(g, ell, s, attr_in, *io, static_cast<cudaStream_t>(stream), phase_mask);
    return launch_status();
}

int tarl_store_noise(const tarl_dual_csr* g, int32_t n_replicas, uint64_t seed, uint32_t step_id, float* noise,
                     void* stream) {
    if (g == nullptr || g->n_links < 0 || n_replicas < 1 || n_replicas > 65535) return TARL_E_BADARG;
    if (g->n_links == 0 || g->n_edges == 0) return TARL_OK;
    if (noise == nullptr || g->in_ptr == nullptr || g->in_eid == nullptr) return TARL_E_BADARG;
    k_store_noise<<<dim3(blocks_for(g->n_links), n_replicas), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        *g, n_replicas, (uint32_t)seed, (uint32_t)(seed >> 32), step_id, noise);
    return launch_status();
}

int tarl_expand_delta_tt(const int32_t* edge_src, int32_t n_edges, int32_t n_links, int32_t n_replicas,
                         const float* delta_tt_link, float* delta_tt, void* stream) {
    if (n_edges < 0 || n_links < 0 || n_replicas < 1 || n_replicas > 65535) return TARL_E_BADARG;
    if (n_edges == 0) return TARL_OK;
    if (edge_src == nullptr || delta_tt_link == nullptr || delta_tt == nullptr) return TARL_E_BADARG;
    k_expand_delta_tt<<<dim3((n_edges + 255) / 256, n_replicas), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        edge_src, n_edges, n_links, delta_tt_link, delta_tt);
    return launch_status();
}

int tarl_store_step_withdraw(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                             const float* attr_in, const tarl_step_io* io, const tarl_agent_table* agents,
                             const tarl_csr* adjacency, int32_t n_nodes, uint8_t* withdrawn, int32_t* counters,
                             float* num_out, int32_t* occupancy, void* stream) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if ((rc = check_step(g, ell, s, attr_in, io)) != TARL_OK) return rc;
    if (ell == nullptr || s.slot_link != nullptr) return TARL_E_BADARG;          // ELL kernels, links in link-id order
    if (agents == nullptr || agents->agent_features == nullptr || agents->n_rows < 1 || adjacency == nullptr ||
        adjacency->ptr == nullptr || (adjacency->n_edges > 0 && adjacency->idx == nullptr) || withdrawn == nullptr ||
        num_out == nullptr || occupancy == nullptr || n_nodes < s.N || n_nodes - s.N > s.N)
        return TARL_E_BADARG;
    if (s.R > 1 && agents->replica_stride < (int64_t)agents->n_rows * 9) return TARL_E_BADARG;
    if (s.N == 0) return TARL_OK;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(occupancy, 0, sizeof(int32_t) * (size_t)s.R, cs) != cudaSuccess) return TARL_E_LAUNCH;
    launch_step(g, ell, s, attr_in, *io, cs, TARL_PHASE_SELECT_APPEND);
    const WithdrawArgs wa = {AgentTable{agents->agent_features, s.R > 1 ? agents->replica_stride : 0, agents->n_rows},
                             *adjacency, withdrawn, counters, num_out, occupancy, nullptr, n_nodes};
    const dim3 grid(blocks_for(s.N), s.R);
    if (ell->width == 4)
        launch_pdl(k_ell_respond_pop_withdraw<4>, grid, cs, *g, *ell, s, io->t, io->pop, io->pop_bits, io->flags, wa,
                   (s.pol_state == kPolKeep ? 1 : 0) * ahead_respond());
    else
        launch_pdl(k_ell_respond_pop_withdraw<8>, grid, cs, *g, *ell, s, io->t, io->pop, io->pop_bits, io->flags, wa,
                   (s.pol_state == kPolKeep ? 1 : 0) * ahead_respond());
    return launch_status();
}

struct tarl_host_pipe {
    cudaStream_t h2d, d2h;
    cudaEvent_t in[2], computed[2], out[2];
    bool used[2];
};

int tarl_host_pipe_create(tarl_host_pipe** pipe) {
    if (pipe == nullptr) return TARL_E_BADARG;
    tarl_host_pipe* p = new tarl_host_pipe();
    bool ok = cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = cudaEventCreateWithFlags(&p->in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->computed[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->out[i], cudaEventDisableTiming) == cudaSuccess;
        p->used[i] = false;
    }
    if (!ok) { delete p; return TARL_E_LAUNCH; }
    *pipe = p;
    return TARL_OK;
}

int tarl_host_pipe_destroy(tarl_host_pipe* p) {
    if (p == nullptr) return TARL_OK;
    cudaStreamSynchronize(p->h2d);
    cudaStreamSynchronize(p->d2h);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(p->in[i]); cudaEventDestroy(p->computed[i]); cudaEventDestroy(p->out[i]); }
    cudaStreamDestroy(p->h2d);
    cudaStreamDestroy(p->d2h);
    delete p;
    return TARL_OK;
}

int tarl_host_pipe_join(tarl_host_pipe* p, void* stream) {
    if (p == nullptr) return TARL_E_BADARG;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < 2; ++i) {
        if (!p->used[i]) continue;
        if (cudaStreamWaitEvent(cs, p->in[i], 0) != cudaSuccess || cudaStreamWaitEvent(cs, p->out[i], 0) != cudaSuccess)
            return TARL_E_LAUNCH;
    }
    return TARL_OK;
}

int tarl_store_step_host(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                         const float* attr_in, const tarl_step_io* io, tarl_host_pipe* p, int32_t slot,
                         const float* sel_host, float* sel_stage, float* dtt_host, uint32_t* pop_bits_host,
                         void* stream) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if ((rc = check_step(g, ell, s, attr_in, io)) != TARL_OK) return rc;
    if (p == nullptr || slot < 0 || slot > 1) return TARL_E_BADARG;
    if ((sel_host != nullptr) != (sel_stage != nullptr)) return TARL_E_BADARG;
    if ((dtt_host != nullptr && io->delta_tt_link == nullptr) || (pop_bits_host != nullptr && io->pop_bits == nullptr))
        return TARL_E_BADARG;
    if (s.N == 0) return TARL_OK;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)s.N * s.R;
    bool ok = true;
    if (sel_host != nullptr) {
        // the staging buffer of this slot was last read by the step taken two calls ago
        if (p->used[slot]) ok = ok && cudaStreamWaitEvent(p->h2d, p->computed[slot], 0) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(sel_stage, sel_host, n * sizeof(float), cudaMemcpyHostToDevice, p->h2d) == cudaSuccess;
        ok = ok && cudaEventRecord(p->in[slot], p->h2d) == cudaSuccess;
        ok = ok && cudaStreamWaitEvent(cs, p->in[slot], 0) == cudaSuccess;
        s.sel = sel_stage;
    }
    // ... and its device outputs are still being downloaded until out[slot] fires
    if (p->used[slot] && (dtt_host != nullptr || pop_bits_host != nullptr))
        ok = ok && cudaStreamWaitEvent(cs, p->out[slot], 0) == cudaSuccess;
    if (!ok) return TARL_E_LAUNCH;
    launch_step(g, ell, s, attr_in, *io, cs, TARL_PHASE_SELECT_APPEND | TARL_PHASE_RESPOND_SHIFT);
    ok = cudaEventRecord(p->computed[slot], cs) == cudaSuccess;
    if (dtt_host != nullptr || pop_bits_host != nullptr) {
        ok = ok && cudaStreamWaitEvent(p->d2h, p->computed[slot], 0) == cudaSuccess;
        if (dtt_host != nullptr)
            ok = ok && cudaMemcpyAsync(dtt_host, io->delta_tt_link, n * sizeof(float), cudaMemcpyDeviceToHost, p->d2h) == cudaSuccess;
        if (pop_bits_host != nullptr)
            ok = ok && cudaMemcpyAsync(pop_bits_host, io->pop_bits, (size_t)s.R * ((s.N + 31) >> 5) * sizeof(uint32_t),
                                       cudaMemcpyDeviceToHost, p->d2h) == cudaSuccess;
    }
    ok = ok && cudaEventRecord(p->out[slot], p->d2h) == cudaSuccess;
    p->used[slot] = true;
    return ok ? launch_status() : TARL_E_LAUNCH;
}

int tarl_store_run(const tarl_dual_csr* g, const tarl_dual_ell* ell, const tarl_link_store* store,
                   const float* attr_in, const tarl_step_io* io, float dt, int32_t n_steps,
                   const float* const* sel_bank, int32_t n_bank, void* stream) {
    Store s;
    int rc = make_store(store, &s);
    if (rc != TARL_OK) return rc;
    if ((rc = check_step(g, ell, s, attr_in, io)) != TARL_OK) return rc;
    if (io->noise != nullptr || n_steps < 0 || n_bank < 0 || (n_bank > 0 && sel_bank == nullptr)) return TARL_E_BADARG;
    if (s.N == 0) return TARL_OK;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    tarl_step_io step = *io;
    for (int i = 0; i < n_steps; ++i) {
        if (n_bank > 0) s.sel = const_cast<float*>(sel_bank[i % n_bank]);
        step.step_id = io->step_id + (uint32_t)i;
        step.t = io->t + dt * (float)i;
        launch_step(g, ell, s, attr_in, step, cs, TARL_PHASE_SELECT_APPEND | TARL_PHASE_RESPOND_SHIFT);
        float4* written = s.hot_next;                       // ping-pong: what this step wrote is the next step's input
        s.hot_next = const_cast<float4*>(s.hot_cur);
        s.hot_cur = written;
    }
    return launch_status();
}

}  // extern "C"

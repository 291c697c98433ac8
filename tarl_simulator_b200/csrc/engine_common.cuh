// engine_common.cuh — shared device helpers of the resident link store kernels (engine.cu, agents.cu).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace tarl {

#ifndef TARL_THREADS
#define TARL_THREADS 128
#endif
constexpr int kThreads = TARL_THREADS;
constexpr int kMetaRingMask = 0xffff;
constexpr int kMetaGarbage = 1 << 16;

// torch.maximum / torch.clamp(min=) propagate NaN; fmaxf does not.
__device__ __forceinline__ float max_propagate_nan(float a, float b) {
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: one call yields four uniforms in [2^-24, 1 - 2^-24].
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, float out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;   // one IMAD.WIDE each
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0; c1 = (uint32_t)p1; c2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = ((float)(c[i] >> 9) + 0.5f) * (1.0f / 8388608.0f);   // exact: in [2^-24, 1-2^-24]
}

// Device view of struct tarl_link_store. One 32-byte record per (replica, link), two float4 halves:
//   A = {head id, head exit time, NUM, MAXN}          everything a NEIGHBOUR needs: gathered with one 128-bit load
//   B = {head arrival time, tail id, pending tail-garbage exit time, meta}   owner only
// meta = ring head (16 bit) | kMetaGarbage ("the reference's (0, t, t+tt) tail write is pending at slot int(NUM)").
struct Store {
    int N, R, Nmax, M;       // M = Nmax-1 ring slots per link
    const float4* hot_cur;   // [R*N*2]
    float4* hot_next;        // [R*N*2]
    float* sel;              // [R*N]
    const float4* stat_a;    // [N] {FFTT, cc, ROAD_INDEX, weight shared by all in-edges of the link or NaN (uni_hint)}
    const float4* stat_b;    // [N] {LENGTH, MAX_FLOW, 0, 0}
    float4* queue;           // [R*N*M]
    float2* post;            // [R*N] {NUM, tail id} after the direction phase
    bool uni_hint;           // stat_a.w is filled in (TARL_STORE_UNIFORM_WEIGHTS): the ELL direction kernel reads it
                             // instead of the link's edge-weight column
    const int32_t* slot_link;   // [N] slot -> link id (nullptr = identity): the store's own locality order
    const int32_t* link_slot;   // [N] link id -> slot
    int pol_state, pol_static;  // L2 eviction policy (kPol*) of the records / summaries and of topology + statics
    __device__ __forceinline__ int link_of(int slot) const { return slot_link != nullptr ? slot_link[slot] : slot; }
    __device__ __forceinline__ int slot_of(int link) const { return link_slot != nullptr ? link_slot[link] : link; }
};

__device__ __forceinline__ int ring_pos(int rh, int logical, int M) {  // logical slot 1..M -> physical 0..M-1
    int p = rh + logical - 1;
    return p >= M ? p - M : p;
}

// (long)a == (long)b and (long)a > 0 of src/response_mpnn.py:66-83 on fp32 operands, without 64-bit conversions
__device__ __forceinline__ bool same_id(float a, float b) { return truncf(a) == truncf(b); }
__device__ __forceinline__ bool at_least_one(float a) { return a >= 1.0f; }

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while the previous kernel of the stream is still draining. pdl_trigger() lets the NEXT kernel's CTAs be
// scheduled as soon as every CTA of this grid has started; pdl_wait() blocks until the PREVIOUS grid has completed and
// its writes are visible. Everything a kernel does before pdl_wait() must therefore touch only data no kernel of the
// step writes (topology, link statics). Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- L2 residency (sm_100a)
// One step of one million links streams ~200 MB through a 126 MB L2, of which ~75 MB are the state that the NEXT kernel
// reads again (the 32-byte records in both ping-pong buffers, the {NUM, tail} summaries); the rest is topology and link
// statics, read once per step. Left to the default policy the stream evicts the state between two kernels and every
// kernel pays DRAM latency on both of its dependent load levels. make_store therefore picks, per store,
//   state fits (R*N*80 B <= 96 MB): records / summaries L2::evict_last, topology and statics L2::evict_first when only
//                                   one replica reads them (R == 1), default otherwise;
//   state does not fit (many replicas): records default, topology / statics (shared by all R replicas) L2::evict_last;
// as createpolicy descriptors handed to every access (ld/st.global.L2::cache_hint). Topology loads also skip the L1
// (L1::no_allocate): the L1 is kept for the neighbours' records, which a tile re-reads.
#ifndef TARL_L2_KEEP
#define TARL_L2_KEEP 1
#endif
enum { kPolDefault = 0, kPolKeep = 1, kPolStream = 2 };

__device__ __forceinline__ uint64_t l2_policy(int kind) {
    uint64_t pol;
    if (kind == kPolKeep) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == kPolStream) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ float4 ld_hint(const float4* p, uint64_t pol) {
#if TARL_L2_KEEP
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ float2 ld_hint(const float2* p, uint64_t pol) {
#if TARL_L2_KEEP
    float2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_hint(float4* p, const float4 v, uint64_t pol) {
#if TARL_L2_KEEP
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_hint(float2* p, const float2 v, uint64_t pol) {
#if TARL_L2_KEEP
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
// The 32-byte record {A, B} of one (replica, link), p = &hot[2 * L] (32-byte aligned: make_store checks the base), moved
// by ONE 256-bit access (sm_100a: ld/st.global.v8.b32, which also takes the L2 eviction priority as a qualifier). Half
// the L1 requests of two 128-bit accesses: measured 59.8 -> 55.7 us per step of one million links.
// NOT for use inside __noinline__ device functions: ptxas 12.9 emits a 32-bit STG for the 256-bit store there (the CSR
// fallback lost 28 of every 32 bytes; caught by tests/test_link_store_gpu.py) — those use the *_narrow forms.
__device__ __forceinline__ void ld_record(const float4* p, float4& A, float4& B, bool keep) {
#if TARL_L2_KEEP
    uint32_t r[8];
    if (keep)
        asm volatile("ld.global.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    else
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    A = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
    B = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
#else
    A = p[0]; B = p[1];
#endif
}
__device__ __forceinline__ void st_record(float4* p, const float4 A, const float4 B, bool keep) {
#if TARL_L2_KEEP
    if (keep)
        asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "r"(__float_as_uint(A.x)), "r"(__float_as_uint(A.y)), "r"(__float_as_uint(A.z)),
                        "r"(__float_as_uint(A.w)), "r"(__float_as_uint(B.x)), "r"(__float_as_uint(B.y)),
                        "r"(__float_as_uint(B.z)), "r"(__float_as_uint(B.w)) : "memory");
    else
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "r"(__float_as_uint(A.x)), "r"(__float_as_uint(A.y)), "r"(__float_as_uint(A.z)),
                        "r"(__float_as_uint(A.w)), "r"(__float_as_uint(B.x)), "r"(__float_as_uint(B.y)),
                        "r"(__float_as_uint(B.z)), "r"(__float_as_uint(B.w)) : "memory");
#else
    p[0] = A; p[1] = B;
#endif
}
__device__ __forceinline__ void st_record_narrow(float4* p, const float4 A, const float4 B, uint64_t pol) {
    st_hint(p, A, pol); st_hint(p + 1, B, pol);
}
// topology columns and link statics: never written by a step
__device__ __forceinline__ float4 ld_static(const float4* p, uint64_t pol) {
#if TARL_L2_KEEP
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ int ld_static(const int32_t* p, uint64_t pol) {
#if TARL_L2_KEEP
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ float ld_static(const float* p, uint64_t pol) {
#if TARL_L2_KEEP
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}

// Ask the L2 for [base + byte_off, + bytes) (shrunk to 16-byte boundaries): one instruction, no register, no wait.
__device__ __forceinline__ void l2_prefetch_span(const void* base, size_t byte_off, uint32_t bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base) + byte_off;
    const uintptr_t lo = (a + 15) & ~(uintptr_t)15, hi = (a + bytes) & ~(uintptr_t)15;
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
}

int make_store(const tarl_link_store* p, Store* s);

}  // namespace tarl

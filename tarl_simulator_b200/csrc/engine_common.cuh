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
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
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
    const float4* stat_a;    // [N] {FFTT, cc, ROAD_INDEX, MAXN}
    const float4* stat_b;    // [N] {LENGTH, MAX_FLOW, 0, 0}
    float4* queue;           // [R*N*M]
    float2* post;            // [R*N] {NUM, tail id} after the direction phase
    uint8_t* hint;           // [R*N] 1 = a downstream link admitted this link's head in the direction phase
    const int32_t* slot_link;   // [N] slot -> link id (nullptr = identity): the store's own locality order
    const int32_t* link_slot;   // [N] link id -> slot
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

int make_store(const tarl_link_store* p, Store* s);

}  // namespace tarl

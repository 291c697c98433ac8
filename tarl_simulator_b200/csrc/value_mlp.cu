// value_mlp.cu — MPNNValueNetSimple's forward on the 5th-generation tensor cores (sm_100a: tcgen05.mma, TMEM, TMA).
//
// Reference semantics: /root/reference/src/agents/mpnn_agent.py:407-450 —
//   v = Linear(64,1)( relu( Linear(64,64)( relu( Linear(N_tot+1, 64)( [NUMBER_OF_AGENT[:, 0:N_tot] ‖ time] ) ) ) ) )
// evaluated for M observation rows at once (every frame of a PPO rollout: M = replicas, K = N_tot + 1 ~ 6e4..1.5e6).
// The first layer is a skinny GEMM  H[M,64] = A[M,K] · W1[64,K]^T  that reads A exactly once: 32 flop per byte of A,
// i.e. HBM-bound on the tensor pipe and compute-bound on the fp32 SIMT pipe — the one place on this path where tensor
// cores pay. The parity bar is 1e-5 relative in fp32, which plain TF32 (10-bit mantissa) misses, so the product is
// error-compensated: a = a_hi + a_lo, w = w_hi + w_lo with *_hi the value rounded to TF32 and *_lo the (rounded)
// remainder, and
//   a·w ~= a_hi·w_hi + a_lo·w_hi + a_hi·w_lo          (three kind::tf32 MMAs, fp32 accumulation in TMEM)
// which leaves a relative error of ~2^-23 per product with a random sign.
//
// Data flow per CTA (one 128-row tile of A x one K-slice), 7 warps:
//   warps 0, 6  TMA producers: A tiles [128 x 32] fp32 into an 8-slot shared-memory ring, the matching [64 x 32]
//               tiles of W_hi / W_lo into a 4-slot ring (128-byte swizzle), completion on mbarriers
//               (cp.async.bulk.tensor).
//   warps 2-5   splitters: thread r owns row r of the tile — reads its 128 bytes from shared memory, splits each value
//               into hi/lo and stores both into TENSOR MEMORY (tcgen05.st, lane r, 32 + 32 columns per stage). A is
//               then an MMA operand straight from TMEM: its hi/lo copies never touch shared memory, whose bandwidth
//               would otherwise cap the kernel (3 MMAs x (A + B) re-read per k-step).
//   warp 1      MMA issuer (one thread): per k-step of 8 columns two tcgen05.mma.kind::tf32 (A from TMEM, B from
//               shared memory through a K-major SWIZZLE_128B descriptor): a_hi x [w_hi; w_lo] (N = 128) and
//               a_lo x w_hi (N = 64) into a 128 x 128 fp32 accumulator in TMEM; tcgen05.commit releases the ring
//               slots / publishes the accumulator.
//   warps 2-5   promotion: the tensor core adds into its fp32 accumulator with truncation, a bias that over the
//               ~2e4 accumulation steps of a K = 6e4 row reaches 1e-4 relative (measured). Every 256 columns the MMA
//               therefore switches between two TMEM accumulators and these warps tcgen05.ld the finished one and
//               add it into fp32 registers with round-to-nearest (the same remedy as for FP8 accumulators); at the
//               end one row per thread goes to the split-K partial buffer.
// A second small kernel sums the K-slices in a fixed order (deterministic), adds the time column and the bias and
// runs the 64x64 and 64x1 layers per row.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tarl_b200.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 32;
#ifndef TARL_VM_GROUP
#define TARL_VM_GROUP 2
#endif
constexpr int kGroup = TARL_VM_GROUP;                              // A tiles requested together (NA must be a multiple)
// ring depths in k-blocks: A tiles (smem), A hi/lo (TMEM: 2 pairs), W tiles (smem: 3 pairs). The last two are released
// by the same event (the MMAs of a pair completing) through ONE barrier per pair, pair p on barrier p % kPairBars: a
// tcgen05.commit costs the issuing thread ~150 cycles (measured), more than a third of the 384-cycle MMA floor of a
// k-block, so there is exactly one per pair.
#ifndef TARL_VM_NA
#define TARL_VM_NA 8
#endif
#ifndef TARL_VM_NW
#define TARL_VM_NW 6
#endif
constexpr int NA = TARL_VM_NA, NT = 4, NW = TARL_VM_NW;
constexpr int kPairBars = 6;                           // a multiple of NT / 2 and NW / 2, at least their maximum + 1
constexpr int kHidden = 64;
constexpr int kSplitSets = 2;                          // splitter warp quartets, k-block kb goes to set kb % kSplitSets
constexpr int kThreadsGemm = 32 * (2 + 4 * kSplitSets + 1);
constexpr uint32_t kABytes = BM * BK * 4, kBBytes = BN * BK * 4;
constexpr uint32_t kTmemCols = 512;                    // 2 x 128 accumulator + 4 x (32 hi + 32 lo)
constexpr uint32_t kColAcc = 0, kAccCols = 2 * BN, kColA = 2 * kAccCols;
constexpr int kChunk = 8;                              // k-blocks (256 columns of A) accumulated in TMEM before promotion
constexpr size_t kSmemBytes = (size_t)NA * kABytes + (size_t)NW * 2 * kBBytes + 1024 /* alignment slack */ + 512 /* barriers */;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, one 128 x 64 x 8 TF32 step
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// K-major operand tile in shared memory, rows of 128 bytes, SWIZZLE_128B (what the TMA wrote): 8-row groups are
// 1024 bytes apart (stride byte offset), the leading offset is unused for swizzled K-major layouts, descriptor
// version 1 (sm_100). The tile base must be 1024-byte aligned. Stepping 8 TF32 columns = +32 bytes on the address.
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (ignored), bits [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version, bits [46,48)
    d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }
constexpr uint32_t kIdesc64 = idesc_tf32(BN), kIdesc128 = idesc_tf32(2 * BN);

constexpr uint32_t kHiMask = 0xFFFFE000u;               // sign + exponent + the 10 mantissa bits TF32 keeps
// Round to nearest TF32 (ties away from zero). Truncation would do for the arithmetic, but its error always points
// towards zero: over K ~ 6e4 same-signed products the 2^-21 relative bias does not average out (measured 3e-5 on the
// output), whereas rounded splits leave ~2^-23 per product with a random sign. The MMA then finds nothing to truncate.
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & kHiMask); }

// ------------------------------------------------------------------------------------------------ weight split
// W1 [64, K] (K = n_nodes + 1, rows not 16-byte aligned in general) -> W_hi, W_lo [64, Kp] (Kp a multiple of 32, zero
// padded, TMA-addressable) and the time column w_time [64].
__global__ void __launch_bounds__(256) k_value_mlp_split_w(const float* __restrict__ w1, int n_nodes, int Kp,
                                                           float* __restrict__ w_hi, float* __restrict__ w_lo,
                                                           float* __restrict__ w_time) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)kHidden * Kp) return;
    const int j = (int)(i / Kp), k = (int)(i - (int64_t)j * Kp);
    float w = 0.0f;
    if (k < n_nodes) w = w1[(int64_t)j * (n_nodes + 1) + k];
    const float hi = tf32_rn(w);
    w_hi[i] = hi;
    w_lo[i] = tf32_rn(w - hi);
    if (k == 0) w_time[j] = w1[(int64_t)j * (n_nodes + 1) + n_nodes];
}

// ------------------------------------------------------------------------------------------------ first layer
#ifdef TARL_VM_PROFILE                                 // per-role cycle accounting of one CTA (profiles/value_mlp_roles.py)
#define PROF_DECL long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc_ = clock64(), pstart_ = pc_
#define PROF(i) do { const long long n_ = clock64(); pt_[i] += n_ - pc_; pc_ = n_; } while (0)
#define PROF_OUT(name) do { if (blockIdx.x == 1 && blockIdx.y == 3) printf("%s nkb %d total %lld | %lld %lld %lld %lld %lld %lld %lld %lld\n", name, nkb, clock64() - pstart_, pt_[0], pt_[1], pt_[2], pt_[3], pt_[4], pt_[5], pt_[6], pt_[7]); } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_OUT(name)
#endif
// Three rings decouple the three latencies: A tiles in shared memory (TMA -> splitters; a slot is free again as soon
// as its rows sit in registers, so its cycle is HBM latency + split, not + MMA), A hi/lo in TMEM (splitters -> MMA),
// W tiles in shared memory (TMA from L2 -> MMA).
//
// Everything downstream of the A ring works on PAIRS of k-blocks (2 x 32 columns). Measured with the per-role counters
// above: the issuing thread is the pacemaker — an mbarrier test costs it ~200 cycles even when the barrier completed
// long ago (its shared-memory round trip queues behind the splitter warps' LDS bursts), a tcgen05.commit ~150, and a
// tcgen05.mma blocks until the previous one is about to finish (64 cycles at N = 128, 45 at N = 64:
// profiles/micro/mma_rate.cu), so that per k-block two waits + eight MMAs + one commit came to ~1300 cycles against
// the MMAs' own 384. Per pair there is ONE wait — `ready[p % 6]` collects the 256 splitter arrivals AND the W tiles'
// bytes (the W producer's arrive.expect_tx) — sixteen MMAs and ONE commit: `pair_done[p % 6]`, which frees the TMEM
// columns two pairs later (2 pair slots) and the W tiles three pairs later (3 pair slots) through the same barrier.
__global__ void __launch_bounds__(kThreadsGemm, 1) k_value_mlp_gemm(const __grid_constant__ CUtensorMap map_a,
                                                                    const __grid_constant__ CUtensorMap map_wh,
                                                                    const __grid_constant__ CUtensorMap map_wl, int M,
                                                                    int kb_total, int kb_per_slice,
                                                                    float* __restrict__ partials) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B tiles: 1024-byte aligned
    const uint32_t w_base = base + NA * kABytes;
    const uint32_t bars = w_base + NW * 2 * kBBytes;
    auto sm_a = [&](int s) { return base + s * kABytes; };
    auto sm_wh = [&](int s) { return w_base + s * 2 * kBBytes; };
    auto sm_wl = [&](int s) { return w_base + s * 2 * kBBytes + kBBytes; };
    auto a_full = [&](int s) { return bars + 8u * s; };
    auto a_empty = [&](int s) { return bars + 8u * (NA + s); };
    auto ready = [&](int s) { return bars + 8u * (2 * NA + s); };             // pair p -> ready(p % kPairBars)
    auto pair_done = [&](int s) { return bars + 8u * (2 * NA + kPairBars + s); };
    auto accfull = [&](int b) { return bars + 8u * (2 * NA + 2 * kPairBars + b); };
    auto accfree = [&](int b) { return bars + 8u * (2 * NA + 2 * kPairBars + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * NA + 2 * kPairBars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int kb0 = blockIdx.y * kb_per_slice;
    const int nkb = min(kb_per_slice, kb_total - kb0);
    const int n_chunks = (nkb + kChunk - 1) / kChunk;
    const int n_pairs = (nkb + 1) / 2;
    // the pair whose MMAs must be complete before pair p may reuse its TMEM columns / its W slots
    auto wait_pair_done = [&](int p) { if (p >= 0) mbar_wait(pair_done(p % kPairBars), (p / kPairBars) & 1); };

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");      // descriptors: kernel parameters
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        for (int s = 0; s < NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 128); }
        for (int s = 0; s < kPairBars; ++s) { mbar_init(ready(s), 128 * kSplitSets + 1); mbar_init(pair_done(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(accfull(b), 1); mbar_init(accfree(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                     // TMEM allocation: one whole warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // launched with programmatic stream serialization: barrier init and the TMEM allocation above overlap the drain
    // of whatever kernel precedes (the previous call's tail kernel in a rollout); its results — the observation, the
    // split weights, the partial-sum buffer the previous tail still reads — are touched only after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the tail kernel may stage its weights now

    if (warp == 0) {
        if (lane == 0) {                                                 // ===== TMA producer, A tiles (HBM)
            // kGroup consecutive k-blocks are requested back to back: per row of A that is kGroup x 128 contiguous
            // bytes arriving at the memory controller together instead of 128-byte pieces one k-block period apart
            PROF_DECL;
            for (int kg = 0; kg < nkb; kg += kGroup) {
                const int n = min(kGroup, nkb - kg);
                PROF(0);
                for (int i = 0; i < n; ++i) mbar_wait(a_empty((kg + i) % NA), (((kg + i) / NA) & 1) ^ 1);
                PROF(1);
                for (int i = 0; i < n; ++i) {
                    const int kb = kg + i, s = kb % NA;
                    mbar_expect_tx(a_full(s), kABytes);
                    tma_load_2d(sm_a(s), &map_a, a_full(s), (kb0 + kb) * BK, m0);
                }
            }
            PROF_OUT("tmaA [other, wait a_empty]");
        }
    } else if (warp == 2 + 4 * kSplitSets) {
        if (lane == 0) {                                                 // ===== TMA producer, W tiles (L2)
            for (int p = 0; p < n_pairs; ++p) {
                const int n = min(2, nkb - 2 * p);
                wait_pair_done(p - NW / 2);                              // the pair that read these W slots
                mbar_expect_tx(ready(p % kPairBars), (uint32_t)n * 2 * kBBytes);
                for (int j = 0; j < n; ++j) {
                    const int kb = 2 * p + j, s = kb % NW;
                    tma_load_2d(sm_wh(s), &map_wh, ready(p % kPairBars), (kb0 + kb) * BK, 0);
                    tma_load_2d(sm_wl(s), &map_wl, ready(p % kPairBars), (kb0 + kb) * BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                                 // ===== MMA issuer
            PROF_DECL;
            const uint64_t dw0 = kmajor_sw128_desc(sm_wh(0));
            for (int p = 0; p < n_pairs; ++p) {
                const int kb = 2 * p;
                const int chunk = kb / kChunk, b = chunk & 1, first = (kb % kChunk) == 0;
                PROF(0);
                if (first) mbar_wait(accfree(b), ((chunk >> 1) & 1) ^ 1);    // accumulator b drained (free at start)
                PROF(1);
                mbar_wait(ready(p % kPairBars), (p / kPairBars) & 1);    // A hi/lo in TMEM, W tiles in shared memory
                PROF(2);
                tc_fence_after();
                // The W slot holds W_hi (rows 0-63) directly followed by W_lo (rows 64-127): ONE N = 128 MMA gives
                // a_hi.w_hi (accumulator columns 0-63) and a_hi.w_lo (columns 64-127); a second, N = 64, adds
                // a_lo.w_hi onto columns 0-63.
                const uint32_t acc = tmem_base + kColAcc + b * kAccCols;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (kb + j < nkb) {
                        const uint64_t dw = dw0 + (uint64_t)((((kb + j) % NW) * 2 * kBBytes) >> 4);
                        const uint32_t a_hi = tmem_base + kColA + ((kb + j) % NT) * 64, a_lo = a_hi + 32;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {     // interleaved: 48 cycles per MMA against 64 / 45 back to back
                            tc_mma_tf32_ts(acc, a_hi + 8 * k, dw + 2 * k, kIdesc128, (first && j == 0 && k == 0) ? 0u : 1u);
                            tc_mma_tf32_ts(acc, a_lo + 8 * k, dw + 2 * k, kIdesc64, 1u);
                        }
                    }
                }
                PROF(3);
                tc_commit(pair_done(p % kPairBars));                     // TMEM A columns and W slots reusable
                if ((kb % kChunk) == kChunk - 2 || p == n_pairs - 1) tc_commit(accfull(b));
                PROF(4);
            }
            PROF_OUT("mma [other, wait accfree, wait ready, mma issue, commit]");
        }
    } else {                                                             // ===== splitters + promotion (warps 2-9)
        // Two quartets of splitter warps: quartet s takes k-block 2p + s of pair p. Quartet s also promotes the
        // chunks with (chunk & 1) == s, i.e. always accumulator s, into its own fp32 sums; the two sums are added
        // through shared memory at the end (fixed order: deterministic).
        const int q = warp & 3;                                          // the TMEM lane quarter this warp may touch
        const int set = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.0f;
        auto promote = [&](int chunk) {                                  // sum += accumulator of `chunk` (fp32, RN)
            const int b = chunk & 1;
            mbar_wait(accfull(b), (chunk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 4; ++h) {                                // columns 64-127 (a_hi.w_lo) fold onto 0-63
                uint32_t v[32];
                tc_ld32(lane_addr + kColAcc + b * kAccCols + h * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[(h & 1) * 32 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            mbar_arrive(accfree(b));
        };
        int next_chunk = set;                                            // the next chunk this quartet promotes
        PROF_DECL;
        for (int p = 0; p < n_pairs; ++p) {
            const int kb = 2 * p + set;
            if (kb >= nkb) {                                             // odd tail: the pair's barrier still counts us
                mbar_arrive(ready(p % kPairBars));
                break;
            }
            const int sa = kb % NA, pa = (kb / NA) & 1, st = kb % NT;
            PROF(0);
            mbar_wait(a_full(sa), pa);
            PROF(1);
            const uint32_t row_addr = sm_a(sa) + row * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t chunk = (uint32_t)c ^ (uint32_t)(row & 7);                      // SWIZZLE_128B
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(row_addr + (chunk << 4)));
                const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = tf32_rn(f[e]);
                    hi[4 * c + e] = __float_as_uint(h);
                    lo[4 * c + e] = __float_as_uint(tf32_rn(f[e] - h));
                }
            }
            mbar_arrive(a_empty(sa));                                    // the row is in registers: slot back to the TMA
            PROF(2);
            wait_pair_done(p - NT / 2);                                  // MMAs that read these TMEM columns are done
            PROF(3);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t h16[16], l16[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { h16[j] = hi[half * 16 + j]; l16[j] = lo[half * 16 + j]; }
                tc_st16(lane_addr + kColA + st * 64 + half * 16, h16);
                tc_st16(lane_addr + kColA + st * 64 + 32 + half * 16, l16);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(ready(p % kPairBars));
            PROF(4);
            // Half a chunk behind the MMAs: pair 2 of chunk c+1 got its TMEM columns, so the MMAs of chunk c are
            // complete and committed — the wait inside promote() returns at once and the issuing thread finds the
            // accumulator free long before it needs it again (chunk c+2).
            if ((kb % kChunk) / 2 == kChunk / 4 && next_chunk <= kb / kChunk - 1) {
                promote(next_chunk);
                next_chunk += kSplitSets;
                PROF(5);
            }
        }
        for (; next_chunk < n_chunks; next_chunk += kSplitSets) promote(next_chunk);
        if (lane == 0 && q == 0) PROF_OUT(set ? "split1 [other, wait a_full, lds+split, wait pair_done, st, promote]" : "split0 [other, wait a_full, lds+split, wait pair_done, st, promote]");
        // quartet 1 parks its sums in the (drained) A ring, [j][row]; quartet 0 adds them
        float* xch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
        asm volatile("bar.sync 1, %0;" ::"r"(128 * kSplitSets) : "memory");     // every quartet has read its last tile
        if (set > 0) {
#pragma unroll
            for (int j = 0; j < BN; ++j) xch[((set - 1) * BN + j) * BM + row] = sum[j];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(128 * kSplitSets) : "memory");
        if (set == 0) {
#pragma unroll
            for (int s2 = 1; s2 < kSplitSets; ++s2)
#pragma unroll
                for (int j = 0; j < BN; ++j) sum[j] += xch[((s2 - 1) * BN + j) * BM + row];
            const int m = m0 + row;
            if (m < M) {
                float4* out = reinterpret_cast<float4*>(partials + ((size_t)blockIdx.y * M + m) * kHidden);
#pragma unroll
                for (int c = 0; c < BN / 4; ++c) out[c] = make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------------------------------------ layers 2 and 3
// One WARP per observation row, lane j owns hidden units j and j + 32: h1 = relu(sum of the K-slices (ascending,
// coalesced 256-byte rows) + time * w_time + b1); h2 = relu(W2 h1 + b2) with W2 transposed in shared memory (lanes
// read consecutive words, h1 is a broadcast); v = w3 . h2 + b3 by a shuffle tree.
constexpr int kTailWarps = 8;
__global__ void __launch_bounds__(kTailWarps * 32) k_value_mlp_tail(const float* __restrict__ partials, int n_slices, int M,
                                                                    const float* __restrict__ time, int64_t time_stride,
                                                                    const float* __restrict__ w_time,
                                                                    const float* __restrict__ b1,
                                                                    const float* __restrict__ w2,
                                                                    const float* __restrict__ b2,
                                                                    const float* __restrict__ w3,
                                                                    const float* __restrict__ b3, float* __restrict__ out,
                                                                    float* __restrict__ save_z1,
                                                                    float* __restrict__ save_z2) {
    __shared__ float s_w2t[kHidden][kHidden + 1];        // [i][j] = W2[j][i]
    __shared__ float s_h1[kTailWarps][kHidden];
    // launched with programmatic stream serialization behind the GEMM: the weights (which no kernel of the call
    // writes) are staged while the GEMM's last CTAs drain, the partial sums are touched only after the wait
    for (int i = threadIdx.x; i < kHidden * kHidden; i += blockDim.x) s_w2t[i % kHidden][i / kHidden] = w2[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * kTailWarps + warp;
    const float wt0 = w_time[lane], wt1 = w_time[lane + 32], b1a = b1[lane], b1b = b1[lane + 32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the next call's GEMM may set itself up
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    if (m >= M) return;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 9
    for (int s = 0; s < n_slices; ++s) {
        const float* p = partials + ((size_t)s * M + m) * kHidden;
        a0 += p[lane];
        a1 += p[lane + 32];
    }
    const float t = time[m * time_stride];
    const float z1a = a0 + t * wt0 + b1a, z1b = a1 + t * wt1 + b1b;
    if (save_z1 != nullptr) { save_z1[(size_t)m * kHidden + lane] = z1a; save_z1[(size_t)m * kHidden + lane + 32] = z1b; }
    s_h1[warp][lane] = fmaxf(z1a, 0.0f);
    s_h1[warp][lane + 32] = fmaxf(z1b, 0.0f);
    __syncwarp();
    float c0 = b2[lane], c1 = b2[lane + 32];
#pragma unroll 16
    for (int i = 0; i < kHidden; ++i) {
        const float h = s_h1[warp][i];
        c0 += s_w2t[i][lane] * h;
        c1 += s_w2t[i][lane + 32] * h;
    }
    if (save_z2 != nullptr) { save_z2[(size_t)m * kHidden + lane] = c0; save_z2[(size_t)m * kHidden + lane + 32] = c1; }
    float v = w3[lane] * fmaxf(c0, 0.0f) + w3[lane + 32] * fmaxf(c1, 0.0f);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[m] = v + b3[0];
}

// ------------------------------------------------------------------------------------------------ backward
// Gradient of sum_m grad_out[m] * out[m] w.r.t. the six parameter tensors (observations are leaves: no input gradient).
// With z1 / z2 the pre-activations the forward pass kept:
//   g_z2 = grad_out (x) w3 . [z2 > 0]        dW3 = sum_m grad_out relu(z2)     db3 = sum grad_out    db2 = sum_m g_z2
//   g_z1 = (g_z2 W2) . [z1 > 0]              dW2 = g_z2^T relu(z1)             db1 = sum_m g_z1
//   dW1[:, :N] = g_z1^T A   (k_value_mlp_dw1)                                  dW1[:, N] = g_z1^T time
// Everything but dW1 is a few thousand flops per row: one CTA, sums over the rows in ascending order (deterministic).
constexpr int kBwdThreads = 1024;
__global__ void __launch_bounds__(kBwdThreads) k_value_mlp_bwd_small(int M, const float* __restrict__ z1,
                                                                     const float* __restrict__ z2,
                                                                     const float* __restrict__ grad_out,
                                                                     const float* __restrict__ time, int64_t time_stride,
                                                                     const float* __restrict__ w2,
                                                                     const float* __restrict__ w3,
                                                                     float* __restrict__ g_z1, float* __restrict__ g_z2,
                                                                     float* __restrict__ d_wtime, float* __restrict__ db1,
                                                                     float* __restrict__ dw2, float* __restrict__ db2,
                                                                     float* __restrict__ dw3, float* __restrict__ db3) {
    __shared__ float s_w2[kHidden][kHidden + 1];          // [j][i] = W2[j][i]
    const int tid = threadIdx.x;
    for (int i = tid; i < kHidden * kHidden; i += kBwdThreads) s_w2[i / kHidden][i % kHidden] = w2[i];
    for (int i = tid; i < M * kHidden; i += kBwdThreads) {
        const int m = i / kHidden, j = i - m * kHidden;
        g_z2[i] = z2[i] > 0.0f ? grad_out[m] * w3[j] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < M * kHidden; i += kBwdThreads) {
        const int m = i / kHidden, c = i - m * kHidden;
        float acc = 0.0f;
#pragma unroll 8
        for (int j = 0; j < kHidden; ++j) acc += g_z2[m * kHidden + j] * s_w2[j][c];
        g_z1[i] = z1[i] > 0.0f ? acc : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kHidden * kHidden; i += kBwdThreads) {      // dW2[j][c] = sum_m g_z2[m][j] relu(z1[m][c])
        const int j = i / kHidden, c = i - j * kHidden;
        float acc = 0.0f;
        for (int m = 0; m < M; ++m) acc += g_z2[m * kHidden + j] * fmaxf(z1[m * kHidden + c], 0.0f);
        dw2[i] = acc;
    }
    if (tid < kHidden) {
        float a3 = 0.0f, a2 = 0.0f, a1 = 0.0f, at = 0.0f;
        for (int m = 0; m < M; ++m) {
            a3 += grad_out[m] * fmaxf(z2[m * kHidden + tid], 0.0f);
            a2 += g_z2[m * kHidden + tid];
            const float g = g_z1[m * kHidden + tid];
            a1 += g;
            at += g * time[m * time_stride];
        }
        dw3[tid] = a3; db2[tid] = a2; db1[tid] = a1; d_wtime[tid] = at;
    }
    if (tid == kHidden) {
        float a = 0.0f;
        for (int m = 0; m < M; ++m) a += grad_out[m];
        db3[0] = a;
    }
}

// dW1[j, n] = sum_m g_z1[m, j] * A[m, n]: the occupancy matrix is read once (coalesced over n), the [64, N + 1] gradient
// written once — 16 flop per byte moved with a reduction only M (= 32 frames in the PPO update) deep: HBM-bound on the
// plain fp32 pipe, exact fp32 accumulation in ascending row order, no tensor cores needed. One thread per column n
// keeps the 64 partial sums of its column in registers; g_z1 goes through shared memory in chunks of kDwRows rows.
constexpr int kDwThreads = 128, kDwRows = 32;
__global__ void __launch_bounds__(kDwThreads) k_value_mlp_dw1(const float* __restrict__ a, int64_t a_stride, int M,
                                                              int n_nodes, const float* __restrict__ g_z1,
                                                              const float* __restrict__ d_wtime,
                                                              float* __restrict__ dw1) {
    __shared__ float s_g[kDwRows][kHidden];
    const int n = blockIdx.x * kDwThreads + threadIdx.x;
    const bool live = n < n_nodes;
    float acc[kHidden];
#pragma unroll
    for (int j = 0; j < kHidden; ++j) acc[j] = 0.0f;
    for (int m0 = 0; m0 < M; m0 += kDwRows) {
        const int rows = min(kDwRows, M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * kHidden; i += kDwThreads) s_g[i / kHidden][i % kHidden] = g_z1[(size_t)m0 * kHidden + i];
        __syncthreads();
        if (live) {
            float av[kDwRows];
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) av[r] = r < rows ? a[(int64_t)(m0 + r) * a_stride + n] : 0.0f;
#pragma unroll
            for (int r = 0; r < kDwRows; ++r) {
                if (r < rows) {
#pragma unroll
                    for (int j = 0; j < kHidden; ++j) acc[j] += s_g[r][j] * av[r];
                }
            }
        }
    }
    const int64_t pitch = (int64_t)n_nodes + 1;
    if (live) {
#pragma unroll
        for (int j = 0; j < kHidden; ++j) dw1[j * pitch + n] = acc[j];
    }
    if (blockIdx.x == 0 && threadIdx.x < kHidden) dw1[threadIdx.x * pitch + n_nodes] = d_wtime[threadIdx.x];   // time column
}

// ------------------------------------------------------------------------------------------------ host side
struct Plan {
    int Kp, kb_total, tiles, slices, kb_per_slice;
    size_t off_wh, off_wl, off_wt, off_part, bytes;
};

Plan make_plan(int M, int n_nodes) {
    Plan p;
    p.kb_total = (n_nodes + BK - 1) / BK;
    p.Kp = p.kb_total * BK;
    p.tiles = (M + BM - 1) / BM;
    int want = 148 / (p.tiles > 0 ? p.tiles : 1);        // K-slices so that tiles x slices fills the 148 SMs once
    if (want < 1) want = 1;
    if (want > p.kb_total) want = p.kb_total > 0 ? p.kb_total : 1;
    p.kb_per_slice = (p.kb_total + want - 1) / want;
    if (p.kb_per_slice < 1) p.kb_per_slice = 1;
    p.slices = p.kb_total > 0 ? (p.kb_total + p.kb_per_slice - 1) / p.kb_per_slice : 0;
    p.off_wh = 0;
    p.off_wl = align_up(p.off_wh + (size_t)kHidden * p.Kp * 4, 1024);
    p.off_wt = align_up(p.off_wl + (size_t)kHidden * p.Kp * 4, 1024);
    p.off_part = align_up(p.off_wt + kHidden * 4, 1024);
    p.bytes = align_up(p.off_part + (size_t)p.slices * M * kHidden * 4, 1024);
    return p;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    if (fn == nullptr) {        // resolved at run time: the library carries no link-time dependency on libcuda
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiled>(p);
    }
    return fn;
}

// fp32 [rows, cols] with `row_stride` elements between rows, boxes of [box_rows x 32], 128-byte swizzle, zero fill
bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t row_stride, int box_rows) {
    EncodeTiled enc = encode_fn();
    if (enc == nullptr) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)row_stride * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

size_t tarl_value_mlp_workspace_bytes(int32_t n_rows, int32_t n_nodes) {
    if (n_rows <= 0 || n_nodes <= 0) return 0;
    return make_plan(n_rows, n_nodes).bytes;
}

int tarl_value_mlp_forward(const float* occupancy, int64_t occ_row_stride, const float* time, int64_t time_stride,
                           int32_t n_rows, int32_t n_nodes, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int32_t weights_changed, void* workspace,
                           size_t workspace_bytes, float* out, float* save_z1, float* save_z2, void* stream) {
    if (n_rows < 0 || n_nodes <= 0) return TARL_E_BADARG;
    if (n_rows == 0) return TARL_OK;
    if (!occupancy || !time || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !out || !workspace) return TARL_E_BADARG;
    // TMA addressing: 16-byte aligned base and row pitch
    if ((reinterpret_cast<uintptr_t>(occupancy) & 15) != 0 || (occ_row_stride & 3) != 0 || occ_row_stride < n_nodes ||
        (reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
        return TARL_E_BADARG;
    const Plan p = make_plan(n_rows, n_nodes);
    if (workspace_bytes < p.bytes) return TARL_E_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* w_hi = reinterpret_cast<float*>(ws + p.off_wh);
    float* w_lo = reinterpret_cast<float*>(ws + p.off_wl);
    float* w_time = reinterpret_cast<float*>(ws + p.off_wt);
    float* partials = reinterpret_cast<float*>(ws + p.off_part);
    CUtensorMap map_a, map_wh, map_wl;
    if (!make_map(&map_a, occupancy, n_rows, n_nodes, occ_row_stride, BM) ||
        !make_map(&map_wh, w_hi, kHidden, p.Kp, p.Kp, BN) || !make_map(&map_wl, w_lo, kHidden, p.Kp, p.Kp, BN))
        return TARL_E_LAUNCH;
    static const bool smem_ok =                                          // once per process (thread-safe static)
        cudaFuncSetAttribute(k_value_mlp_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) == cudaSuccess;
    if (!smem_ok) return TARL_E_LAUNCH;
    const int64_t n_w = (int64_t)kHidden * p.Kp;
    if (weights_changed)
        k_value_mlp_split_w<<<(unsigned)((n_w + 255) / 256), 256, 0, s>>>(w1, n_nodes, p.Kp, w_hi, w_lo, w_time);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(p.tiles, p.slices); cfg.blockDim = dim3(kThreadsGemm);
        cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = s;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_value_mlp_gemm, map_a, map_wh, map_wl, (int)n_rows, (int)p.kb_total, (int)p.kb_per_slice,
                           partials);
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((n_rows + kTailWarps - 1) / kTailWarps); cfg.blockDim = dim3(kTailWarps * 32);
        cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_value_mlp_tail, (const float*)partials, (int)p.slices, (int)n_rows, time, time_stride,
                           (const float*)w_time, b1, w2, b2, w3, b3, out, save_z1, save_z2);
    }
    return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
}

int tarl_value_mlp_backward(const float* occupancy, int64_t occ_row_stride, const float* time, int64_t time_stride,
                            int32_t n_rows, int32_t n_nodes, const float* w2, const float* w3, const float* z1,
                            const float* z2, const float* grad_out, float* scratch, float* grad_w1, float* grad_b1,
                            float* grad_w2, float* grad_b2, float* grad_w3, float* grad_b3, void* stream) {
    if (n_rows < 0 || n_nodes <= 0) return TARL_E_BADARG;
    if (!grad_w1 || !grad_b1 || !grad_w2 || !grad_b2 || !grad_w3 || !grad_b3) return TARL_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_rows == 0) {
        cudaMemsetAsync(grad_w1, 0, sizeof(float) * (size_t)kHidden * ((size_t)n_nodes + 1), s);
        cudaMemsetAsync(grad_b1, 0, sizeof(float) * kHidden, s); cudaMemsetAsync(grad_w2, 0, sizeof(float) * kHidden * kHidden, s);
        cudaMemsetAsync(grad_b2, 0, sizeof(float) * kHidden, s); cudaMemsetAsync(grad_w3, 0, sizeof(float) * kHidden, s);
        cudaMemsetAsync(grad_b3, 0, sizeof(float), s);
        return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
    }
    if (!occupancy || !time || !w2 || !w3 || !z1 || !z2 || !grad_out || !scratch || occ_row_stride < n_nodes)
        return TARL_E_BADARG;
    float* g_z1 = scratch;                                   // [n_rows, 64]
    float* g_z2 = scratch + (size_t)n_rows * kHidden;        // [n_rows, 64]
    float* d_wtime = g_z2 + (size_t)n_rows * kHidden;        // [64]
    k_value_mlp_bwd_small<<<1, kBwdThreads, 0, s>>>(n_rows, z1, z2, grad_out, time, time_stride, w2, w3, g_z1, g_z2, d_wtime,
                                                    grad_b1, grad_w2, grad_b2, grad_w3, grad_b3);
    k_value_mlp_dw1<<<(n_nodes + kDwThreads - 1) / kDwThreads, kDwThreads, 0, s>>>(occupancy, occ_row_stride, n_rows, n_nodes,
                                                                                 g_z1, d_wtime, grad_w1);
    return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH;
}

}  // extern "C"

// edge_mlp_tc.cu — MPNNPolicyNet.edge_mlp forward on the 5th-generation tensor cores (sm_100a: tcgen05.mma, TMEM).
//
// Reference semantics: /root/reference/src/agents/mpnn_agent.py:38-44 (the module) and :227-231 (its use, commented out
// in the reference): logit[b,e] = L3(relu(L2(relu(L1([x[b,src e] ‖ x[b,dst e] ‖ edge_attr[b,e]]))))), 33 -> 64 -> 32 -> 1.
// Per (row, edge) pair that is 4.2 k multiply-adds on 132 bytes of gathered input: the one dense per-element MLP of the
// path. Both hidden layers run as tcgen05.mma.kind::tf32 over tiles of 128 pairs (one pair per TMEM lane):
//
//   warps 0-3 (thread r = pair r = TMEM lane r)                         warp 4, one thread
//   gather the pair's 34 inputs (+ the constant 1 that carries b1),
//   split hi/lo (3xTF32, as csrc/value_mlp.cu), tcgen05.st  A1 -> TMEM
//                                                                       D1[128 x 128] = A1_hi x [W1_hi; W1_lo]^T (N = 128)
//                                                                                     + A1_lo x W1_hi^T         (N = 64)
//   tcgen05.ld D1, fold hi.lo columns, ReLU, split, tcgen05.st A2 -> TMEM
//                                                                       D2[128 x 64]  = A2_hi x [W2_hi; W2_lo]^T (N = 64)
//                                                                                     + A2_lo x W2_hi^T         (N = 32)
//   tcgen05.ld D2, + b2, ReLU, the 32 -> 1 layer in registers, store
//
// The activations never touch shared memory: A operands are read by the MMA from TMEM, where the worker warps wrote
// them; the B operands (the two weight matrices, split and laid out K-major with the 128-byte swizzle by a small
// preparation kernel) sit in shared memory for the CTA's lifetime. TMEM: 256 columns per CTA (A1 and A2 share one
// 128-column region, D1 and D2 the other), so two CTAs per SM overlap one's gathers with the other's MMAs.
// Precision: a.w ~ a_hi.w_hi + a_hi.w_lo + a_lo.w_hi with rounded splits, fp32 accumulation over K = 40 / 64: ~1e-7
// relative (tests: 1e-5 against float64).
#include <cuda_runtime.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace {

constexpr int kX = 16;
constexpr int kTile = 128;                 // pairs per tile = TMEM lanes
constexpr int kThreads = 160;              // 4 worker warps + 1 MMA warp
constexpr int kK1 = 40, kH1 = 64, kH2 = 32;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColA = 0, kColD = 128;         // A1_hi [0,40) A1_lo [40,80) | A2_hi [0,64) A2_lo [64,128); D1 [128,256) / D2 [128,192)
// shared memory: B1 = two K-tiles of [128 rows x 32 floats] (W1'_hi rows 0-63, W1'_lo rows 64-127), B2 = two K-tiles of
// [64 rows x 32 floats] (W2_hi rows 0-31, W2_lo rows 32-63); then b2, w3, b3 and the barriers
constexpr uint32_t kB1Tile = 128 * 128, kB2Tile = 64 * 128;
constexpr uint32_t kOffB2 = 2 * kB1Tile, kOffVec = kOffB2 + 2 * kB2Tile, kOffBar = kOffVec + 68 * 4;
constexpr uint32_t kPrepFloats = kOffVec / 4 + 68;         // what the preparation kernel writes (floats)
// Requested shared memory: the image + barriers + alignment slack, padded to 100 KB so that at most TWO CTAs fit an SM —
// a third would sit in tcgen05.alloc (512 TMEM columns per SM, 256 per CTA) holding its share of the tiles until another
// CTA exits: measured 8.4 ms instead of 4.6 ms for 8 rows x 6.0 M edges.
constexpr size_t kSmemBytes = 100 * 1024;
static_assert(kOffBar + 64 + 1024 <= kSmemBytes, "weight image does not fit");

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, one 128 x N x 8 TF32 step
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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

// K-major operand tile in shared memory, rows of 128 bytes, SWIZZLE_128B: 8-row groups 1024 bytes apart (stride byte
// offset), descriptor version 1 (sm_100); tile base 1024-byte aligned; stepping 8 TF32 columns = +32 bytes.
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24); }
constexpr uint32_t kIdescL1a = idesc_tf32(2 * kH1), kIdescL1b = idesc_tf32(kH1), kIdescL2a = idesc_tf32(2 * kH2),
                   kIdescL2b = idesc_tf32(kH2);

constexpr uint32_t kHiMask = 0xFFFFE000u;
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & kHiMask); }

// float offset of element (row r, column c in [0, 32)) inside a swizzled K-major tile
__device__ __forceinline__ int sw128(int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }

// ------------------------------------------------------------------------------------------------ weight preparation
// prep (floats): B1 tile 0 / 1 (4096 each), B2 tile 0 / 1 (2048 each), then b2[32], w3[32], b3 — the image a CTA copies
// into its shared memory. W1' = [W1 (33 columns) | b1 | 0 ...] (40 columns: the pair's input vector carries a constant 1).
__global__ void __launch_bounds__(256) k_edge_mlp_tc_prep(const float* __restrict__ w1, const float* __restrict__ b1,
                                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                                          const float* __restrict__ w3, const float* __restrict__ b3,
                                                          float* __restrict__ prep) {
    for (int i = threadIdx.x; i < (int)kPrepFloats; i += 256) prep[i] = 0.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < kH1 * kK1; i += 256) {
        const int r = i / kK1, c = i - r * kK1;
        const float w = c < 33 ? w1[r * 33 + c] : (c == 33 ? b1[r] : 0.0f);
        const float hi = tf32_rn(w), lo = tf32_rn(w - hi);
        float* tile = prep + (c >> 5) * (kB1Tile / 4);
        tile[sw128(r, c & 31)] = hi;
        tile[sw128(kH1 + r, c & 31)] = lo;
    }
    for (int i = threadIdx.x; i < kH2 * kH1; i += 256) {
        const int r = i / kH1, c = i - r * kH1;
        const float w = w2[i];
        const float hi = tf32_rn(w), lo = tf32_rn(w - hi);
        float* tile = prep + kOffB2 / 4 + (c >> 5) * (kB2Tile / 4);
        tile[sw128(r, c & 31)] = hi;
        tile[sw128(kH2 + r, c & 31)] = lo;
    }
    float* vec = prep + kOffVec / 4;
    for (int i = threadIdx.x; i < kH2; i += 256) { vec[i] = b2[i]; vec[kH2 + i] = w3[i]; }
    if (threadIdx.x == 0) vec[2 * kH2] = b3[0];
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 2) k_edge_mlp_tc(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                             int E, int B, const float* __restrict__ x, int64_t x_bs,
                                                             const float* __restrict__ ea, int64_t ea_bs,
                                                             const float* __restrict__ prep, float* __restrict__ out,
                                                             int64_t out_bs, int64_t out_es) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + kOffBar;
    const uint32_t a1_ready = bars, d1_full = bars + 8, a2_ready = bars + 16, d2_full = bars + 24, tmem_slot = bars + 32;
    const float* vec = reinterpret_cast<const float*>(base_ptr + kOffVec);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {   // the weight image: 49 424 bytes, copied as 128-bit words
        const float4* g = reinterpret_cast<const float4*>(prep);
        float4* s = reinterpret_cast<float4*>(base_ptr);
        for (int i = threadIdx.x; i < (int)kPrepFloats / 4; i += kThreads) s[i] = g[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(a1_ready, kTile); mbar_init(a2_ready, kTile);
        mbar_init(d1_full, 1); mbar_init(d2_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the MMA reads the weight tiles through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_per_row = (E + kTile - 1) / kTile;
    const long long n_tiles = (long long)tiles_per_row * B;
    // Worker threads run their gathers ahead of the tile they feed to the tensor core: the edge's two node ids two tiles
    // ahead, the 34 inputs (two dependent load levels behind the ids) one tile ahead — requested right after this tile's
    // A1 went to TMEM, in flight while the thread waits for the MMAs and folds D1 / D2. Without it every tile began
    // with ~1.5 us of exposed gather latency and an SM had two tiles (one per CTA) to hide it with.
    struct Ids { int b, e, s, d; bool live; };
    auto load_ids = [&](long long tile) {
        Ids r = {0, 0, 0, 0, false};
        if (tile < n_tiles) {
            r.b = (int)(tile / tiles_per_row);
            r.e = (int)(tile - (long long)r.b * tiles_per_row) * kTile + (int)threadIdx.x;
            r.live = r.e < E;
            if (r.live) { r.s = src[r.e]; r.d = dst[r.e]; }
        }
        return r;
    };
    auto gather = [&](const Ids& id, float (&a)[kK1]) {      // A1 = [x_i | x_j | attr | 1 | 0 ...]
#pragma unroll
        for (int c = 0; c < kK1; ++c) a[c] = 0.0f;
        if (id.live) {
            const float4* xi = reinterpret_cast<const float4*>(x + id.b * x_bs + (int64_t)id.s * kX);
            const float4* xj = reinterpret_cast<const float4*>(x + id.b * x_bs + (int64_t)id.d * kX);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 u = xi[q], v = xj[q];
                a[4 * q] = u.x; a[4 * q + 1] = u.y; a[4 * q + 2] = u.z; a[4 * q + 3] = u.w;
                a[kX + 4 * q] = v.x; a[kX + 4 * q + 1] = v.y; a[kX + 4 * q + 2] = v.z; a[kX + 4 * q + 3] = v.w;
            }
            a[32] = ea[id.b * ea_bs + id.e];
            a[33] = 1.0f;
        }
    };
    Ids id_cur = {0, 0, 0, 0, false}, id_next = id_cur;
    float a[kK1];
    if (warp < 4) {
        id_cur = load_ids(blockIdx.x);
        id_next = load_ids((long long)blockIdx.x + gridDim.x);
        gather(id_cur, a);
    }
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        if (warp == 4) {
            if (lane == 0) {                                             // ===== MMA issuer
                mbar_wait(a1_ready, ph);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kK1 / 8; ++ks) {
                    const uint64_t bd = kmajor_sw128_desc(base + (ks >> 2) * kB1Tile) + 2 * (ks & 3);
                    tc_mma_tf32_ts(tmem_base + kColD, tmem_base + kColA + 8 * ks, bd, kIdescL1a, ks > 0 ? 1u : 0u);
                    tc_mma_tf32_ts(tmem_base + kColD, tmem_base + kColA + kK1 + 8 * ks, bd, kIdescL1b, 1u);
                }
                tc_commit(d1_full);
                mbar_wait(a2_ready, ph);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kH1 / 8; ++ks) {
                    const uint64_t bd = kmajor_sw128_desc(base + kOffB2 + (ks >> 2) * kB2Tile) + 2 * (ks & 3);
                    tc_mma_tf32_ts(tmem_base + kColD, tmem_base + kColA + 8 * ks, bd, kIdescL2a, ks > 0 ? 1u : 0u);
                    tc_mma_tf32_ts(tmem_base + kColD, tmem_base + kColA + kH1 + 8 * ks, bd, kIdescL2b, 1u);
                }
                tc_commit(d2_full);
            }
            continue;
        }
        // ===== worker warps: thread r <-> pair r of the tile <-> TMEM lane r
        const int b = id_cur.b, e = id_cur.e;
        const bool live = id_cur.live;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        {   // A1 split hi / lo, 40 + 40 columns
#pragma unroll
            for (int c8 = 0; c8 < kK1 / 8; ++c8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = tf32_rn(a[8 * c8 + j]);
                    hi[j] = __float_as_uint(h);
                    lo[j] = __float_as_uint(tf32_rn(a[8 * c8 + j] - h));
                }
                tc_st8(lane_addr + kColA + 8 * c8, hi);
                tc_st8(lane_addr + kColA + kK1 + 8 * c8, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(a1_ready);
        id_cur = id_next;                                                // the next tile's inputs: requested now
        id_next = load_ids(tile + 2 * (long long)gridDim.x);
        gather(id_cur, a);
        // h1 = relu(D1[:, 0:64] + D1[:, 64:128]) -> A2 hi / lo (64 + 64 columns, over the A1 columns: layer 1 is through)
        mbar_wait(d1_full, ph);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t p[32], q[32];
            tc_ld32(lane_addr + kColD + 32 * half, p);
            tc_ld32(lane_addr + kColD + kH1 + 32 * half, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h1 = fmaxf(__uint_as_float(p[8 * c8 + j]) + __uint_as_float(q[8 * c8 + j]), 0.0f);
                    const float h = tf32_rn(h1);
                    hi[j] = __float_as_uint(h);
                    lo[j] = __float_as_uint(tf32_rn(h1 - h));
                }
                tc_st8(lane_addr + kColA + 32 * half + 8 * c8, hi);
                tc_st8(lane_addr + kColA + kH1 + 32 * half + 8 * c8, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(a2_ready);
        // h2 = relu(D2[:, 0:32] + D2[:, 32:64] + b2), logit = w3 . h2 + b3
        mbar_wait(d2_full, ph);
        tc_fence_after();
        {
            uint32_t p[32], q[32];
            tc_ld32(lane_addr + kColD, p);
            tc_ld32(lane_addr + kColD + kH2, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float o = vec[2 * kH2];
#pragma unroll
            for (int j = 0; j < kH2; ++j)
                o += vec[kH2 + j] * fmaxf(__uint_as_float(p[j]) + __uint_as_float(q[j]) + vec[j], 0.0f);
            if (live) out[b * out_bs + e * out_es] = o;
        }
        tc_fence_before();         // the next tile's tcgen05.st / MMAs overwrite what was just read
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

}  // namespace

namespace tarl {

// weights: {W1 [64,33], b1, W2 [32,64], b2, W3 [1,32], b3}. scratch: tarl_edge_mlp_tc_scratch_floats() floats owned by
// the caller, 16-byte aligned: the weight image (49 KB: split, swizzled tiles) is rebuilt there on every call.
int edge_mlp_forward_tc(const int32_t* src, const int32_t* dst, int E, int B, const float* x, int64_t x_bs, const float* ea,
                        int64_t ea_bs, const float* const* weights, float* scratch, float* out, int64_t out_bs,
                        int64_t out_es, cudaStream_t s) {
    if (scratch == nullptr || (reinterpret_cast<uintptr_t>(scratch) & 15) != 0) return TARL_E_BADARG;
    static const bool ok =
        cudaFuncSetAttribute(k_edge_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) == cudaSuccess;
    if (!ok) return TARL_E_LAUNCH;
    k_edge_mlp_tc_prep<<<1, 256, 0, s>>>(weights[0], weights[1], weights[2], weights[3], weights[4], weights[5], scratch);
    const long long n_tiles = (long long)((E + kTile - 1) / kTile) * B;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cap = 2LL * sms;                       // two CTAs per SM (256 TMEM columns each), persistent over tiles
    k_edge_mlp_tc<<<(unsigned)(n_tiles < cap ? n_tiles : cap), kThreads, kSmemBytes, s>>>(src, dst, E, B, x, x_bs, ea, ea_bs,
                                                                                       scratch, out, out_bs, out_es);
    return launch_status();
}

}  // namespace tarl

extern "C" int32_t tarl_edge_mlp_tc_available(void) { return 1; }
extern "C" int32_t tarl_edge_mlp_tc_scratch_floats(void) { return (int32_t)kPrepFloats; }

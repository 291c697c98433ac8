// Microbenchmark: issue / completion rate of tcgen05.mma kind::tf32 on sm_100a for the operand forms the value MLP and
// the edge MLP use (A from TMEM or from shared memory, N = 64 / 128 / 256, K = 8 per instruction).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024u >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// MODE 0: ts N=128; 1: ts N=64; 2: ts N=256; 3: ss N=128; 4: ts N=128 then N=64 alternating (value MLP pattern);
// 6: ts bf16 N=128 (K=16); 7: ss N=256; 8: ts N=128 alternating two accumulators; 9: ts N=128 then N=64, commit per 8
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int n_mma, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 96 * 1024, slot = bar + 16;
    float* z = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
    for (int i = threadIdx.x; i < 96 * 256; i += 128) z[i] = 0.0f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(slot));
    if (threadIdx.x == 0) {
        const uint64_t db = desc_sw128(base), da = desc_sw128(base + 32 * 1024);
        const uint32_t acc = tm, a_t = tm + 256;
        uint32_t ph = 0;
        long long t0 = clock64(), t_issue = 0;
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (MODE == 0) { mma_ts(acc, a_t + 8 * k, db + 2 * k, idesc_tf32(128), 1); mma_ts(acc, a_t + 32 + 8 * k, db + 2 * k, idesc_tf32(128), 1); }
                if (MODE == 1) { mma_ts(acc, a_t + 8 * k, db + 2 * k, idesc_tf32(64), 1); mma_ts(acc, a_t + 32 + 8 * k, db + 2 * k, idesc_tf32(64), 1); }
                if (MODE == 2) { mma_ts(acc, a_t + 8 * k, db + 2 * k, idesc_tf32(256), 1); mma_ts(acc, a_t + 32 + 8 * k, db + 2 * k, idesc_tf32(256), 1); }
                if (MODE == 3) { mma_ss(acc, da + 2 * k, db + 2 * k, idesc_tf32(128), 1); mma_ss(acc, da + 512 + 2 * k, db + 2 * k, idesc_tf32(128), 1); }
                if (MODE == 4 || MODE == 9) { mma_ts(acc, a_t + 8 * k, db + 2 * k, idesc_tf32(128), 1); mma_ts(acc, a_t + 32 + 8 * k, db + 2 * k, idesc_tf32(64), 1); }
                if (MODE == 6) { mma_ts_f16(acc, a_t + 8 * k, db + 2 * k, idesc_bf16(128), 1); mma_ts_f16(acc, a_t + 32 + 8 * k, db + 2 * k, idesc_bf16(128), 1); }
                if (MODE == 7) { mma_ss(acc, da + 2 * k, db + 2 * k, idesc_tf32(256), 1); mma_ss(acc, da + 512 + 2 * k, db + 2 * k, idesc_tf32(256), 1); }
                if (MODE == 8) { mma_ts(acc, a_t + 8 * k, db + 2 * k, idesc_tf32(128), 1); mma_ts(acc + 128, a_t + 32 + 8 * k, db + 2 * k, idesc_tf32(128), 1); }
            }
            if (MODE == 9) { tc_commit(bar); }
        }
        t_issue = clock64() - t0;
        if (MODE == 9) { for (int i = 0; i < n_mma / 8; ++i) { mbar_wait(bar, ph); ph ^= 1; } }
        else { tc_commit(bar); mbar_wait(bar, 0); }
        long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t_issue; out[1] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
template <int MODE>
int run(const char* name, long long* out) {
    const int smem = 98 * 1024 + 1024, n = 2048;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148}) {
        k<MODE><<<grid, 128, smem>>>(n, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return 1; }
        printf("grid %3d  %-44s issue %.1f cyc/mma, complete %.1f cyc/mma\n", grid, name, (double)out[0] / n, (double)out[1] / n);
    }
    return 0;
}
int main() {
    long long* out; cudaMallocManaged(&out, 16);
    return run<0>("ts tf32 N=128", out) || run<1>("ts tf32 N=64", out) || run<2>("ts tf32 N=256", out) || run<3>("ss tf32 N=128", out) ||
           run<4>("ts tf32 N=128 / N=64 alternating", out) || run<6>("ts bf16 N=128 K=16", out) || run<7>("ss tf32 N=256", out) ||
           run<8>("ts tf32 N=128 two accumulators", out);
}

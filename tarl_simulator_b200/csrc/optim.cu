// optim.cu — the device side of the PPO update that is not a network (sm_100a): generalized advantage estimation over a
// [T, R] trajectory and the Adam step over one flat fp32 parameter bucket.
//
// Reference semantics (the reference takes both from libraries, restated here as the spec):
//   * torchrl 0.5.0 GAE as wired at /root/reference/src/rl/ppo_trainer.py:21-27 (gamma 0.99, lambda 0.95,
//     average_gae=True): delta_t = r_t + gamma V(s_{t+1}) (1 - terminated_t) - V(s_t);
//     A_t = delta_t + gamma lambda (1 - done_t) A_{t+1}; value_target = A + V. The standardisation of A over the batch
//     needs a reduction over every rank's frames: the kernel leaves per-CTA partial sums (fp64) in a fixed order.
//   * torch.optim.Adam(lr) as constructed at /root/reference/src/rl/ppo_trainer.py:39 (betas 0.9 / 0.999, eps 1e-8, no
//     weight decay, no amsgrad), in the operation order of torch's single-tensor path:
//     m = m + (g - m)(1 - b1); v = b2 v + (1 - b2) g g; denom = sqrt(v) / sqrt(1 - b2^k) + eps; p -= lr/(1 - b1^k) m/denom.
//     One launch over the flat bucket the gradient all-reduce uses, instead of ~6 launches per parameter tensor; the
//     global gradient norm the reference logs (:141, no clipping) comes out of the same pass.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tarl_b200.h"

namespace {

constexpr int kThreads = 128;

// one thread per replica walks its T steps backwards (coalesced over replicas at every t)
__global__ void __launch_bounds__(kThreads) k_gae(const float* __restrict__ value, const float* __restrict__ next_value,
                                                  int64_t v_stride, const float* __restrict__ reward,
                                                  const uint8_t* __restrict__ done, const uint8_t* __restrict__ terminated,
                                                  int T, int R, float gamma, float lmbda, float* __restrict__ adv,
                                                  float* __restrict__ target, double* __restrict__ partials) {
    __shared__ double sm[2][kThreads];
    const int r = blockIdx.x * kThreads + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (r < R) {
        float running = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            const int64_t i = (int64_t)t * R + r;
            const float v = value[(int64_t)t * v_stride + r], nv = next_value[(int64_t)t * v_stride + r];
            const float not_term = 1.0f - (terminated[i] ? 1.0f : 0.0f), not_done = 1.0f - (done[i] ? 1.0f : 0.0f);
            const float delta = reward[i] + gamma * nv * not_term - v;
            running = delta + gamma * lmbda * not_done * running;
            adv[i] = running;
            target[i] = running + v;
            s += (double)running;
            ss += (double)running * (double)running;
        }
    }
    sm[0][threadIdx.x] = s; sm[1][threadIdx.x] = ss;
    __syncthreads();
    for (int off = kThreads / 2; off > 0; off >>= 1) {            // fixed tree: deterministic
        if (threadIdx.x < off) { sm[0][threadIdx.x] += sm[0][threadIdx.x + off]; sm[1][threadIdx.x] += sm[1][threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = sm[0][0]; partials[2 * blockIdx.x + 1] = sm[1][0]; }
}

// stats = {n, sum A, sum A^2} after the all-reduce: A <- (A - mean) / max(std, 1e-4), unbiased std (average_gae=True)
__global__ void __launch_bounds__(256) k_standardise(float* __restrict__ adv, int64_t n, const double* __restrict__ stats) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double cnt = stats[0], mean = stats[1] / cnt;
    double var = (stats[2] - cnt * mean * mean) / (cnt > 1.0 ? cnt - 1.0 : 1.0);
    if (var < 0.0) var = 0.0;
    double sd = sqrt(var);
    if (sd < 1e-4) sd = 1e-4;
    adv[i] = (float)(((double)adv[i] - mean) / sd);
}

constexpr int kAdamThreads = 256, kAdamPerThread = 4;

__global__ void __launch_bounds__(kAdamThreads) k_adam(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n, float b1,
                                                       float b2, float eps, float step_size, float bc2_sqrt,
                                                       float grad_scale, double* __restrict__ norm_partials) {
    __shared__ double sm[kAdamThreads];
    const int64_t base = ((int64_t)blockIdx.x * kAdamThreads + threadIdx.x) * kAdamPerThread;
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < kAdamPerThread; ++k) {
        const int64_t i = base + k;
        if (i < n) {
            const float gi = g[i] * grad_scale;
            ss += (double)gi * (double)gi;
            const float mi = m[i] + (gi - m[i]) * (1.0f - b1);               // exp_avg.lerp_(grad, 1 - beta1)
            const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;              // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            m[i] = mi; v[i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] - step_size * (mi / denom);                         // addcdiv_(exp_avg, denom, value = -step_size)
        }
    }
    sm[threadIdx.x] = ss;
    __syncthreads();
    for (int off = kAdamThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && norm_partials != nullptr) norm_partials[blockIdx.x] = sm[0];
}

// One CTA: thread i sums partials i, i + 256, ... in that order, then a fixed tree over the 256 sums — the same result
// on every rank and every run. (A single thread walking the ~3 800 partials of the grid100 nets one dependent load after
// the other took 100 us, three times the Adam pass itself.)
__global__ void __launch_bounds__(kAdamThreads) k_norm_finish(const double* __restrict__ partials, int n,
                                                              float* __restrict__ out) {
    __shared__ double sm[kAdamThreads];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kAdamThreads) s += partials[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = kAdamThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)sqrt(sm[0]);
}

// torchrl 0.5.0 ClipPPOLoss with the reference's settings (/root/reference/src/rl/ppo_trainer.py:36,135-139: clip 0.2,
// entropy bonus 0.01, critic coefficient 1.0 with smooth-L1, advantages as given), forward AND the gradient of
// loss_objective + loss_critic + loss_entropy with respect to log_prob, entropy and value, for the n frames of one
// minibatch: one CTA, one launch, where the torch formula is ~40 launches of n-element kernels.
//   ratio = exp(lp - lp_old); objective = -mean(min(ratio A, clamp(ratio, 1 - c, 1 + c) A))
//   d objective / d lp = -(1/n) ratio A where the unclipped term is the smaller one (or both are the same term), else 0
//   critic = coef mean(smooth_l1(v - target)), d/dv = coef (|d| < 1 ? d : sign d) / n;  entropy loss = -c_e mean(H)
// out = {loss_objective, loss_entropy, loss_critic, approx_kl, clip_fraction, entropy, #impossible frames}. Sums in fp64 over a fixed
// thread-strided order and a fixed tree: deterministic.
constexpr int kLossThreads = 256;
__global__ void __launch_bounds__(kLossThreads) k_ppo_clip_loss(const float* __restrict__ lp, const float* __restrict__ lp_old,
                                                                const float* __restrict__ adv, const float* __restrict__ ent,
                                                                const float* __restrict__ val, const float* __restrict__ tgt,
                                                                int n, float lo, float hi, float clip, float ent_coef,
                                                                float critic_coef, float* __restrict__ out,
                                                                float* __restrict__ g_lp, float* __restrict__ g_ent,
                                                                float* __restrict__ g_val) {
    __shared__ double sm[6][kLossThreads];
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const float inv_n = 1.0f / (float)n;
    for (int i = threadIdx.x; i < n; i += kLossThreads) {
        // GraphDistribution.log_prob marks an action that selects no edge in some group as impossible: -inf
        // (/root/reference/src/reinforcement_learning.py:82-93), and its sample() produces such an action whenever a
        // group's uniform is not below the group's last cumulative probability (:57-80) — about one draw in 10^8. On the
        // reference's test networks that never happens; over the 2.4 10^8 draws of one 128-replica grid100 rollout it
        // happens a few times, the frame's ratio exp(-inf - -inf) is NaN and one optimiser step later so is every
        // parameter. Such a frame (BOTH log-probabilities -inf, nothing else) takes no part in the objective; it still
        // counts in the critic and entropy terms and in every mean's denominator; out[6] counts them.
        const bool impossible = (lp[i] == -INFINITY) && (lp_old[i] == -INFINITY);
        const float lw = impossible ? 0.0f : lp[i] - lp_old[i];
        const float ratio = expf(lw);
        const float a = impossible ? 0.0f : adv[i];
        const float clamped = ratio < lo ? lo : (ratio > hi ? hi : ratio);      // a NaN ratio stays NaN, as in torch.clamp
        const float g1 = ratio * a, g2 = clamped * a;
        const float gmin = (g1 != g1 || g2 != g2) ? g1 + g2 : fminf(g1, g2);    // torch.minimum propagates NaN, fminf drops it
        const bool inside = (ratio >= lo) && (ratio <= hi);         // clamp passes the gradient on its closed interval
        // torch.minimum splits the gradient evenly on ties; g2's own derivative is ratio A inside the interval, 0 outside
        float w = 0.0f;
        if (g1 < g2) w = 1.0f;
        else if (g1 == g2) w = inside ? 1.0f : 0.5f;
        else w = 0.0f;                                              // g2 < g1 happens only outside the interval
        g_lp[i] = -inv_n * (w * g1);
        g_ent[i] = -ent_coef * inv_n;
        const float d = val[i] - tgt[i], ad = fabsf(d);
        g_val[i] = critic_coef * inv_n * (ad < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f));
        acc[0] += (double)gmin;
        acc[1] += (double)ent[i];
        acc[2] += (double)(ad < 1.0f ? 0.5f * d * d : ad - 0.5f);
        acc[3] += (double)(-lw);
        acc[4] += (fabsf(ratio - 1.0f) > clip) ? 1.0 : 0.0;
        acc[5] += impossible ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) sm[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int off = kLossThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off)
#pragma unroll
            for (int k = 0; k < 6; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double m = 1.0 / (double)n;
        out[0] = (float)(-sm[0][0] * m);
        out[1] = (float)(-(double)ent_coef * sm[1][0] * m);
        out[2] = (float)((double)critic_coef * sm[2][0] * m);
        out[3] = (float)(sm[3][0] * m);
        out[4] = (float)(sm[4][0] * m);
        out[5] = (float)(sm[1][0] * m);
        out[6] = (float)sm[5][0];
    }
}

inline int status() { return cudaGetLastError() == cudaSuccess ? TARL_OK : TARL_E_LAUNCH; }

}  // namespace

extern "C" {

int32_t tarl_gae_partial_count(int32_t n_replicas) { return n_replicas <= 0 ? 0 : (n_replicas + kThreads - 1) / kThreads; }

int tarl_gae(const float* value, const float* next_value, int64_t value_step_stride, const float* reward,
             const uint8_t* done, const uint8_t* terminated, int32_t n_steps, int32_t n_replicas, float gamma, float lmbda,
             float* advantage, float* value_target, double* partials, void* stream) {
    if (n_steps < 0 || n_replicas < 0) return TARL_E_BADARG;
    if (n_steps == 0 || n_replicas == 0) return TARL_OK;
    if (!value || !next_value || !reward || !done || !terminated || !advantage || !value_target || !partials)
        return TARL_E_BADARG;
    k_gae<<<tarl_gae_partial_count(n_replicas), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        value, next_value, value_step_stride, reward, done, terminated, n_steps, n_replicas, gamma, lmbda, advantage,
        value_target, partials);
    return status();
}

int tarl_standardise(float* advantage, int64_t n, const double* stats, void* stream) {
    if (n < 0) return TARL_E_BADARG;
    if (n == 0) return TARL_OK;
    if (!advantage || !stats) return TARL_E_BADARG;
    k_standardise<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(advantage, n, stats);
    return status();
}

int32_t tarl_adam_partial_count(int64_t n) {
    const int64_t per_cta = (int64_t)kAdamThreads * kAdamPerThread;
    return n <= 0 ? 0 : (int32_t)((n + per_cta - 1) / per_cta);
}

int tarl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, int32_t step, float grad_scale, double* norm_partials, float* grad_norm,
                   void* stream) {
    if (n < 0 || step < 1) return TARL_E_BADARG;
    if (n == 0) return TARL_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || (grad_norm != nullptr && norm_partials == nullptr)) return TARL_E_BADARG;
    // bias corrections in double, as Python floats are in torch's single-tensor path
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const int ctas = tarl_adam_partial_count(n);
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    k_adam<<<ctas, kAdamThreads, 0, cs>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, (float)((double)lr / bc1),
                                          (float)sqrt(bc2), grad_scale, norm_partials);
    if (grad_norm != nullptr) k_norm_finish<<<1, kAdamThreads, 0, cs>>>(norm_partials, ctas, grad_norm);
    return status();
}

int tarl_ppo_clip_loss(const float* log_prob, const float* sample_log_prob, const float* advantage, const float* entropy,
                       const float* value, const float* value_target, int32_t n, double clip_epsilon, float entropy_coef,
                       float critic_coef, float* out, float* grad_log_prob, float* grad_entropy, float* grad_value,
                       void* stream) {
    if (n < 1) return TARL_E_BADARG;             // the mean over an empty minibatch is not a number
    if (!log_prob || !sample_log_prob || !advantage || !entropy || !value || !value_target || !out || !grad_log_prob ||
        !grad_entropy || !grad_value)
        return TARL_E_BADARG;
    // the interval bounds as torch forms them: Python doubles, rounded to fp32 when they meet the tensor
    const float lo = (float)(1.0 - clip_epsilon), hi = (float)(1.0 + clip_epsilon);
    k_ppo_clip_loss<<<1, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        log_prob, sample_log_prob, advantage, entropy, value, value_target, n, lo, hi, (float)clip_epsilon, entropy_coef,
        critic_coef,
        out, grad_log_prob, grad_entropy, grad_value);
    return status();
}

}  // extern "C"

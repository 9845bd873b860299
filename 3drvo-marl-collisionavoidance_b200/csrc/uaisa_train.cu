// uaisa_train.cu -- optimiser epilogue of the PPO update (SURVEY.md section 5 / 8a row a-P3) for sm_100a.
//
// The reference's update iteration ends with  clip_grad_norm_(ac.parameters(), 2.0); pi_optimizer.step()  preceded by a host-side
// "if kl > target_kl: break" (train/policy/multi_ppo.py:355-376).  Here all parameters, gradients and Adam moments live in FLAT
// fp32 buckets (one NCCL all-reduce per iteration covers the gradient + the KL scalar), and the tail of an iteration is two small
// kernels with NO host round trip:
//
//   uaisa_opt_prepare   fixed-partition partial sums of |g|^2 in fp64 (deterministic: every rank gets the same bits from the same reduced
//                       bucket) + ONE thread that evaluates the early-stop rule on the device: latches `stopped` once the (all-reduced) KL
//                       exceeds the target, otherwise bumps the optimiser's step counter and the count of applied iterations.
//   uaisa_clip_adam     total norm from the partials (fixed order) -> clip coefficient min(1, max_norm / (norm + 1e-6)) applied IN PLACE to
//                       the whole gradient bucket (clip_grad_norm_ scales every gradient, also the stale ones of the other head) -> Adam
//                       (torch.optim.Adam defaults: betas .9/.999, eps 1e-8, no weight decay, no amsgrad) on the optimiser's own range,
//                       skipped entirely when `stopped` is set.
//
// Both are asynchronous on the caller's stream and capturable in a CUDA graph (the whole multi-GPU iteration is one graph).
#include <cuda_runtime.h>
#include <stdint.h>

namespace uaisa {

constexpr int OPT_BLOCKS = 296;      // 2 x 148 SMs; also the length of the partial-sum array
constexpr int OPT_THREADS = 256;

// ctl[0] = stopped latch, ctl[1] = iterations applied since the caller cleared it
__global__ void __launch_bounds__(OPT_THREADS) uaisa_opt_prepare_kernel(const float *__restrict__ g, long long n, int want_norm, double *__restrict__ partials,
                                                                          const float *__restrict__ kl, float target_kl, int *__restrict__ ctl, int *__restrict__ step) {
    __shared__ double red[OPT_THREADS / 32];
    double acc = 0.0;
    if (want_norm) {
        // fixed partition: block b owns [b*chunk, (b+1)*chunk), thread t strides inside it -> the same summation order on every rank and run
        const long long chunk = (n + OPT_BLOCKS - 1) / OPT_BLOCKS, lo = (long long)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
        for (long long i = lo + threadIdx.x; i < hi; i += OPT_THREADS) { const double x = (double)g[i]; acc += x * x; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < OPT_THREADS / 32; w++) s += red[w];
            partials[blockIdx.x] = s;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int stopped = ctl ? ctl[0] : 0;
        if (ctl && kl && !(kl[0] <= target_kl)) stopped = 1;      // multi_ppo.py:362-364 (a NaN KL also stops)
        if (ctl) ctl[0] = stopped;
        if (!stopped) { step[0] += 1; if (ctl) ctl[1] += 1; }
    }
}

__global__ void __launch_bounds__(OPT_THREADS) uaisa_clip_adam_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                                                       long long n_total, long long r_off, long long r_cnt, const double *__restrict__ partials,
                                                                       float max_norm, float lr, float b1, float b2, float eps, const int *__restrict__ step,
                                                                       const int *__restrict__ ctl) {
    if (ctl && ctl[0]) return;                       // early stop: the iteration is not applied (uniform over the grid)
    __shared__ float s_coef;
    if (threadIdx.x == 0) {
        float coef = 1.0f;
        if (max_norm > 0.0f) {
            double s = 0.0;
            for (int b = 0; b < OPT_BLOCKS; b++) s += partials[b];       // fixed order
            const float total = (float)sqrt(s);
            coef = max_norm / (total + 1e-6f);                           // torch.nn.utils.clip_grad_norm_
            coef = coef < 1.0f ? coef : 1.0f;
        }
        s_coef = coef;
    }
    __syncthreads();
    const float coef = s_coef;
    const int t = step[0];
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride) {
        float gi = g[i];
        if (max_norm > 0.0f) { gi = gi * coef; g[i] = gi; }
        if (i >= r_off && i < r_off + r_cnt) {
            float mi = m[i], vi = v[i];
            mi = mi + (1.0f - b1) * (gi - mi);                           // exp_avg.lerp_(grad, 1 - beta1)
            vi = vi * b2 + (1.0f - b2) * gi * gi;                        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] - step_size * (mi / denom);                      // param.addcdiv_(exp_avg, denom, value=-step_size)
            m[i] = mi; v[i] = vi;
        }
    }
}

}  // namespace uaisa

extern "C" {

int uaisa_opt_partials_len(void) { return uaisa::OPT_BLOCKS; }

// g: the flat gradient bucket [n]; partials: double[uaisa_opt_partials_len()] scratch; kl: device scalar (the all-reduced approx-KL, NULL = no
// early-stop rule); ctl: int32[2] {stopped latch, applied count} (NULL = always apply); step: int32[1] device step counter of the optimiser.
int uaisa_opt_prepare(const float *g, int64_t n, int32_t want_norm, double *partials, const float *kl, float target_kl, int32_t *ctl, int32_t *step, void *stream) {
    if (!g || n < 1 || !partials || !step) return -1;
    uaisa::uaisa_opt_prepare_kernel<<<uaisa::OPT_BLOCKS, uaisa::OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(g, (long long)n, want_norm, partials, kl, target_kl, ctl, step);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// p/g/m/v: flat buckets [n_total]; Adam runs on [r_off, r_off + r_cnt); max_norm > 0: clip the WHOLE gradient bucket in place by the norm in `partials`.
int uaisa_clip_adam(float *p, float *g, float *m, float *v, int64_t n_total, int64_t r_off, int64_t r_cnt, const double *partials, float max_norm, float lr,
                    float beta1, float beta2, float eps, const int32_t *step, const int32_t *ctl, void *stream) {
    if (!p || !g || !m || !v || n_total < 1 || r_off < 0 || r_cnt < 0 || r_off + r_cnt > n_total || !step || (max_norm > 0.0f && !partials)) return -1;
    uaisa::uaisa_clip_adam_kernel<<<uaisa::OPT_BLOCKS, uaisa::OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, (long long)n_total, (long long)r_off, (long long)r_cnt,
                                                                                                                  partials, max_norm, lr, beta1, beta2, eps, step, ctl);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"

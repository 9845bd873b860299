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

// =====================================================================================================================================================
// Reader forward / backward for update minibatches whose samples all have <= 1 neighbour row (the captured-graph path of the PPO update): the
// bi-GRU from h0 = 0 (policy_rnn_ac.py:102-127 with L = 1: W_hh never enters), the sum of both directions, the concat with the 12 self values and
// LayerNorm(268) -- and their gradients -- in two kernels instead of ~60 element-wise / reduction launches over [minibatch x 768] tensors, which were
// 70 % of an update iteration.  fp32 on the CUDA cores (the update keeps the reference's precision: 1e-5 on the features against the torch ops, tests/test_ppo_gpu.py).
// Parameters are read in torch's own layouts (weight_ih [768][9] gate-major r|z|n, bias_ih / bias_hh [768]); gradients are ACCUMULATED (atomicAdd) into
// caller buffers of the same layouts.  `idx` (optional) gathers rows of `obs` on the fly, so the minibatch is never materialised.
namespace uaisa {

constexpr int RD_H = 256, RD_IN = 9, RD_SELF = 12, RD_FEAT = 268;

struct ReaderW {
    const float *w_ih[2], *b_ih[2], *b_hh[2], *ln_w, *ln_b;
};
struct ReaderG {
    float *w_ih[2], *b_ih[2], *b_hh[2], *ln_w, *ln_b;
};

// exp on the SFU (2 ulp) + fast division: absolute error ~1e-7 on sigmoid / tanh, far inside fp32 training noise; the IEEE divide and libm expf / tanhf
// forms were most of these kernels' instructions
__device__ __forceinline__ float rd_sig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float rd_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// Both kernels use the same mapping.  A CTA of 512 threads walks tiles of RB_T samples; thread = (direction d, hidden unit u) keeps ITS 27 input weights
// and 4 biases in registers, the tile's x rows (9 values each) are broadcast from shared memory:
//   gates phase   every thread evaluates r, z, n of its unit for each sample of the tile (K = 9) and leaves h = (1 - z) n in shared memory
//   row phase     one warp per sample: direction sum, concat with the 12 self values, LayerNorm statistics (forward: writes the features; backward: the
//                 LayerNorm backward -> dh[256], and dfeat * xhat for the LayerNorm weight gradient)
//   (backward)    gradient phase: thread = (d, u) again, gates recomputed, 27 weight + 4 bias gradients accumulated in registers over ALL the CTA's tiles;
//                 threads < 268 accumulate the LayerNorm weight / bias gradients; one atomicAdd per value per CTA at the end.
#ifndef UAISA_RB_T
#define UAISA_RB_T 32
#endif
constexpr int RB_T = UAISA_RB_T;   // samples per tile: three CTA barriers are paid per tile (15, what 48 KB of static shared memory held, measured 126 / 337 us
                                   // per 65 536 samples forward / backward); the arrays now live in dynamic shared memory (101 KB for the backward)
constexpr int RB_SMEM_FWD = (RB_T * (RD_IN + 3) + RB_T * 2 * RD_H) * 4, RB_SMEM_BWD = RB_SMEM_FWD + RB_T * RD_FEAT * 4;

struct ReaderThread {                 // thread (d, u): its weights, biases
    float wv[3][RD_IN], br, bz, bn, bhn;
    __device__ __forceinline__ void load(const ReaderW &W, int d, int u) {
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
            for (int k = 0; k < RD_IN; k++) wv[g][k] = __ldg(W.w_ih[d] + (size_t)(g * 256 + u) * RD_IN + k);
        br = __ldg(W.b_ih[d] + u) + __ldg(W.b_hh[d] + u); bz = __ldg(W.b_ih[d] + 256 + u) + __ldg(W.b_hh[d] + 256 + u);
        bn = __ldg(W.b_ih[d] + 512 + u); bhn = __ldg(W.b_hh[d] + 512 + u);
    }
    __device__ __forceinline__ void gates(const float *x, float &r, float &z, float &n) const {
        float ar = br, az = bz, an = bn;
#pragma unroll
        for (int k = 0; k < RD_IN; k++) { ar = fmaf(x[k], wv[0][k], ar); az = fmaf(x[k], wv[1][k], az); an = fmaf(x[k], wv[2][k], an); }
        r = rd_sig(ar); z = rd_sig(az); n = rd_tanh(an + r * bhn);
    }
};

__device__ __forceinline__ bool rd_all_zero(const float *x) {
    bool z = true;
#pragma unroll
    for (int k = 0; k < RD_IN; k++) z = z && (x[k] == 0.0f);
    return z;
}

// x rows of a tile -> shared memory (threads 0 .. ns*9-1), zero if the observation has no neighbour slot
__device__ __forceinline__ void load_x_tile(float (*s_x)[RD_IN + 3], const float *obs, const long long *idx, int obs_dim, int s0, int ns, int tid) {
    if (tid < ns * RD_IN) {
        const int j = tid / RD_IN, k = tid - j * RD_IN;
        const float *o = obs + (size_t)(idx ? idx[s0 + j] : s0 + j) * obs_dim;
        s_x[j][k] = (obs_dim >= RD_SELF + RD_IN) ? __ldg(o + RD_SELF + k) : 0.0f;
    }
}

__global__ void __launch_bounds__(512) uaisa_reader_fwd_kernel(const float *__restrict__ obs, const long long *__restrict__ idx, int B, int obs_dim, ReaderW W,
                                                                 float *__restrict__ feat, float *__restrict__ stats, float *__restrict__ hsum) {
    extern __shared__ __align__(16) float rd_smem[];
    float (*s_x)[RD_IN + 3] = reinterpret_cast<float (*)[RD_IN + 3]>(rd_smem);
    float (*s_h)[2][RD_H] = reinterpret_cast<float (*)[2][RD_H]>(rd_smem + RB_T * (RD_IN + 3));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, d = tid >> 8, u = tid & 255;
    ReaderThread T; T.load(W, d, u);
    // A sample without neighbour rows feeds a ZERO row (ir_gym.py:211-212): its gates are constants of the biases -- fmaf(0, w, b) = b exactly, so evaluating them
    // once per thread gives the very bits of the general path.  Sparse airspace: almost every sample of a configs[1] / configs[4] epoch.
    float r0, z0, n0;
    { float xz[RD_IN];
#pragma unroll
      for (int k = 0; k < RD_IN; k++) xz[k] = 0.0f;
      T.gates(xz, r0, z0, n0); }
    const float h0 = (1.0f - z0) * n0;
    const int n_tiles = (B + RB_T - 1) / RB_T;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s0 = tile * RB_T, ns = min(RB_T, B - s0);
        load_x_tile(s_x, obs, idx, obs_dim, s0, ns, tid);
        __syncthreads();
        for (int j = 0; j < ns; j++) {
            float x[RD_IN];
#pragma unroll
            for (int k = 0; k < RD_IN; k++) x[k] = s_x[j][k];
            if (rd_all_zero(x)) { s_h[j][d][u] = h0; continue; }       // (uniform over the CTA)
            float r, z, n; T.gates(x, r, z, n);
            s_h[j][d][u] = (1.0f - z) * n;
        }
        __syncthreads();
        for (int j = warp; j < ns; j += 16) {              // row phase: a warp per sample; lane l owns units l, l+32, ... and (l < 12) self value l
            const int s = s0 + j;
            const float *o = obs + (size_t)(idx ? idx[s] : s) * obs_dim;
            float h[8];
#pragma unroll
            for (int i = 0; i < 8; i++) h[i] = s_h[j][0][lane + 32 * i] + s_h[j][1][lane + 32 * i];
            if (hsum) {                                    // saved for the backward: its LayerNorm part then needs no second gate evaluation
#pragma unroll
                for (int i = 0; i < 8; i++) hsum[(size_t)s * RD_H + lane + 32 * i] = h[i];
            }
            const float sv = (lane < RD_SELF) ? __ldg(o + lane) : 0.0f;
            float sum = sv;
#pragma unroll
            for (int i = 0; i < 8; i++) sum += h[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            const float mean = sum * (1.0f / RD_FEAT);
            float ss = (lane < RD_SELF) ? (sv - mean) * (sv - mean) : 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) ss += (h[i] - mean) * (h[i] - mean);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            const float rstd = rsqrtf(ss * (1.0f / RD_FEAT) + 1e-5f);          // nn.LayerNorm: biased variance, eps 1e-5
            float *f = feat + (size_t)s * RD_FEAT;
            if (lane < RD_SELF) f[lane] = (sv - mean) * rstd * __ldg(W.ln_w + lane) + __ldg(W.ln_b + lane);
#pragma unroll
            for (int i = 0; i < 8; i++) { const int q = RD_SELF + lane + 32 * i; f[q] = (h[i] - mean) * rstd * __ldg(W.ln_w + q) + __ldg(W.ln_b + q); }
            if (lane == 0) { stats[2 * s] = mean; stats[2 * s + 1] = rstd; }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(512) uaisa_reader_bwd_kernel(const float *__restrict__ obs, const long long *__restrict__ idx, int B, int obs_dim, ReaderW W,
                                                                 const float *__restrict__ dfeat, const float *__restrict__ stats, const float *__restrict__ hsum, ReaderG G) {
    extern __shared__ __align__(16) float rd_smem[];
    float (*s_x)[RD_IN + 3] = reinterpret_cast<float (*)[RD_IN + 3]>(rd_smem);                               // x of the tile's samples
    float (*s_h)[2][RD_H] = reinterpret_cast<float (*)[2][RD_H]>(rd_smem + RB_T * (RD_IN + 3));              // gates phase: h per direction; row phase overwrites [j][0][.] with d loss / d (h_fwd + h_bwd)
    float (*s_gw)[RD_FEAT] = reinterpret_cast<float (*)[RD_FEAT]>(rd_smem + RB_T * (RD_IN + 3) + RB_T * 2 * RD_H);   // dfeat * xhat  (LayerNorm weight gradient terms)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, d = tid >> 8, u = tid & 255;
    ReaderThread T; T.load(W, d, u);
    float r0, z0, n0;                                      // gates of a zero row (see the forward kernel)
    { float xz[RD_IN];
#pragma unroll
      for (int k = 0; k < RD_IN; k++) xz[k] = 0.0f;
      T.gates(xz, r0, z0, n0); }
    float aw[3][RD_IN], ab[4], alw = 0.0f, alb = 0.0f;
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
        for (int k = 0; k < RD_IN; k++) aw[g][k] = 0.0f;
    ab[0] = ab[1] = ab[2] = ab[3] = 0.0f;
    const int n_tiles = (B + RB_T - 1) / RB_T;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s0 = tile * RB_T, ns = min(RB_T, B - s0);
        load_x_tile(s_x, obs, idx, obs_dim, s0, ns, tid);
        __syncthreads();
        if (!hsum) {                                       // gates phase (only when the forward did not save h_fwd + h_bwd)
            for (int j = 0; j < ns; j++) {
                float x[RD_IN];
#pragma unroll
                for (int k = 0; k < RD_IN; k++) x[k] = s_x[j][k];
                if (rd_all_zero(x)) { s_h[j][d][u] = (1.0f - z0) * n0; continue; }
                float r, z, n; T.gates(x, r, z, n);
                s_h[j][d][u] = (1.0f - z) * n;
            }
            __syncthreads();
        }
        for (int j = warp; j < ns; j += 16) {              // row phase: LayerNorm backward, a warp per sample
            const int s = s0 + j;
            const float *o = obs + (size_t)(idx ? idx[s] : s) * obs_dim;
            const float mean = stats[2 * s], rstd = stats[2 * s + 1];
            const float sv = (lane < RD_SELF) ? __ldg(o + lane) : 0.0f;
            const float *df = dfeat + (size_t)s * RD_FEAT;
            // g = dfeat * ln_w ; dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
            float gq[8], xh[8], sg = 0.0f, sgx = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int q = RD_SELF + lane + 32 * i;
                const float dfq = df[q], hq = hsum ? __ldg(hsum + (size_t)s * RD_H + lane + 32 * i) : s_h[j][0][lane + 32 * i] + s_h[j][1][lane + 32 * i];
                xh[i] = (hq - mean) * rstd; gq[i] = dfq * __ldg(W.ln_w + q);
                sg += gq[i]; sgx += gq[i] * xh[i];
                s_gw[j][q] = dfq * xh[i];
            }
            if (lane < RD_SELF) {
                const float dfq = df[lane], xs = (sv - mean) * rstd, gs = dfq * __ldg(W.ln_w + lane);
                sg += gs; sgx += gs * xs;
                s_gw[j][lane] = dfq * xs;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, off); sgx += __shfl_xor_sync(0xffffffffu, sgx, off); }
            const float mg = sg * (1.0f / RD_FEAT), mgx = sgx * (1.0f / RD_FEAT);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) s_h[j][0][lane + 32 * i] = rstd * (gq[i] - mg - xh[i] * mgx);      // dh (no gradient w.r.t. the observation is needed)
        }
        __syncthreads();
        for (int j = 0; j < ns; j++) {                     // gradient phase
            float x[RD_IN];
#pragma unroll
            for (int k = 0; k < RD_IN; k++) x[k] = s_x[j][k];
            const bool zero = rd_all_zero(x);                          // (uniform over the CTA) a zero row: constant gates, no weight gradient -- the same expressions, so the same bits
            float r = r0, z = z0, n = n0;
            if (!zero) T.gates(x, r, z, n);
            const float dh = s_h[j][0][u];
            const float dpn = dh * (1.0f - z) * (1.0f - n * n);        // d / d (i_n + b_in + r * b_hn)
            const float dpz = -dh * n * z * (1.0f - z);
            const float dpr = dpn * T.bhn * r * (1.0f - r);
            if (!zero) {
#pragma unroll
                for (int k = 0; k < RD_IN; k++) { aw[0][k] = fmaf(dpr, x[k], aw[0][k]); aw[1][k] = fmaf(dpz, x[k], aw[1][k]); aw[2][k] = fmaf(dpn, x[k], aw[2][k]); }
            }
            ab[0] += dpr; ab[1] += dpz; ab[2] += dpn; ab[3] += dpn * r;
            if (tid < RD_FEAT) { alw += s_gw[j][tid]; alb += dfeat[(size_t)(s0 + j) * RD_FEAT + tid]; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
        for (int k = 0; k < RD_IN; k++) atomicAdd(G.w_ih[d] + (size_t)(g * 256 + u) * RD_IN + k, aw[g][k]);
    atomicAdd(G.b_ih[d] + u, ab[0]); atomicAdd(G.b_hh[d] + u, ab[0]);
    atomicAdd(G.b_ih[d] + 256 + u, ab[1]); atomicAdd(G.b_hh[d] + 256 + u, ab[1]);
    atomicAdd(G.b_ih[d] + 512 + u, ab[2]); atomicAdd(G.b_hh[d] + 512 + u, ab[3]);
    if (tid < RD_FEAT) { atomicAdd(G.ln_w + tid, alw); atomicAdd(G.ln_b + tid, alb); }
}

// ---- PPO losses and their gradients with respect to the heads' outputs (multi_ppo.py:379-404), one thread per minibatch sample --------------------
// scal[0] <- max(1, number of usable samples): samples whose advantage / return is not finite are masked out (ok[i] == 0).  scal[1], scal[2] are an
// integer counter and an arrival ticket (zero between launches: the last CTA to arrive publishes the count and re-arms them).
__global__ void __launch_bounds__(256) uaisa_ppo_wsum_kernel(const uint8_t *__restrict__ ok, const long long *__restrict__ idx, int B, float *__restrict__ scal) {
    __shared__ int red[8];
    int c = 0;
    for (int s = blockIdx.x * 256 + threadIdx.x; s < B; s += gridDim.x * 256) c += ok[idx ? idx[s] : s] ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0; for (int w = 0; w < 8; w++) t += red[w];
        int *cnt = reinterpret_cast<int *>(scal + 1), *ticket = reinterpret_cast<int *>(scal + 2);
        atomicAdd(cnt, t);
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            const int tot = atomicAdd(cnt, 0);
            scal[0] = (float)(tot > 1 ? tot : 1);
            *cnt = 0; *ticket = 0;
        }
    }
}

__device__ __forceinline__ float nan_to_zero(float x) { return (x == x && fabsf(x) <= 3.402823466e38f) ? x : 0.0f; }   // torch.nan_to_num(x, 0, 0, 0)

// compute_loss_pi (multi_ppo.py:385-404): y [B][3] = the actor's pre-tanh output.  dy [B][3] = d(loss / W) / dy; dlogstd[3], slots[0] (approx-KL / W) and
// slots[1] (loss / W) are ADDED to (the caller zeroes them with the gradient bucket).  inv = 1 / (wsum * W).
__global__ void __launch_bounds__(256) uaisa_ppo_pi_loss_kernel(const float *__restrict__ y, const float *__restrict__ log_std, float std_factor,
                                                               const float *__restrict__ act, const float *__restrict__ logp_old, const float *__restrict__ adv,
                                                               const uint8_t *__restrict__ ok, const long long *__restrict__ idx, int B, float clip, float inv_world,
                                                               const float *__restrict__ scal, float *__restrict__ dy, float *__restrict__ dlogstd, float *__restrict__ slots) {
    __shared__ float red[8][5];
    const int s = blockIdx.x * 256 + threadIdx.x;
    float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};       // loss, kl, dlogstd[0..2]
    if (s < B) {
        const long long i = idx ? idx[s] : s;
        const float inv = inv_world / scal[0], w = ok[i] ? 1.0f : 0.0f;
        float mu[3], st[3], e[3], a[3], logp = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            mu[j] = tanhf(y[(size_t)s * 3 + j]);
            e[j] = std_factor * expf(__ldg(log_std + j)) + 1e-6f;
            st[j] = fminf(fmaxf(e[j], 1e-4f), 10.0f);
            a[j] = act[i * 3 + j];
            const float d = a[j] - mu[j];
            logp += -(d * d) / (2.0f * st[j] * st[j]) - logf(st[j]) - 0.9189385332046727f;
        }
        const float lo = logp_old[i], ratio = expf(logp - lo), av = nan_to_zero(adv[i]);
        const float ra = ratio * av, cl = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip) * av;
        part[0] = -fminf(ra, cl) * w;
        part[1] = (lo - logp) * w;
        const float dterm = -w * inv;
        const bool inside = ratio >= 1.0f - clip && ratio <= 1.0f + clip;
        const float g_ra = av * dterm, g_cl = inside ? av * dterm : 0.0f;           // d term / d ratio through either argument of min()
        const float dratio = ra < cl ? g_ra : (ra > cl ? g_cl : 0.5f * (g_ra + g_cl));   // torch.min splits the gradient at ties
        const float dlogp = ratio * dratio;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float d = a[j] - mu[j], is2 = 1.0f / (st[j] * st[j]);
            dy[(size_t)s * 3 + j] = dlogp * d * is2 * (1.0f - mu[j] * mu[j]);
            const float dstd = dlogp * (d * d * is2 / st[j] - 1.0f / st[j]);
            part[2 + j] = (e[j] >= 1e-4f && e[j] <= 10.0f) ? dstd * (e[j] - 1e-6f) : 0.0f;    // d std / d log_std = std_factor * exp(log_std) inside the clamp
        }
    }
#pragma unroll
    for (int q = 0; q < 5; q++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], off);
    }
    if ((threadIdx.x & 31) == 0) for (int q = 0; q < 5; q++) red[threadIdx.x >> 5][q] = part[q];
    __syncthreads();
    if (threadIdx.x < 5) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        const float inv = inv_world / scal[0];
        if (threadIdx.x == 0) atomicAdd(slots + 1, t * inv);            // loss / W
        else if (threadIdx.x == 1) atomicAdd(slots + 0, t * inv);       // approx-KL / W
        else atomicAdd(dlogstd + (threadIdx.x - 2), t);
    }
}

// compute_loss_v (multi_ppo.py:379-383): y [B] = the critic's output.  dy [B] = d(loss / W) / dy; loss_out[0] += sum((v - ret)^2 w) / wsum
__global__ void __launch_bounds__(256) uaisa_ppo_v_loss_kernel(const float *__restrict__ y, const float *__restrict__ ret, const uint8_t *__restrict__ ok,
                                                              const long long *__restrict__ idx, int B, float inv_world, const float *__restrict__ scal,
                                                              float *__restrict__ dy, float *__restrict__ loss_out) {
    __shared__ float red[8];
    const int s = blockIdx.x * 256 + threadIdx.x;
    float part = 0.f;
    if (s < B) {
        const long long i = idx ? idx[s] : s;
        const float w = ok[i] ? 1.0f : 0.0f, d = y[s] - nan_to_zero(ret[i]);
        part = d * d * w;
        dy[s] = 2.0f * d * w * (inv_world / scal[0]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; w++) t += red[w]; atomicAdd(loss_out, t / scal[0]); }
}

}  // namespace uaisa

extern "C" {

static int reader_attrs() {      // dynamic shared memory above 48 KB is an opt-in, per device
    static bool set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (dev < 0 || dev >= 64 || !set[dev]) {
        if (cudaFuncSetAttribute(uaisa::uaisa_reader_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, uaisa::RB_SMEM_FWD) != cudaSuccess) return -2;
        if (cudaFuncSetAttribute(uaisa::uaisa_reader_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, uaisa::RB_SMEM_BWD) != cudaSuccess) return -2;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    return 0;
}

// params: 8 DEVICE pointers weight_ih_l0, bias_ih_l0, bias_hh_l0, weight_ih_l0_reverse, bias_ih_l0_reverse, bias_hh_l0_reverse, ln.weight, ln.bias (torch layouts).
// obs [*, obs_dim]; idx (may be NULL): int64[B] row of `obs` each sample reads; feat [B][268]; stats [B][2] = LayerNorm mean, rstd (saved for the backward).
int uaisa_reader_fwd(const float *obs, const int64_t *idx, int32_t B, int32_t obs_dim, const float *const *params, float *feat, float *stats, float *hsum, void *stream) {
    if (!obs || !params || !feat || !stats || B < 1 || obs_dim < uaisa::RD_SELF) return -1;
    uaisa::ReaderW W{};
    for (int d = 0; d < 2; d++) { W.w_ih[d] = params[3 * d]; W.b_ih[d] = params[3 * d + 1]; W.b_hh[d] = params[3 * d + 2]; }
    W.ln_w = params[6]; W.ln_b = params[7];
    const int tiles_f = (B + uaisa::RB_T - 1) / uaisa::RB_T;
    if (reader_attrs()) return -2;
    uaisa::uaisa_reader_fwd_kernel<<<tiles_f < 592 ? tiles_f : 592, 512, uaisa::RB_SMEM_FWD, static_cast<cudaStream_t>(stream)>>>(obs, reinterpret_cast<const long long *>(idx), B, obs_dim, W, feat, stats, hsum);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// grads: 8 DEVICE pointers in the order of `params`; the gradients of this minibatch are ADDED to them.  dfeat [B][268] contiguous.
int uaisa_reader_bwd(const float *obs, const int64_t *idx, int32_t B, int32_t obs_dim, const float *const *params, const float *dfeat, const float *stats,
                     const float *hsum, float *const *grads, void *stream) {
    if (!obs || !params || !dfeat || !stats || !grads || B < 1 || obs_dim < uaisa::RD_SELF) return -1;
    uaisa::ReaderW W{}; uaisa::ReaderG G{};
    for (int d = 0; d < 2; d++) {
        W.w_ih[d] = params[3 * d]; W.b_ih[d] = params[3 * d + 1]; W.b_hh[d] = params[3 * d + 2];
        G.w_ih[d] = grads[3 * d]; G.b_ih[d] = grads[3 * d + 1]; G.b_hh[d] = grads[3 * d + 2];
    }
    W.ln_w = params[6]; W.ln_b = params[7]; G.ln_w = grads[6]; G.ln_b = grads[7];
    const int tiles = (B + uaisa::RB_T - 1) / uaisa::RB_T;
    const int grid = tiles < 444 ? tiles : 444;      // 3 CTAs of 512 threads per SM
    if (reader_attrs()) return -2;
    uaisa::uaisa_reader_bwd_kernel<<<grid, 512, uaisa::RB_SMEM_BWD, static_cast<cudaStream_t>(stream)>>>(obs, reinterpret_cast<const long long *>(idx), B, obs_dim, W, dfeat, stats, hsum, G);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// PPO losses of one update iteration and their gradients with respect to the heads' outputs (multi_ppo.py:379-404), minibatch gathered through idx
// (may be NULL).  scal: float[4] workspace, ZEROED once by the caller ([0] = the usable-sample count written by uaisa_ppo_wsum, [1..2] its counters).  See the kernels above for the outputs.
int uaisa_ppo_wsum(const uint8_t *ok, const int64_t *idx, int32_t B, float *scal, void *stream) {
    if (!ok || !scal || B < 1) return -1;
    const int grid = (B + 1023) / 1024 < 148 ? (B + 1023) / 1024 : 148;
    uaisa::uaisa_ppo_wsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ok, reinterpret_cast<const long long *>(idx), B, scal);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int uaisa_ppo_pi_loss(const float *y, const float *log_std, float std_factor, const float *act, const float *logp_old, const float *adv, const uint8_t *ok,
                      const int64_t *idx, int32_t B, float clip_ratio, float inv_world, const float *scal, float *dy, float *dlogstd, float *slots, void *stream) {
    if (!y || !log_std || !act || !logp_old || !adv || !ok || !scal || !dy || !dlogstd || !slots || B < 1) return -1;
    uaisa::uaisa_ppo_pi_loss_kernel<<<(B + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, log_std, std_factor, act, logp_old, adv, ok,
        reinterpret_cast<const long long *>(idx), B, clip_ratio, inv_world, scal, dy, dlogstd, slots);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int uaisa_ppo_v_loss(const float *y, const float *ret, const uint8_t *ok, const int64_t *idx, int32_t B, float inv_world, const float *scal, float *dy,
                     float *loss_out, void *stream) {
    if (!y || !ret || !ok || !scal || !dy || !loss_out || B < 1) return -1;
    uaisa::uaisa_ppo_v_loss_kernel<<<(B + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, ret, ok, reinterpret_cast<const long long *>(idx), B, inv_world,
                                                                                                  scal, dy, loss_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"

// uaisa_policy.cu -- fused rollout forward of the reference's actor-critic (train/policy/policy_rnn_ac.py) for sm_100a.
//
//   K6a  uaisa_policy_features_kernel   obs row -> bi-GRU step 0 (both directions, h0 = 0) -> sum -> concat 12 self
//                                       values -> LayerNorm(268): one warp per agent, fp32 on the CUDA cores.
//                                       Exact for every agent with <= 1 neighbour row (an agent without neighbours feeds
//                                       ONE zero row, ir_gym.py:211-212); the few agents with more rows are patched by
//                                       the host path (uaisa_ppo/policy.py).
//   K6b  uaisa_policy_heads_kernel      actor 268-256-256-3 (ReLU, ReLU, Tanh) and critic 268-256-256-1 for a tile of 128
//                                       agents per CTA on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with both
//                                       operands in shared memory (K-major, 128-byte swizzle), fp32 accumulators in TMEM,
//                                       tcgen05.ld epilogues that apply bias + activation and write the next layer's A operand
//                                       straight back into the swizzled shared-memory tile (activations never leave the SM).
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor).
#include <cuda_runtime.h>
#include <stdint.h>

namespace uaisa {

constexpr int PF_H = 256;        // GRU hidden size
constexpr int PF_IN = 9;         // GRU input size
constexpr int PF_SELF = 12;
constexpr int PF_FEAT = PF_SELF + PF_H;   // 268
constexpr int PF_KPAD = 288;     // feature row stride: 9 chunks of 32 tf32 (zero padded)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// wih_t: [2 dirs][9][768] (gate-major r|z|n inside 768); bias: [2][4][256] = b_ir+b_hr, b_iz+b_hz, b_in, b_hn
__global__ void __launch_bounds__(256) uaisa_policy_features_kernel(const float *__restrict__ obs, int A, int obs_dim, const float *__restrict__ wih_t,
                                                                      const float *__restrict__ bias, const float *__restrict__ ln_w,
                                                                      const float *__restrict__ ln_b, float *__restrict__ feat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= A) return;
    const float *o = obs + (size_t)warp * obs_dim;
    float x[PF_IN];
#pragma unroll
    for (int k = 0; k < PF_IN; k++) x[k] = (obs_dim >= PF_SELF + PF_IN) ? __ldg(o + PF_SELF + k) : 0.0f;
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int u = lane + 32 * i;
        float hs = 0.0f;
#pragma unroll
        for (int dir = 0; dir < 2; dir++) {
            const float *w = wih_t + (size_t)dir * PF_IN * 768;
            float ar = 0.f, az = 0.f, an = 0.f;
#pragma unroll
            for (int k = 0; k < PF_IN; k++) {
                ar = fmaf(x[k], __ldg(w + k * 768 + u), ar);
                az = fmaf(x[k], __ldg(w + k * 768 + 256 + u), az);
                an = fmaf(x[k], __ldg(w + k * 768 + 512 + u), an);
            }
            const float *b = bias + (size_t)dir * 4 * PF_H;
            const float r = sigmoidf_(ar + __ldg(b + u)), z = sigmoidf_(az + __ldg(b + PF_H + u));
            const float n = tanhf(an + __ldg(b + 2 * PF_H + u) + r * __ldg(b + 3 * PF_H + u));
            hs += (1.0f - z) * n;                      // h' = (1-z)*n + z*h0, h0 = 0
        }
        h[i] = hs;
    }
    const float s_self = (lane < PF_SELF) ? __ldg(o + lane) : 0.0f;
    float sum = s_self;
#pragma unroll
    for (int i = 0; i < 8; i++) sum += h[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float mean = sum * (1.0f / PF_FEAT);
    float ss = (lane < PF_SELF) ? (s_self - mean) * (s_self - mean) : 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) ss += (h[i] - mean) * (h[i] - mean);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float rstd = rsqrtf(ss * (1.0f / PF_FEAT) + 1e-5f);   // nn.LayerNorm: biased variance, eps 1e-5
    float *f = feat + (size_t)warp * PF_KPAD;
    if (lane < PF_SELF) f[lane] = (s_self - mean) * rstd * __ldg(ln_w + lane) + __ldg(ln_b + lane);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int q = PF_SELF + lane + 32 * i;
        f[q] = (h[i] - mean) * rstd * __ldg(ln_w + q) + __ldg(ln_b + q);
    }
    if (lane < PF_KPAD - PF_FEAT) f[PF_FEAT + lane] = 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------------
// K6b: tcgen05 heads
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HD_M = 128;                 // agents per CTA tile
constexpr int HD_N = 256;                 // hidden width
constexpr int HD_KC = 32;                 // tf32 elements per K chunk = one 128-byte swizzle row
constexpr int HD_A_CHUNK = HD_M * 128;    // bytes of one A chunk  [128 rows x 128 B]
constexpr int HD_B_TILE = HD_N * 128;     // bytes of one B chunk  [256 rows x 128 B]
constexpr int HD_NCH1 = PF_KPAD / HD_KC;  // 9 K-chunks for layer 1
constexpr int HD_NCH2 = HD_N / HD_KC;     // 8 K-chunks for layer 2
constexpr int HD_SMEM = 1024 + HD_NCH1 * HD_A_CHUNK + 2 * HD_B_TILE + 64;

struct HeadsParams {
    const float *feat; int A;
    const float *W1[2], *b1[2], *W2[2], *b2[2], *W3[2], *b3[2];   // [0] actor, [1] critic; W1 [256][288], W2 [256][256], W3 [nout][256]
    float *mu, *v;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_128B shared-memory matrix descriptor (UMMA::SmemDescriptor): start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 |
// version 1 <<46 | layout SWIZZLE_128B(2) <<61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// UMMA::InstrDescriptor for kind::tf32: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 <<17, M>>4 <<24
constexpr uint32_t HD_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HD_N >> 3) << 17) | ((uint32_t)(HD_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(HD_IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// element (row, 4 consecutive k starting at k4*4) of a [rows x 32] fp32 chunk in the K-major SWIZZLE_128B layout
__device__ __forceinline__ float4 *swz(char *chunk_base, int row, int k4) {
    return reinterpret_cast<float4 *>(chunk_base + (row >> 3) * 1024 + (row & 7) * 128 + ((k4 ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(128, 1) uaisa_policy_heads_kernel(const __grid_constant__ HeadsParams P) {
    extern __shared__ unsigned char hd_smem_raw[];
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(hd_smem_raw) + 1023) & ~uintptr_t(1023));   // SW128 needs 1024-byte alignment
    char *As = base;                                   // 9 chunks [128 x 32] fp32
    char *Bs = As + HD_NCH1 * HD_A_CHUNK;              // 2 buffers [256 x 32] fp32
    uint64_t *mbar = reinterpret_cast<uint64_t *>(Bs + 2 * HD_B_TILE);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a0 = blockIdx.x * HD_M;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // feature tile -> A chunks (row = thread)
    {
        const bool ok = a0 + tid < P.A;
        const float4 *src = reinterpret_cast<const float4 *>(P.feat + (size_t)(a0 + tid) * PF_KPAD);
        for (int c = 0; c < HD_NCH1; c++) {
#pragma unroll
            for (int k4 = 0; k4 < 8; k4++) *swz(As + c * HD_A_CHUNK, tid, k4) = ok ? __ldg(src + c * 8 + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    uint32_t phase[2] = {0u, 0u};
    int n_issued = 0;            // weight tiles issued so far (buffer = n_issued & 1)

    // one K-chunk GEMM step: W[rows 0..255][kc*32 .. +32) (row stride ldw floats) -> B buffer -> 4 MMAs into tmem_d
    auto gemm_chunk = [&](const float *W, int ldw, int kc, int a_chunk, uint32_t tmem_d, bool accumulate) {
        const int buf = n_issued & 1;
        if (n_issued >= 2) { mbar_wait(smem_u32(&mbar[buf]), phase[buf]); phase[buf] ^= 1u; }   // the MMAs that read this buffer are done
        char *Bb = Bs + buf * HD_B_TILE;
#pragma unroll 2
        for (int rr = 0; rr < 2; rr++) {
            const int n = tid + rr * 128;
            const float4 *src = reinterpret_cast<const float4 *>(W + (size_t)n * ldw + kc * HD_KC);
#pragma unroll
            for (int k4 = 0; k4 < 8; k4++) *swz(Bb, n, k4) = __ldg(src + k4);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(As + a_chunk * HD_A_CHUNK), sb = smem_u32(Bb);
#pragma unroll
            for (int k = 0; k < HD_KC / 8; k++)           // UMMA K = 8 for tf32: advance 32 bytes inside the swizzle row
                umma_tf32(tmem_d, umma_desc(sa + k * 32), umma_desc(sb + k * 32), (accumulate || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(&mbar[buf]));
        }
        n_issued++;
    };
    auto drain = [&]() {         // wait until every issued MMA has completed, then make TMEM readable by all threads
        const int last = (n_issued - 1) & 1;
        if (n_issued >= 2) { mbar_wait(smem_u32(&mbar[last ^ 1]), phase[last ^ 1]); phase[last ^ 1] ^= 1u; }
        mbar_wait(smem_u32(&mbar[last]), phase[last]); phase[last] ^= 1u;
        n_issued = 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    // ---- layer 1 of both heads from the same A (features): D0 = actor, D1 = critic
    for (int c = 0; c < HD_NCH1; c++) {
        gemm_chunk(P.W1[0], PF_KPAD, c, c, tmem + 0, c > 0);
        gemm_chunk(P.W1[1], PF_KPAD, c, c, tmem + HD_N, c > 0);
    }
    drain();
    for (int head = 0; head < 2; head++) {
        // epilogue of layer 1: bias + ReLU -> A chunks 0..7 (the features are no longer needed)
        const uint32_t d1 = tmem + (head ? HD_N : 0);
        for (int cc = 0; cc < HD_NCH2; cc++) {
            float v[32];
            tmem_ld32(d1 + lane_base + cc * 32, v);
#pragma unroll
            for (int k4 = 0; k4 < 8; k4++) {
                float4 o4;
                o4.x = fmaxf(v[4 * k4 + 0] + __ldg(P.b1[head] + cc * 32 + 4 * k4 + 0), 0.f);
                o4.y = fmaxf(v[4 * k4 + 1] + __ldg(P.b1[head] + cc * 32 + 4 * k4 + 1), 0.f);
                o4.z = fmaxf(v[4 * k4 + 2] + __ldg(P.b1[head] + cc * 32 + 4 * k4 + 2), 0.f);
                o4.w = fmaxf(v[4 * k4 + 3] + __ldg(P.b1[head] + cc * 32 + 4 * k4 + 3), 0.f);
                *swz(As + cc * HD_A_CHUNK, tid, k4) = o4;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // layer 2 -> D0 (the actor's layer-1 accumulator has been consumed; the critic's still sits in D1)
        for (int c = 0; c < HD_NCH2; c++) gemm_chunk(P.W2[head], HD_N, c, c, tmem + 0, c > 0);
        drain();
        // epilogue of layer 2: bias + ReLU, then the tiny output layer on the CUDA cores (row = thread)
        float acc[3] = {0.f, 0.f, 0.f};
        const int nout = head ? 1 : 3;
        for (int cc = 0; cc < HD_NCH2; cc++) {
            float v[32];
            tmem_ld32(tmem + lane_base + cc * 32, v);
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const float hq = fmaxf(v[q] + __ldg(P.b2[head] + cc * 32 + q), 0.f);
                acc[0] = fmaf(hq, __ldg(P.W3[head] + cc * 32 + q), acc[0]);
                if (!head) {
                    acc[1] = fmaf(hq, __ldg(P.W3[0] + HD_N + cc * 32 + q), acc[1]);
                    acc[2] = fmaf(hq, __ldg(P.W3[0] + 2 * HD_N + cc * 32 + q), acc[2]);
                }
            }
        }
        if (a0 + tid < P.A) {
            if (!head) {
#pragma unroll
                for (int q = 0; q < 3; q++) P.mu[(size_t)(a0 + tid) * 3 + q] = tanhf(acc[q] + __ldg(P.b3[0] + q));
            } else {
                P.v[a0 + tid] = acc[0] + __ldg(P.b3[1]);
            }
        }
        (void)nout;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();       // D0 fully read before the next head's layer 2 overwrites it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace uaisa

// ---- C ABI -----------------------------------------------------------------------------------------------------------
extern "C" {

int uaisa_policy_features(const float *obs, int32_t A, int32_t obs_dim, const float *wih_t, const float *bias, const float *ln_w,
                          const float *ln_b, float *feat, void *stream) {
    if (!obs || !wih_t || !bias || !ln_w || !ln_b || !feat || A < 1 || obs_dim < uaisa::PF_SELF) return -1;
    const int threads = 256, warps_per_block = threads / 32;
    uaisa::uaisa_policy_features_kernel<<<(A + warps_per_block - 1) / warps_per_block, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        obs, A, obs_dim, wih_t, bias, ln_w, ln_b, feat);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int uaisa_policy_heads(const float *feat, int32_t A, const float *const *weights /* 12 device pointers: W1a b1a W2a b2a W3a b3a W1c b1c W2c b2c W3c b3c */,
                       float *mu, float *v, void *stream) {
    if (!feat || !weights || !mu || !v || A < 1) return -1;
    static bool attr_set[64] = {};      // the attribute is per device (the launch runs on the device current in the calling thread)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(uaisa::uaisa_policy_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, uaisa::HD_SMEM) != cudaSuccess) return -2;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    uaisa::HeadsParams P{};
    P.feat = feat; P.A = A; P.mu = mu; P.v = v;
    for (int h = 0; h < 2; h++) {
        P.W1[h] = weights[6 * h + 0]; P.b1[h] = weights[6 * h + 1]; P.W2[h] = weights[6 * h + 2];
        P.b2[h] = weights[6 * h + 3]; P.W3[h] = weights[6 * h + 4]; P.b3[h] = weights[6 * h + 5];
    }
    uaisa::uaisa_policy_heads_kernel<<<(A + uaisa::HD_M - 1) / uaisa::HD_M, 128, uaisa::HD_SMEM, static_cast<cudaStream_t>(stream)>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"

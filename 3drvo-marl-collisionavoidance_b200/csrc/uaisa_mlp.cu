// uaisa_mlp.cu -- forward and backward of the actor / critic MLPs of the PPO update (train/policy/policy_rnn_ac.py:206-257 driven by
// multi_ppo.py:379-404) on the 5th-generation tensor cores of sm_100a.  Round 1/2a ran these GEMMs through torch autograd = cuBLAS SIMT fp32
// (60 % of an update iteration, profiles/r2/ppo_update_profile.txt).
//
// One kernel, uaisa_mlp_gemm_kernel<NT>, computes  C[M][N] (+)= A[M][K] * B[N][K]^T  with
//   * tcgen05.mma kind::tf32 (fp32 values, TF32 operand precision, fp32 accumulation in TMEM), CTA tile 128 x NT (NT = 256 or 288 columns:
//     one MMA of N = 256 plus, for the 268-wide feature side, one of N = 32), K in chunks of 32 floats = one 128-byte swizzle row;
//   * operands staged global -> registers -> shared memory (K-major, SWIZZLE_128B) by all 256 threads, the loads of chunk c+1 in flight while
//     chunk c is stored and multiplied; two stages, released by tcgen05.commit on an mbarrier;
//   * EITHER memory orientation of each operand: "K-major source" X[row][k] (16-byte loads along k) or "MN-major source" X[k][row] (the
//     transposing path: every thread gathers four k for one row with coalesced 4-byte loads and writes ONE 16-byte swizzled chunk -- consecutive
//     lanes own consecutive rows, so the 8 lanes of a quarter-warp hit 8 different chunk slots: conflict-free).  That is what lets the
//     backward pass run without a single transposed copy:  dW = dZ^T X  reads dZ and X as they lie (both MN-major sources, K = the batch),
//     dX = dZ W reads W[out][in] as an MN-major B operand;
//   * split-K over blockIdx.y with vector atomics (red.global.add.v4.f32) for the weight gradients;
//   * epilogues read the accumulator with tcgen05.ld (two warps per TMEM lane quarter split the columns):
//       EPI_STORE                      C = acc
//       EPI_BIAS_RELU                  C = relu(acc + bias[n])
//       EPI_BIAS_RELU_OUT              the same, plus the 256 -> nout output layer on the CUDA cores: Y[m][o] = sum_n C[m][n] W3[o][n] + b3[o]
//       EPI_MASK                       C = mask[m][n] > 0 ? acc : 0            (ReLU backward against the saved activation)
//       EPI_ATOMIC                     C += acc                                 (split-K)
// Small CUDA-core kernels around it: uaisa_mlp_out_bwd (output layer backward + ReLU mask + db2 + dW3 + db3, one thread per hidden unit) and
// uaisa_colsum (bias gradient of layer 1).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "uaisa_tc.cuh"

namespace uaisa {
using namespace tc;

enum GemmEpi : int { EPI_STORE = 0, EPI_BIAS_RELU = 1, EPI_BIAS_RELU_OUT = 2, EPI_MASK = 3, EPI_ATOMIC = 4 };

struct GemmParams {
    const float *A; int lda, a_mn;      // a_mn = 0: A[m][k] (k contiguous);  1: A[k][m] (m contiguous)
    const float *B; int ldb, b_mn;      // b_mn = 0: B[n][k];                 1: B[k][n]
    float *C; int ldc;
    int M, N, K;
    int k_per_split;                    // K range of one blockIdx.y, a multiple of 32
    int tile_m;                         // output rows one CTA owns (<= 128 = the MMA's M; the rows above it are zero operand rows nobody stores): chosen by the
                                        // launcher so that the row tiles fill whole waves of resident CTAs
    int epi;
    const float *bias;                  // [N]
    const float *mask; int ldm;         // [M][N]
    const float *W3, *b3; float *Y; int nout;   // EPI_BIAS_RELU_OUT: W3 [nout][N], b3 [nout], Y [M][nout]
#ifdef UAISA_GEMM_DBG
    int dbg;                            // timing experiments (results are WRONG): 1 = no operand loads after chunk 0, 2 = no MMAs, 4 = no epilogue loads / stores
#endif
};

constexpr int GM_M = 128;               // rows of the CTA tile = TMEM lanes
constexpr int GM_T = 256;               // threads
constexpr int GM_A_STAGE = GM_M * 128;  // bytes of one A chunk [128 rows x 32 floats]

template <int NT> constexpr int gemm_smem() { return 1024 + 2 * (GM_A_STAGE + NT * 128) + 128 * 4 * 4 + 64; }

// ---- operand chunk loader: [ROWS x 32 floats] of one K chunk into registers (zero beyond the matrix), then into the swizzled stage -----------------
// Everything that does not depend on the chunk -- the item's element offset in global memory, its swizzled byte offset in the stage, whether its row
// exists -- is computed ONCE per thread; a full chunk then costs one load and one store per item (the first version recomputed indices and bounds for
// every item of every chunk: 600 instructions per thread and chunk, the kernel was issue-bound on them).
template <int ROWS>
struct Loader {
    static constexpr int NI = ROWS * 8 / GM_T;
    uint32_t goff[NI], soff[NI], ok;
    int ld, mn, step;
    bool vec;
    const float *X;
    __device__ __forceinline__ void init(const float *X_, int ld_, int mn_, int row0, int R, int kb, bool vec_, int tid) {
        X = X_; ld = ld_; mn = mn_; vec = vec_; ok = 0u; step = mn_ ? 32 * ld_ : 32;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int item = tid + i * GM_T;
            const int row = mn_ ? item % ROWS : item >> 3, k4 = mn_ ? item / ROWS : item & 7, g = row0 + row;
            if (g < R) ok |= 1u << i;
            goff[i] = mn_ ? (uint32_t)(kb + 4 * k4) * (uint32_t)ld_ + (uint32_t)g : (uint32_t)g * (uint32_t)ld_ + (uint32_t)(kb + 4 * k4);
            soff[i] = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((k4 ^ (row & 7)) << 4));
        }
    }
    // chunk c of this CTA's K range; krem = K elements left from the chunk's first one (>= 32: a full chunk)
    __device__ __forceinline__ void load(float4 (&r)[NI], int c, int krem, int tid) const {
        const float *base = X + (size_t)c * step;
        if (krem >= 32) {
#pragma unroll
            for (int i = 0; i < NI; i++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((ok >> i) & 1u) {
                    const float *p = base + goff[i];
                    if (!mn) {
                        if (vec) v = __ldg(reinterpret_cast<const float4 *>(p));
                        else { v.x = __ldg(p); v.y = __ldg(p + 1); v.z = __ldg(p + 2); v.w = __ldg(p + 3); }
                    } else { v.x = __ldg(p); v.y = __ldg(p + ld); v.z = __ldg(p + 2 * ld); v.w = __ldg(p + 3 * ld); }
                }
                r[i] = v;
            }
        } else {                      // the ragged last chunk: per-element bounds
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const int item = tid + i * GM_T, k = (mn ? item / ROWS : item & 7) << 2, st = mn ? ld : 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((ok >> i) & 1u) {
                    const float *p = base + goff[i];
                    if (k < krem) v.x = __ldg(p);
                    if (k + 1 < krem) v.y = __ldg(p + st);
                    if (k + 2 < krem) v.z = __ldg(p + 2 * st);
                    if (k + 3 < krem) v.w = __ldg(p + 3 * st);
                }
                r[i] = v;
            }
        }
    }
    __device__ __forceinline__ void store(const float4 (&r)[NI], char *tile) const {
#pragma unroll
        for (int i = 0; i < NI; i++) *reinterpret_cast<float4 *>(tile + soff[i]) = r[i];
    }
};

template <int NT>
__global__ void __launch_bounds__(GM_T, NT <= 256 ? 2 : 1) uaisa_mlp_gemm_kernel(const __grid_constant__ GemmParams P) {
    constexpr int TM_COLS = NT <= 256 ? 256 : 512;        // TMEM allocation (a power of two): two CTAs per SM when NT <= 256
    constexpr int B_STAGE = NT * 128;
    extern __shared__ unsigned char gm_smem_raw[];
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~uintptr_t(1023));
    char *As = base, *Bs = base + 2 * GM_A_STAGE;
    float *s_y = reinterpret_cast<float *>(Bs + 2 * B_STAGE);                  // [128][4] partial output-layer sums of the upper column half
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_y + 128 * 4);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                   // (provably warp-uniform: the MMA issue below stays on the uniform datapath)
    const int m0 = blockIdx.x * P.tile_m, m_end = min(P.M, m0 + P.tile_m);
    const int kb = blockIdx.y * P.k_per_split, Kend = min(P.K, kb + P.k_per_split);
    const int nch = (Kend - kb + 31) >> 5;
    if (nch <= 0) return;                                 // (uniform) an empty split adds nothing
    const bool vecA = !P.a_mn && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0) && (P.lda & 3) == 0;
    const bool vecB = !P.b_mn && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0) && (P.ldb & 3) == 0;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
    if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_init_fence(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t empty0 = smem_u32(&bars[0]);
    constexpr uint32_t IDESC_MAIN = umma_idesc_tf32(GM_M, NT <= 256 ? NT : 256), IDESC_TAIL = umma_idesc_tf32(GM_M, NT > 256 ? NT - 256 : 8);

    float4 ra[GM_M * 8 / GM_T], rb[NT * 8 / GM_T];
    Loader<GM_M> la; la.init(P.A, P.lda, P.a_mn, m0, m_end, kb, vecA, tid);
    Loader<NT> lb; lb.init(P.B, P.ldb, P.b_mn, 0, P.N, kb, vecB, tid);
    la.load(ra, 0, Kend - kb, tid); lb.load(rb, 0, Kend - kb, tid);
    uint32_t par[2] = {0u, 0u};
    for (int c = 0; c < nch; c++) {
        const int s = c & 1;
        if (c >= 2) { mbar_wait(empty0 + 8 * s, par[s]); par[s] ^= 1u; }       // the MMAs that read this stage have completed
        la.store(ra, As + s * GM_A_STAGE);
        lb.store(rb, Bs + s * B_STAGE);
#ifdef UAISA_GEMM_DBG
        if (!(P.dbg & 1))
#endif
        if (c + 1 < nch) {                                                     // next chunk's loads fly while this one is multiplied
            la.load(ra, c + 1, Kend - kb - 32 * (c + 1), tid); lb.load(rb, c + 1, Kend - kb - 32 * (c + 1), tid);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {                                                       // warp-uniform descriptors; lane 0 issues (uaisa_policy_fwd.cu: the same change there was 14-19 %)
            fence_after_sync();
            const uint32_t sa = __shfl_sync(0xffffffffu, smem_u32(As + s * GM_A_STAGE), 0), sb = __shfl_sync(0xffffffffu, smem_u32(Bs + s * B_STAGE), 0);
            const uint64_t da = umma_desc(sa), db = umma_desc(sb), dt = umma_desc(sb + 256 * 128);
            if (lane == 0) {
#ifdef UAISA_GEMM_DBG
                if (!(P.dbg & 2))
#endif
#pragma unroll
                for (int k = 0; k < 4; k++) {                                  // UMMA K = 8 tf32 = 32 bytes inside the swizzle row: + 2 in the (address >> 4) field
                    umma_tf32(tmem, da + 2 * k, db + 2 * k, IDESC_MAIN, (c > 0 || k > 0) ? 1u : 0u);
                    if (NT > 256) umma_tf32(tmem + 256, da + 2 * k, dt + 2 * k, IDESC_TAIL, (c > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
            }
        }
    }
    {   // drain: every issued MMA has completed
        const int last = (nch - 1) & 1;
        if (nch >= 2) mbar_wait(empty0 + 8 * (last ^ 1), par[last ^ 1]);
        mbar_wait(empty0 + 8 * last, par[last]);
        fence_after_sync();
    }

    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31; warps w and w + 4 split the columns ---------------------------------------
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane, m = m0 + row;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool mok = m < m_end;
    const bool vecC = ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) && (P.ldc & 3) == 0;
    const bool vecM = P.mask && ((reinterpret_cast<uintptr_t>(P.mask) & 15) == 0) && (P.ldm & 3) == 0;
    float yacc[3] = {0.f, 0.f, 0.f};
    if (P.epi == EPI_ATOMIC) {                 // split-K: the partial tile is added straight from the registers (small outputs: the weight gradients)
        constexpr int HC = NT / 2;             // columns per half: 128 or 144 (multiples of 16)
#pragma unroll 1
        for (int cb = 0; cb < HC; cb += 16) {
            const int n0 = half * HC + cb;
            float v[16];
            __syncwarp();                      // tcgen05.ld is warp-collective: reconverge after the per-row branches below
            tmem_ld16(tmem + lane_base + n0, v);
            if (n0 >= P.N || !mok) continue;
            float *cp = P.C + (size_t)m * P.ldc + n0;
            if (vecC && n0 + 15 < P.N) {
#pragma unroll
                for (int i4 = 0; i4 < 4; i4++)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * i4), "f"(v[4 * i4]), "f"(v[4 * i4 + 1]), "f"(v[4 * i4 + 2]), "f"(v[4 * i4 + 3]) : "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) if (n0 + i < P.N) atomicAdd(cp + i, v[i]);
            }
        }
    } else {
        // The tile leaves through shared memory (the operand stages are free once the MMAs have drained), 128 columns per pass: a thread owns one accumulator
        // ROW, so direct stores put the 32 lanes of a warp on 32 different rows -- 16 useful bytes per 32-byte sector and request; timing experiments put the
        // epilogue's global traffic at 31 of 68 us (51 of 92 us with the mask loads, profiles/r2/gemm_dbg.txt).  Pass: TMEM -> registers -> (bias, ReLU, output
        // layer partials) -> staging tile [128][132]; then each warp streams 16 rows out, 512 contiguous bytes per instruction, the ReLU mask read the same way.
        constexpr int PITCH = 132;             // floats: consecutive rows start 4 banks apart -> conflict-free float4 column writes and row reads
        float *stg = reinterpret_cast<float *>(base);
        static_assert(GM_M * PITCH * 4 <= 2 * (GM_A_STAGE + NT * 128), "the staging tile reuses the operand stages");
#pragma unroll 1
        for (int pass = 0; pass * 128 < NT; pass++) {
            const int cbase = pass * 128;
#pragma unroll 1
            for (int b = 0; b < 4; b++) {
                const int n0 = cbase + half * 64 + b * 16;         // (warp-uniform)
                if (n0 >= NT) break;
                float v[16];
                tmem_ld16(tmem + lane_base + n0, v);
                if (n0 >= P.N) continue;
                if (P.epi == EPI_BIAS_RELU || P.epi == EPI_BIAS_RELU_OUT) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = (n0 + i < P.N) ? fmaxf(v[i] + __ldg(P.bias + n0 + i), 0.f) : 0.f;
                    if (P.epi == EPI_BIAS_RELU_OUT) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            if (n0 + i < P.N) {
#pragma unroll
                                for (int o = 0; o < 3; o++) if (o < P.nout) yacc[o] = fmaf(v[i], __ldg(P.W3 + (size_t)o * P.N + n0 + i), yacc[o]);
                            }
                        }
                    }
                }
                float4 *sp = reinterpret_cast<float4 *>(stg + row * PITCH + (n0 - cbase));
#pragma unroll
                for (int i4 = 0; i4 < 4; i4++) sp[i4] = make_float4(v[4 * i4], v[4 * i4 + 1], v[4 * i4 + 2], v[4 * i4 + 3]);
            }
            __syncthreads();
#ifdef UAISA_GEMM_DBG
            if (!(P.dbg & 4))
#endif
            if (P.C) {
                const int n = cbase + 4 * lane;
                const bool full = n + 3 < P.N, mv = P.epi == EPI_MASK && vecM && full;
                constexpr int RG = 8;                              // rows per group: the group's mask loads are all in flight before the first is used
#pragma unroll 1
                for (int r0 = warp; r0 < GM_M; r0 += RG * (GM_T / 32)) {
                    float4 h[RG];
#pragma unroll
                    for (int u = 0; u < RG; u++) {
                        const int mr = m0 + r0 + u * (GM_T / 32);
                        if (mv && mr < m_end) h[u] = __ldg(reinterpret_cast<const float4 *>(P.mask + (size_t)mr * P.ldm + n));
                    }
#pragma unroll
                    for (int u = 0; u < RG; u++) {
                        const int r = r0 + u * (GM_T / 32), mr = m0 + r;
                        if (mr >= m_end || n >= P.N) continue;
                        float4 o = *reinterpret_cast<const float4 *>(stg + r * PITCH + 4 * lane);
                        float *cp = P.C + (size_t)mr * P.ldc + n;
                        if (P.epi == EPI_MASK) {
                            if (mv) {
                                o.x = h[u].x > 0.f ? o.x : 0.f; o.y = h[u].y > 0.f ? o.y : 0.f; o.z = h[u].z > 0.f ? o.z : 0.f; o.w = h[u].w > 0.f ? o.w : 0.f;
                            } else {
                                const float *mp = P.mask + (size_t)mr * P.ldm + n;
                                o.x = __ldg(mp) > 0.f ? o.x : 0.f;
                                if (n + 1 < P.N) o.y = __ldg(mp + 1) > 0.f ? o.y : 0.f;
                                if (n + 2 < P.N) o.z = __ldg(mp + 2) > 0.f ? o.z : 0.f;
                                if (n + 3 < P.N) o.w = __ldg(mp + 3) > 0.f ? o.w : 0.f;
                            }
                        }
                        if (vecC && full) *reinterpret_cast<float4 *>(cp) = o;
                        else { cp[0] = o.x; if (n + 1 < P.N) cp[1] = o.y; if (n + 2 < P.N) cp[2] = o.z; if (n + 3 < P.N) cp[3] = o.w; }
                    }
                }
            }
            __syncthreads();                   // the staging tile is rewritten by the next pass
        }
    }
    if (P.epi == EPI_BIAS_RELU_OUT) {          // the two column halves of a row meet in shared memory
        if (half == 1) { s_y[row * 4 + 0] = yacc[0]; s_y[row * 4 + 1] = yacc[1]; s_y[row * 4 + 2] = yacc[2]; }
        __syncthreads();
        if (half == 0 && mok) {
#pragma unroll
            for (int o = 0; o < 3; o++) if (o < P.nout) P.Y[(size_t)m * P.nout + o] = yacc[o] + s_y[row * 4 + o] + __ldg(P.b3 + o);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}
template __global__ void uaisa_mlp_gemm_kernel<256>(const __grid_constant__ GemmParams);
template __global__ void uaisa_mlp_gemm_kernel<288>(const __grid_constant__ GemmParams);

// Output layer backward + ReLU backward of hidden layer 2, one thread per hidden unit j, one CTA per 128 samples:
//   dZ2[b][j] = H2[b][j] > 0 ? sum_o dy[b][o] W3[o][j] : 0;   db2[j] += sum_b dZ2[b][j];   dW3[o][j] += sum_b dy[b][o] H2[b][j];   db3[o] += sum_b dy[b][o]
// (the sums over the batch stay in the thread's registers: no cross-thread reduction; one atomic per value per CTA).
__global__ void __launch_bounds__(256) uaisa_mlp_out_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ H2, const float *__restrict__ W3, int Bn, int nout,
                                                               float *__restrict__ dZ2, float *__restrict__ dW3, float *__restrict__ db3, float *__restrict__ db2) {
    __shared__ float s_dy[128 * 3];
    const int j = threadIdx.x, b0 = blockIdx.x * 128, nb = min(128, Bn - b0);
    for (int i = j; i < nb * nout; i += 256) s_dy[i] = __ldg(dy + (size_t)b0 * nout + i);
    __syncthreads();
    float w3[3] = {0.f, 0.f, 0.f}, aw[3] = {0.f, 0.f, 0.f}, ab = 0.f;
    for (int o = 0; o < nout; o++) w3[o] = __ldg(W3 + o * 256 + j);
#pragma unroll 4
    for (int r = 0; r < nb; r++) {
        const float h = __ldg(H2 + (size_t)(b0 + r) * 256 + j);
        float dh = 0.f;
#pragma unroll
        for (int o = 0; o < 3; o++) if (o < nout) { const float g = s_dy[r * nout + o]; dh = fmaf(g, w3[o], dh); aw[o] = fmaf(g, h, aw[o]); }
        const float dz = h > 0.f ? dh : 0.f;
        dZ2[(size_t)(b0 + r) * 256 + j] = dz;
        ab += dz;
    }
    atomicAdd(db2 + j, ab);
    for (int o = 0; o < nout; o++) atomicAdd(dW3 + o * 256 + j, aw[o]);
    if (j < nout) { float s = 0.f; for (int r = 0; r < nb; r++) s += s_dy[r * nout + j]; atomicAdd(db3 + j, s); }
}

// out[j] += sum over rows of X[r][j]   (bias gradient of hidden layer 1); one thread per column, rows split over the grid
__global__ void __launch_bounds__(256) uaisa_colsum_kernel(const float *__restrict__ X, int rows, int cols, int rows_per_cta, float *__restrict__ out) {
    const int j = blockIdx.y * 256 + threadIdx.x, r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    if (j >= cols) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += __ldg(X + (size_t)r * cols + j); s1 += __ldg(X + (size_t)(r + 1) * cols + j);
        s2 += __ldg(X + (size_t)(r + 2) * cols + j); s3 += __ldg(X + (size_t)(r + 3) * cols + j);
    }
    for (; r < r1; r++) s0 += __ldg(X + (size_t)r * cols + j);
    atomicAdd(out + j, (s0 + s1) + (s2 + s3));
}

static int set_attr_once() {
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(uaisa_mlp_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<256>()) != cudaSuccess) return -2;
        if (cudaFuncSetAttribute(uaisa_mlp_gemm_kernel<288>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem<288>()) != cudaSuccess) return -2;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    return 0;
}

}  // namespace uaisa

extern "C" {

// C[M][N] (+)= A * B^T on tcgen05 (TF32 operands, fp32 accumulation); see the head of this file.  N <= 288.  splits > 1 requires epi = 4 (atomic
// accumulation into a C the caller has initialised).  Returns 0, -1 (bad arguments) or -2 (launch failure).
int uaisa_mlp_gemm(const float *A, int32_t lda, int32_t a_mn, const float *B, int32_t ldb, int32_t b_mn, float *C, int32_t ldc, int32_t M, int32_t N, int32_t K,
                    int32_t splits, int32_t epi, const float *bias, const float *mask, int32_t ldm, const float *W3, const float *b3, float *Y, int32_t nout,
                    void *stream) {
    using namespace uaisa;
    if (!A || !B || M < 1 || N < 1 || N > 288 || K < 1 || splits < 1 || epi < 0 || epi > 4) return -1;
    if (splits > 1 && epi != EPI_ATOMIC) return -1;
    if ((epi == EPI_BIAS_RELU || epi == EPI_BIAS_RELU_OUT) && !bias) return -1;
    if (epi == EPI_BIAS_RELU_OUT && (!W3 || !b3 || !Y || nout < 1 || nout > 3)) return -1;
    if (epi == EPI_MASK && !mask) return -1;
    if (epi != EPI_BIAS_RELU_OUT && !C) return -1;
    if (set_attr_once()) return -2;
    GemmParams P{};
    P.A = A; P.lda = lda; P.a_mn = a_mn; P.B = B; P.ldb = ldb; P.b_mn = b_mn; P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K;
    P.k_per_split = (((K + splits - 1) / splits) + 31) / 32 * 32;
#ifdef UAISA_GEMM_DBG
    P.dbg = std::getenv("UAISA_GEMM_DBG") ? std::atoi(std::getenv("UAISA_GEMM_DBG")) : 0;
#endif
    P.epi = epi; P.bias = bias; P.mask = mask; P.ldm = ldm; P.W3 = W3; P.b3 = b3; P.Y = Y; P.nout = nout;
    // Wave quantisation: 65 536 rows are 512 tiles of 128 on 296 resident CTAs (2 per SM; 148 for the 288-column kernel) = 1.73 waves that cost 2.  The kernel is
    // bound by its loads and stores, not by the MMAs, so a tile may own FEWER than the MMA's 128 rows: pick the height (multiple of 8, >= 64) that minimises
    // waves x height -- 112 rows -> 586 tiles = 1.98 waves of 0.875 the work each.
    int tile_m = GM_M;
    if (splits == 1) {
        int dev = 0, sms = 148;
        static int sm_count[64] = {};
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
            if (!sm_count[dev]) { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sm_count[dev] = v; }
            if (sm_count[dev]) sms = sm_count[dev];
        }
        const long long slots = (long long)sms * (N <= 256 ? 2 : 1);
        long long best = -1;
        for (int t = GM_M; t >= 64; t -= 8) {
            const long long tiles = (M + t - 1) / t, waves = (tiles + slots - 1) / slots, cost = waves * t;
            if (best < 0 || cost < best) { best = cost; tile_m = t; }
        }
    }
    P.tile_m = tile_m;
    const dim3 grid((M + tile_m - 1) / tile_m, (K + P.k_per_split - 1) / P.k_per_split);
    if (N <= 256) uaisa_mlp_gemm_kernel<256><<<grid, GM_T, gemm_smem<256>(), static_cast<cudaStream_t>(stream)>>>(P);
    else uaisa_mlp_gemm_kernel<288><<<grid, GM_T, gemm_smem<288>(), static_cast<cudaStream_t>(stream)>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// Output-layer backward of a 268-256-256-nout MLP (nout <= 3): dZ2 [B][256] is written, db2 [256], dW3 [nout][256] and db3 [nout] are ADDED to.
int uaisa_mlp_out_bwd(const float *dy, const float *H2, const float *W3, int32_t B, int32_t nout, float *dZ2, float *dW3, float *db3, float *db2, void *stream) {
    if (!dy || !H2 || !W3 || !dZ2 || !dW3 || !db3 || !db2 || B < 1 || nout < 1 || nout > 3) return -1;
    uaisa::uaisa_mlp_out_bwd_kernel<<<(B + 127) / 128, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, H2, W3, B, nout, dZ2, dW3, db3, db2);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// out[cols] += column sums of X [rows][cols]
int uaisa_colsum(const float *X, int32_t rows, int32_t cols, float *out, void *stream) {
    if (!X || !out || rows < 1 || cols < 1) return -1;
    const int rpc = 256;
    uaisa::uaisa_colsum_kernel<<<dim3((rows + rpc - 1) / rpc, (cols + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, rows, cols, rpc, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"

// uaisa_policy_fwd.cu -- K6: the COMPLETE rollout forward of the reference's actor-critic (train/policy/policy_rnn_ac.py:57-69) in one kernel
// for sm_100a:  obs row -> masked bi-GRU over the agent's neighbour rows (policy_rnn_ac.py:102-168: L = max(count, 1) steps per direction, final
// hidden states of both directions SUMMED) -> concat the 12 self values -> LayerNorm(268) -> actor 268-256-256-3 (ReLU, ReLU, Tanh) and critic
// 268-256-256-1.  One CTA = 128 agents (row = thread).  Nothing but the inputs and mu / v touches HBM: the GRU state, the features and the hidden
// activations live in shared memory (as the next GEMM's A operand) and in TMEM.
//
//   * every GEMM runs on the 5th-generation tensor cores: tcgen05.mma kind::tf32, A (activations) and B (weights) K-major with 128-byte swizzle in
//     shared memory, fp32 accumulators in TMEM.  GRU step: for a slice of 64 hidden units the accumulator holds [ r | z | i_n | h_n ] = 256 columns
//     (x_t W_i^T and h W_h^T accumulate into the same r / z columns); the epilogue (tcgen05.ld) applies the gates and stores the NEW hidden state into
//     a second TMEM region (tcgen05.st) because the old state is still the A operand of the other slices; after the 4 slices the new state is copied
//     back into the swizzled A tile.  Steps beyond an agent's own row count leave its state untouched (mask), a CTA runs max(count) steps.
//   * weights are PRE-SWIZZLED on the device into the exact shared-memory image of each [rows x 32] tile (uaisa_policy_pack), so a weight tile is ONE
//     bulk copy (cp.async.bulk global -> shared, the TMA engine; SASS UBLKCP) completing on an mbarrier.  Warp 0 is producer and MMA issuer -- the whole
//     warp runs the ring / issue code with warp-uniform state (uniform datapath), lane 0 executes the instructions with side effects: a 2-stage ring
//     (full / empty mbarriers, tcgen05.commit releases a stage) keeps the next two tiles in flight, also across the epilogues.
//   * a tile without any neighbour row takes the reader's constant output h0 (packed behind the weights by uaisa_policy_pack) and skips the GRU.
//   * 512 threads: 4 threads share an agent row (same TMEM lane quarter, different warps) and split every epilogue by 16-unit blocks, so 16 warps hide
//     the latency of the element-wise work (gates, LayerNorm, bias + ReLU), which is most of the kernel's instructions.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "uaisa_tc.cuh"

namespace uaisa {
using namespace tc;

constexpr int FW_M = 128, FW_H = 256, FW_IN = 9, FW_SELF = 12, FW_FEAT = 268;
constexpr int FW_TILE_G = 192 * 32;                              // floats of a GRU weight tile  [r(64) | z(64) | n(64)] x 32
constexpr int FW_TILE_W = 256 * 32;                              // floats of a heads weight tile
constexpr int FW_OFF_GX = 0;                                     // [2 dir][4 slice]            x FW_TILE_G : weight_ih (k < 9, zero padded to 32)
constexpr int FW_OFF_GH = FW_OFF_GX + 2 * 4 * FW_TILE_G;         // [2 dir][4 slice][8 chunk]   x FW_TILE_G : weight_hh
constexpr int FW_OFF_W1 = FW_OFF_GH + 2 * 4 * 8 * FW_TILE_G;     // [9 chunk][2 head]           x FW_TILE_W : layer 1, K order [h(256) | self(12) | 0(20)]
constexpr int FW_OFF_W2 = FW_OFF_W1 + 18 * FW_TILE_W;            // [2 head][8 chunk]           x FW_TILE_W : layer 2
constexpr int FW_OFF_GB = FW_OFF_W2 + 16 * FW_TILE_W;            // [2 dir][4][256]: b_ir+b_hr, b_iz+b_hz, b_in, b_hn
constexpr int FW_OFF_LN = FW_OFF_GB + 2 * 4 * 256;               // ln weight[288], ln bias[288] in the kernel's K order
constexpr int FW_OFF_B1 = FW_OFF_LN + 2 * 288;                   // [2 head][256]
constexpr int FW_OFF_B2 = FW_OFF_B1 + 512;                       // [2 head][256]
constexpr int FW_OFF_W3 = FW_OFF_B2 + 512;                       // [4][256]: actor rows 0..2, critic row 3
constexpr int FW_OFF_B3 = FW_OFF_W3 + 1024;                      // [4]
constexpr int FW_OFF_H0 = FW_OFF_B3 + 16;                        // [256]: the reader's output for an agent WITHOUT neighbour rows (one zero row, h0 = 0): a constant of the weights
constexpr int FW_PACK_FLOATS = FW_OFF_H0 + 256;

constexpr int FW_A_CHUNK = FW_M * 128;                           // bytes of one A chunk [128 rows x 32 fp32]
constexpr int FW_STAGE = 256 * 128;                              // bytes of one B stage
constexpr int FW_SMEM = 1024 + 9 * FW_A_CHUNK + 2 * FW_STAGE + 64 + 4 * 128 * 4 * 4;      // + barriers + row-reduction scratch [4][128][4] floats

struct PackSrc {       // raw module parameters (device pointers), torch layouts
    const float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2];          // GRU, [768][9] / [768][256] / [768] (gate order r, z, n); [1] = reverse direction
    const float *ln_w, *ln_b;                                    // [268] in [self | h] order
    const float *W1[2], *b1[2], *W2[2], *b2[2], *W3[2], *b3[2];  // [0] actor, [1] critic; W1 [256][268], W2 [256][256], W3 [3 | 1][256]
};

// Gate non-linearities on the SFU: tanh.approx.f32 (one MUFU.TANH, max abs error 2^-10.99 ~ 4.9e-4) and sigmoid(x) = 0.5 tanh(x/2) + 0.5 (error 2.4e-4) --
// the same order as the TF32 rounding of the GEMM operands that feed them (10-bit mantissa).  The exp-and-IEEE-divide forms were 48 % of the kernel's
// executed instructions (profiles/r2/ncu_policy_forward_sparse.txt).
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// (row, k) of float index w inside a swizzled tile image (inverse of tc::swz_float_index)
__device__ __forceinline__ void unswz(int w, int &row, int &k) {
    const int grp = w >> 8, rem = w & 255, r7 = rem >> 5, pos = rem & 31;
    row = grp * 8 + r7; k = (((pos >> 2) ^ r7) << 2) | (pos & 3);
}

__global__ void uaisa_policy_pack_kernel(const __grid_constant__ PackSrc S, float *__restrict__ wp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FW_PACK_FLOATS) return;
    float val = 0.0f;
    if (i < FW_OFF_GH) {                                           // GRU input weights
        const int t = i / FW_TILE_G, dir = t >> 2, slice = t & 3;
        int row, k; unswz(i - t * FW_TILE_G, row, k);
        const int gate = row >> 6, u = slice * 64 + (row & 63);
        if (k < FW_IN) val = S.w_ih[dir][(gate * FW_H + u) * FW_IN + k];
    } else if (i < FW_OFF_W1) {                                    // GRU recurrent weights
        const int t = (i - FW_OFF_GH) / FW_TILE_G, dir = t >> 5, slice = (t >> 3) & 3, chunk = t & 7;
        int row, k; unswz((i - FW_OFF_GH) - t * FW_TILE_G, row, k);
        const int gate = row >> 6, u = slice * 64 + (row & 63);
        val = S.w_hh[dir][(size_t)(gate * FW_H + u) * FW_H + chunk * 32 + k];
    } else if (i < FW_OFF_W2) {                                    // heads layer 1, K re-ordered to [h | self | pad]
        const int t = (i - FW_OFF_W1) / FW_TILE_W, chunk = t >> 1, head = t & 1;
        int n, k; unswz((i - FW_OFF_W1) - t * FW_TILE_W, n, k);
        const int q = chunk * 32 + k;
        if (q < FW_H) val = S.W1[head][n * FW_FEAT + FW_SELF + q];
        else if (q < FW_FEAT) val = S.W1[head][n * FW_FEAT + (q - FW_H)];
    } else if (i < FW_OFF_GB) {                                    // heads layer 2
        const int t = (i - FW_OFF_W2) / FW_TILE_W, head = t >> 3, chunk = t & 7;
        int n, k; unswz((i - FW_OFF_W2) - t * FW_TILE_W, n, k);
        val = S.W2[head][n * FW_H + chunk * 32 + k];
    } else if (i < FW_OFF_LN) {
        const int j = i - FW_OFF_GB, dir = j >> 10, which = (j >> 8) & 3, u = j & 255;
        val = which == 0 ? S.b_ih[dir][u] + S.b_hh[dir][u] : which == 1 ? S.b_ih[dir][256 + u] + S.b_hh[dir][256 + u] : which == 2 ? S.b_ih[dir][512 + u] : S.b_hh[dir][512 + u];
    } else if (i < FW_OFF_B1) {
        const int j = i - FW_OFF_LN, q = j % 288;
        const float *src = j < 288 ? S.ln_w : S.ln_b;
        if (q < FW_H) val = src[FW_SELF + q]; else if (q < FW_FEAT) val = src[q - FW_H];
    } else if (i < FW_OFF_B2) { const int j = i - FW_OFF_B1; val = S.b1[j >> 8][j & 255]; }
    else if (i < FW_OFF_W3) { const int j = i - FW_OFF_B2; val = S.b2[j >> 8][j & 255]; }
    else if (i < FW_OFF_B3) { const int j = i - FW_OFF_W3, r = j >> 8; val = r < 3 ? S.W3[0][r * 256 + (j & 255)] : S.W3[1][j & 255]; }
    else if (i < FW_OFF_H0) { const int j = i - FW_OFF_B3; if (j < 3) val = S.b3[0][j]; else if (j == 3) val = S.b3[1][0]; }
    else {                                                         // one GRU step from h = 0 on a zero row, both directions summed: the operations (and their order) of
        const int u = i - FW_OFF_H0;                               // gates16() in the forward kernel on accumulators that are exactly 0
        float hd[2];
        for (int dir = 0; dir < 2; dir++) {
            const float br = S.b_ih[dir][u] + S.b_hh[dir][u], bz = S.b_ih[dir][256 + u] + S.b_hh[dir][256 + u], bin = S.b_ih[dir][512 + u], bhn = S.b_hh[dir][512 + u];
            float r = sigm(0.0f + br);
            const float z = sigm(0.0f + bz);
            r *= bhn;
            const float nn = tanh_fast(0.0f + bin + r);
            hd[dir] = (1.0f - z) * nn + 0.0f;
        }
        val = hd[0] + hd[1];
    }
    wp[i] = val;
}

struct FwdParams {
    const float *obs; const uint8_t *count; int A, obs_dim, nm;
    const float *wp;              // packed weights (uaisa_policy_pack)
    float *mu, *v;                // [A][3], [A]
    float *feat_out;              // optional [A][268] LayerNorm output in the module's [self | h] order (tests)
    float4 *scratch;              // [ceil(A/128)][64][128] float4: forward-direction state of CTAs that run the recurrence
    int *stats;                   // optional: [0] += CTAs that ran the recurrence, [1] = max steps seen
#ifdef UAISA_FWD_DBG
    int dbg;                      // timing experiments (results are WRONG): 1 = no weight copies after the first two, 2 = no MMAs, 4 = no element-wise epilogue math
#endif
};

// order in which the weight tiles are consumed (thread 0's prefetcher walks it independently of the consumer)
struct TileIter {
    int lmax, phase, dir, step, slice, sub, lin;
    __device__ __forceinline__ void init(int lmax_) { lmax = lmax_; phase = lmax_ == 0 ? 1 : 0; dir = step = slice = sub = lin = 0; }   // (no GRU tiles for a tile without neighbour rows)
    __device__ __forceinline__ bool next(const float *wp, const float *&addr, uint32_t &bytes) {
        if (phase == 0) {
            bytes = FW_TILE_G * 4;
            if (lmax == 1) {                                       // step 0 only: per slice, both directions
                addr = wp + FW_OFF_GX + ((lin & 1) * 4 + (lin >> 1)) * FW_TILE_G;
                if (++lin == 8) { phase = 1; lin = 0; }
                return true;
            }
            addr = sub == 0 ? wp + FW_OFF_GX + (dir * 4 + slice) * FW_TILE_G : wp + FW_OFF_GH + ((dir * 4 + slice) * 8 + (sub - 1)) * FW_TILE_G;
            const int nsub = step > 0 ? 9 : 1;
            if (++sub == nsub) { sub = 0; if (++slice == 4) { slice = 0; if (++step == lmax) { step = 0; if (++dir == 2) { phase = 1; lin = 0; } } } }
            return true;
        }
        if (phase == 1) { addr = wp + FW_OFF_W1 + lin * FW_TILE_W; bytes = FW_TILE_W * 4; if (++lin == 18) { phase = 2; lin = 0; } return true; }
        if (phase == 2) { addr = wp + FW_OFF_W2 + lin * FW_TILE_W; bytes = FW_TILE_W * 4; if (++lin == 16) phase = 3; return true; }
        return false;
    }
};

// FW_S threads share one agent row (thread t, t + 128, ...: same TMEM lane, different warps): the 16-unit blocks of every epilogue are dealt round-robin to
// them, so a CTA has 4 * FW_S warps to hide the latency of the element-wise work (gates, LayerNorm, bias + ReLU) that dominates the kernel's instruction count.
constexpr int FW_S = 4;
constexpr int FW_NT = FW_M * FW_S;

// CL = 2 (optional, UAISA_POLICY_CLUSTER=2): the kernel runs as thread-block clusters of two CTAs (two agent tiles on two SMs) that walk the SAME weight-tile
// sequence; every tile is fetched once per cluster -- each CTA bulk-copies one half and MULTICASTS it into both CTAs' stage (SASS UBLKCP.S.G.MULTICAST) -- which
// halves the L2 -> SM weight traffic.  A stage is refilled only after BOTH CTAs' MMAs have released it (the `empty` barriers count two multicast tcgen05.commit
// arrivals, UTCBAR.MULTICAST).  Measured no faster than CL = 1: the weight stream is bound by the latency of the two stages in flight, not by L2 bandwidth.
template <int CL>
__global__ void __launch_bounds__(FW_NT, 1) uaisa_policy_forward_kernel(const __grid_constant__ FwdParams P) {
    extern __shared__ unsigned char fw_smem_raw[];
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(fw_smem_raw) + 1023) & ~uintptr_t(1023));   // SWIZZLE_128B needs 1024-byte alignment
    char *As = base;                                   // 9 chunks [128 x 32] fp32: chunks 0..7 = GRU state / features h / hidden activations, chunk 8 = x_t / self
    char *Bs = As + 9 * FW_A_CHUNK;                    // 2 stages
    uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + 2 * FW_STAGE);   // full[2], empty[2], done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5);
    int *s_lmax = reinterpret_cast<int *>(tmem_slot + 1);
    float *s_red = reinterpret_cast<float *>(bars + 8);                 // [FW_S][128][4] partial sums of the row reductions (LayerNorm, output layer)
    const int tid = threadIdx.x, lane = tid & 31, row = tid & (FW_M - 1), sub = tid >> 7;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // (provably warp-uniform: the issuer block below stays on the uniform datapath)
    const int a0 = blockIdx.x * FW_M, a = a0 + row;
    const bool ok = a < P.A;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]), done = smem_u32(&bars[4]);

    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    // every global input of the tile's first phases is requested BEFORE the set-up barriers -- the row count, the 12 self values and neighbour row 0 (21
    // consecutive floats of the agent's observation) -- so the tile pays one exposed memory latency instead of three (count -> x tile -> self at the LayerNorm)
    const float *orow = P.obs + (size_t)(ok ? a : 0) * P.obs_dim;
    int cnt = ok ? (int)P.count[a] : 0;
    float in0[FW_SELF + FW_IN];
    if (sub == 0) {
#pragma unroll
        for (int k = 0; k < FW_SELF + FW_IN; k++) in0[k] = (ok && (k < FW_SELF || P.nm > 0)) ? __ldg(orow + k) : 0.0f;
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512u);
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);          // full: my own arrive.expect_tx (+ the tile's bytes)
        mbar_init(smem_u32(&bars[2]), CL); mbar_init(smem_u32(&bars[3]), CL);        // empty: one tcgen05.commit arrival per CTA of the cluster
        mbar_init(smem_u32(&bars[4]), 1);                                            // done
        mbar_init_fence();
        *s_lmax = 0;
    }
    __syncthreads();
    cnt = cnt < P.nm ? cnt : P.nm;
    const int c_eff = cnt > 1 ? cnt : 1;               // an agent without neighbours feeds ONE zero row (ir_gym.py:211-212)
    {   // number of GRU steps = the largest row count in the CLUSTER's tiles (all its CTAs must walk the same weight-tile sequence)
        int m = cnt;
        if (CL > 1) {
            for (int q = tid; q < CL * FW_M; q += FW_NT) {
                const int aq = (int)(blockIdx.x / CL) * CL * FW_M + q;
                int c = aq < P.A ? (int)P.count[aq] : 0;
                c = c < P.nm ? c : P.nm;
                m = c > m ? c : m;
            }
        }
        m = __reduce_max_sync(0xffffffffu, m);
        if ((tid & 31) == 0) atomicMax(s_lmax, m);
    }
    {   // A chunk 8, my row: neighbour row 0 in K columns 0..8 (zero up to 15: the chunk is multiplied with K = 16), the self values parked in columns 16..27
        // (never read by an MMA) until the LayerNorm
        char *t8 = As + 8 * FW_A_CHUNK;
        if (sub == 0) {
            if (cnt == 0) {                            // no neighbours: the reader is fed ONE ZERO row whatever the observation holds there (ir_gym.py:211-212)
#pragma unroll
                for (int k = 0; k < FW_IN; k++) in0[FW_SELF + k] = 0.0f;
            }
            *swz(t8, row, 0) = make_float4(in0[12], in0[13], in0[14], in0[15]); *swz(t8, row, 1) = make_float4(in0[16], in0[17], in0[18], in0[19]);
            *swz(t8, row, 2) = make_float4(in0[20], 0.f, 0.f, 0.f); *swz(t8, row, 3) = make_float4(0.f, 0.f, 0.f, 0.f);
            *swz(t8, row, 4) = make_float4(in0[0], in0[1], in0[2], in0[3]); *swz(t8, row, 5) = make_float4(in0[4], in0[5], in0[6], in0[7]);
            *swz(t8, row, 6) = make_float4(in0[8], in0[9], in0[10], in0[11]);
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (CL > 1) cluster_sync();                        // the peers' barriers are initialised before anything is multicast to them
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // (warp-uniform values, provably)
    const int lmax = __shfl_sync(0xffffffffu, *s_lmax, 0);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;      // a warp reaches the TMEM lane quarter (warp id % 4)
    const float *wp = P.wp;
    uint32_t dpar = 0;                                 // parity of the `done` barrier (all threads)
    // ---- warp 0: weight-tile ring and MMA issue.  The WHOLE warp runs this code (warp-uniform control flow and state, so ring indices, parities and the
    // matrix descriptors live on the uniform datapath); lane 0 alone executes the instructions with side effects (expect_tx, bulk copy, tcgen05.mma / commit).
    // With `if (tid == 0)` around everything the compiler kept the state in vector registers and local memory and moved five registers into uniform ones for
    // every MMA: 6 800 single-thread instructions per tile, half of warp 0's time -- and every other warp waits for it (profiles/r2/policy_dbg.txt).
    TileIter it; it.init(lmax);
    int pf = 0, cs = 0; uint32_t fmask = 0u, emask = 0u;               // tiles requested / consumed; parity bit per stage of the full / empty barriers
    auto topup = [&]() {
        while (pf < cs + 2) {
            const float *addr; uint32_t bytes;
            if (!it.next(wp, addr, bytes)) break;
            const int s = pf & 1;
            if (pf >= 2) { mbar_wait(empty0 + 8 * s, (emask >> s) & 1u); emask ^= 1u << s; }      // the MMAs that read this stage have completed
#ifdef UAISA_FWD_DBG
            if ((P.dbg & 1) && pf >= 2) { pf++; continue; }
#endif
            if (lane == 0) {
                mbar_expect_tx(full0 + 8 * s, bytes);
                if (CL == 1) bulk_g2s(smem_u32(Bs + s * FW_STAGE), addr, bytes, full0 + 8 * s);
                else {                                 // my 1/CL of the tile, into every CTA of the cluster
                    const uint32_t part = bytes / CL;
                    bulk_g2s_multicast(smem_u32(Bs + s * FW_STAGE) + crank * part, reinterpret_cast<const char *>(addr) + crank * part, part, full0 + 8 * s, CMASK);
                }
            }
            pf++;
        }
    };
    auto acquire = [&]() -> int {
        topup();
        const int s = cs & 1;
#ifdef UAISA_FWD_DBG
        if ((P.dbg & 1) && cs >= 2) return s;
#endif
        mbar_wait(full0 + 8 * s, (fmask >> s) & 1u); fmask ^= 1u << s;
        fence_after_sync();
        return s;
    };
    auto release = [&](int s) {
        if (lane == 0) { if (CL == 1) umma_commit(empty0 + 8 * s); else umma_commit_multicast(empty0 + 8 * s, CMASK); }
        cs++;
    };
    auto mma = [&](int a_chunk, uint32_t b_saddr, int ksteps, uint32_t idesc, uint32_t tmem_d, bool acc) {
        const uint64_t da = umma_desc(smem_u32(As + a_chunk * FW_A_CHUNK)), db = umma_desc(b_saddr);
#ifdef UAISA_FWD_DBG
        if (P.dbg & 2) return;
#endif
        if (lane == 0) {                               // UMMA K = 8 tf32 = 32 bytes: + 2 in the descriptor's (address >> 4) field per step
            umma_tf32(tmem_d, da, db, idesc, acc ? 1u : 0u);
            umma_tf32(tmem_d, da + 2, db + 2, idesc, 1u);
            if (ksteps == 4) { umma_tf32(tmem_d, da + 4, db + 4, idesc, 1u); umma_tf32(tmem_d, da + 6, db + 6, idesc, 1u); }
        }
    };
    auto group_done = [&]() { if (lane == 0) umma_commit(done); topup(); };   // (warp 0) accumulators ready -> `done`; next tiles stream during the epilogue
    auto wait_done = [&]() { mbar_wait(done, dpar); dpar ^= 1u; __syncwarp(); fence_after_sync(); };   // (warp 0 reconverges after lane 0's instructions: tcgen05.ld / st are .aligned)
    auto sync_smem_to_mma = [&]() { fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync(); };
    auto sync_tmem = [&]() { fence_before_sync(); __syncthreads(); fence_after_sync(); };
    constexpr uint32_t ID192 = umma_idesc_tf32(128, 192), ID128 = umma_idesc_tf32(128, 128), ID64 = umma_idesc_tf32(128, 64), ID256 = umma_idesc_tf32(128, 256);

    // x_t of my row into the A tile of chunk 8 (9 values, zero padded to 16 = the K the chunk is multiplied with)
    auto write_x = [&](int idx, bool valid) {
        char *t8 = As + 8 * FW_A_CHUNK;
        if (sub == 0) {
            float x[12];
#pragma unroll
            for (int k = 0; k < 12; k++) x[k] = 0.0f;
            if (valid && ok && idx < P.nm && cnt > 0) {         // (cnt == 0: the one row the reader is fed is a zero row, whatever the observation holds)
#pragma unroll
                for (int k = 0; k < FW_IN; k++) x[k] = __ldg(orow + FW_SELF + FW_IN * idx + k);
            }
            *swz(t8, row, 0) = make_float4(x[0], x[1], x[2], x[3]); *swz(t8, row, 1) = make_float4(x[4], x[5], x[6], x[7]);
            *swz(t8, row, 2) = make_float4(x[8], 0.f, 0.f, 0.f);
        }                                              // (K columns 12..15 stay zero from the set-up, 16..31 are never multiplied)
    };
    auto ld4 = [&](const float *p, float *o) { const float4 f = __ldg(reinterpret_cast<const float4 *>(p)); o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = f.w; };
    auto ld16 = [&](const float *p, float *o) { ld4(p, o); ld4(p + 4, o + 4); ld4(p + 8, o + 8); ld4(p + 12, o + 12); };
    // gates of 16 hidden units [u0, u0+16) of direction `dir` from the accumulator columns at `dcol`: out = (1 - z) n + z h_old
    auto gates16 = [&](int dir, int u0, uint32_t dcol, int jb, bool has_h, const float *h_old, float *out) {
        float r[16], z[16], gi[16], gh[16], b[16];
#ifdef UAISA_FWD_DBG
        if (P.dbg & 4) { for (int q = 0; q < 16; q++) out[q] = 0.25f; return; }
#endif
        const float *gb = wp + FW_OFF_GB + dir * 1024 + u0;
        tmem_ld16(tmem + lane_base + dcol + jb * 16, r);
        ld16(gb, b);
#pragma unroll
        for (int q = 0; q < 16; q++) r[q] = sigm(r[q] + b[q]);
        tmem_ld16(tmem + lane_base + dcol + 64 + jb * 16, z);
        ld16(gb + 256, b);
#pragma unroll
        for (int q = 0; q < 16; q++) z[q] = sigm(z[q] + b[q]);
        ld16(gb + 768, b);
        if (has_h) {
            tmem_ld16(tmem + lane_base + dcol + 192 + jb * 16, gh);
#pragma unroll
            for (int q = 0; q < 16; q++) r[q] *= gh[q] + b[q];
        } else {
#pragma unroll
            for (int q = 0; q < 16; q++) r[q] *= b[q];
        }
        tmem_ld16(tmem + lane_base + dcol + 128 + jb * 16, gi);
        ld16(gb + 512, b);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const float nn = tanh_fast(gi[q] + b[q] + r[q]);
            out[q] = (1.0f - z[q]) * nn + (has_h ? z[q] * h_old[q] : 0.0f);
        }
    };
    auto read_h16 = [&](int u0, float *h) {            // my row's state, units [u0, u0+16), from the swizzled A tile
        char *t = As + (u0 >> 5) * FW_A_CHUNK;
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) { const float4 f = *swz(t, row, ((u0 & 31) >> 2) + q4); h[4 * q4] = f.x; h[4 * q4 + 1] = f.y; h[4 * q4 + 2] = f.z; h[4 * q4 + 3] = f.w; }
    };
    auto write_h16 = [&](int u0, const float *h) {
        char *t = As + (u0 >> 5) * FW_A_CHUNK;
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) *swz(t, row, ((u0 & 31) >> 2) + q4) = make_float4(h[4 * q4], h[4 * q4 + 1], h[4 * q4 + 2], h[4 * q4 + 3]);
    };
    // sum over the FW_S sub-threads of a row of up to 4 partial values (through shared memory)
    auto row_reduce = [&](float *v, int n) {
        for (int i = 0; i < n; i++) s_red[(sub * FW_M + row) * 4 + i] = v[i];
        __syncthreads();
        for (int i = 0; i < n; i++) { float t = 0.0f;
#pragma unroll
            for (int q = 0; q < FW_S; q++) t += s_red[(q * FW_M + row) * 4 + i];
            v[i] = t; }
        __syncthreads();
    };

    // =========================================================================================== bi-GRU reader
    if (lmax == 0) {
        // no agent of the tile has a neighbour row: the reader's output is the same vector for every agent, h0 (packed with the weights) -- no GRU tile is fetched,
        // no MMA issued, no gate evaluated.  Bit-identical to the general path, whose accumulators are exactly 0 for a zero row
        // (tests/test_policy.py::test_tiles_without_neighbour_rows_take_the_constant_reader_output).  Sparse airspace: the whole rollout of configs[1] / [4].
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) { float h[16]; ld16(wp + FW_OFF_H0 + ub * 16, h); write_h16(ub * 16, h); }
    } else if (lmax == 1) {
        // every agent of the tile has <= 1 row: one step from h = 0 in both directions (no recurrent GEMM); both directions' accumulators side by side
        for (int slice = 0; slice < 4; slice++) {          // (row 0 is in the A tile since the set-up)
            if (warp == 0) {
                int s = acquire(); mma(8, smem_u32(Bs + s * FW_STAGE), 2, ID192, tmem + 0, false); release(s);
                s = acquire(); mma(8, smem_u32(Bs + s * FW_STAGE), 2, ID192, tmem + 256, false); release(s);
                group_done();
            }
            wait_done();
#pragma unroll 1
            for (int jb = sub; jb < 4; jb += FW_S) {
                const int u0 = slice * 64 + jb * 16;
                float hf[16], hb[16];
                gates16(0, u0, 0, jb, false, nullptr, hf);
                gates16(1, u0, 256, jb, false, nullptr, hb);
#pragma unroll
                for (int q = 0; q < 16; q++) hf[q] += hb[q];
                write_h16(u0, hf);
            }
            sync_tmem();                               // the accumulators are free for the next slice
        }
    } else {
        float4 *stash = P.scratch + (size_t)blockIdx.x * 64 * FW_M;
        for (int dir = 0; dir < 2; dir++) {
            for (int step = 0; step < lmax; step++) {
                const bool active = step < c_eff;
                // ---- this step's x tile; (from step 1 on) the previous step's new state, TMEM -> A tile
                if (dir | step) write_x(dir == 0 ? step : c_eff - 1 - step, active);       // (row 0 of the first step is in the A tile since the set-up)
                if (step > 0) {
#pragma unroll 1
                    for (int ub = sub; ub < 16; ub += FW_S) { float h[16]; tmem_ld16(tmem + lane_base + 256 + ub * 16, h); write_h16(ub * 16, h); }
                }
                sync_smem_to_mma();
                for (int slice = 0; slice < 4; slice++) {
                    if (warp == 0) {
                        int s = acquire(); mma(8, smem_u32(Bs + s * FW_STAGE), 2, ID192, tmem + 0, false); release(s);           // x_t W_i^T -> [r | z | i_n]
                        if (step > 0) {
                            for (int c = 0; c < 8; c++) {                                                                         // h W_h^T
                                s = acquire();
                                const uint32_t b = smem_u32(Bs + s * FW_STAGE);
                                mma(c, b, 4, ID128, tmem + 0, true);                                                              //   -> += [r | z]
                                mma(c, b + 128 * 128, 4, ID64, tmem + 192, c > 0);                                                //   -> h_n (kept apart: n = tanh(i_n + r * h_n))
                                release(s);
                            }
                        }
                        group_done();
                    }
                    wait_done();
#pragma unroll 1
                    for (int jb = sub; jb < 4; jb += FW_S) {
                        const int u0 = slice * 64 + jb * 16;
                        float ho[16], hn[16];
                        if (step > 0) read_h16(u0, ho);
                        gates16(dir, u0, 0, jb, step > 0, ho, hn);
                        if (!active) {
#pragma unroll
                            for (int q = 0; q < 16; q++) hn[q] = ho[q];      // (step > 0 here: every row is active at step 0)
                        }
                        tmem_st16(tmem + lane_base + 256 + u0, hn);
                    }
                    tmem_st_wait();
                    sync_tmem();                       // accumulator columns free; the new state of this slice is in TMEM
                }
            }
            // ---- final state of this direction (TMEM columns 256..511)
            if (dir == 0) {
#pragma unroll 1
                for (int ub = sub; ub < 16; ub += FW_S) {
                    float h[16]; tmem_ld16(tmem + lane_base + 256 + ub * 16, h);
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) if (a0 < P.A) stash[(size_t)(ub * 4 + q4) * FW_M + row] = make_float4(h[4 * q4], h[4 * q4 + 1], h[4 * q4 + 2], h[4 * q4 + 3]);
                }
            } else {
#pragma unroll 1
                for (int ub = sub; ub < 16; ub += FW_S) {
                    float h[16]; tmem_ld16(tmem + lane_base + 256 + ub * 16, h);
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) { const float4 f = a0 < P.A ? stash[(size_t)(ub * 4 + q4) * FW_M + row] : make_float4(0.f, 0.f, 0.f, 0.f); h[4 * q4] += f.x; h[4 * q4 + 1] += f.y; h[4 * q4 + 2] += f.z; h[4 * q4 + 3] += f.w; }
                    write_h16(ub * 16, h);             // h_fwd + h_bwd (policy_rnn_ac.py:122, 163)
                }
            }
            sync_tmem();
        }
        if (P.stats && tid == 0 && a0 < P.A) { atomicAdd(P.stats, 1); atomicMax(P.stats + 1, lmax); }
    }

    // =========================================================================================== concat + LayerNorm(268), in place in the A tile
#ifdef UAISA_FWD_DBG
    if (!(P.dbg & 8))
#endif
    {
        float self[FW_SELF];                           // parked in K columns 16..27 of A chunk 8 by the set-up
#pragma unroll
        for (int k = 0; k < FW_SELF; k++) self[k] = 0.0f;
        if (sub == 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) { const float4 f = *swz(As + 8 * FW_A_CHUNK, row, 4 + j); self[4 * j] = f.x; self[4 * j + 1] = f.y; self[4 * j + 2] = f.z; self[4 * j + 3] = f.w; }
        }
        float red[1] = {0.0f};
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < FW_SELF; k++) red[0] += self[k];
        }
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) { float h[16]; read_h16(ub * 16, h);
#pragma unroll
            for (int q = 0; q < 16; q++) red[0] += h[q]; }
        row_reduce(red, 1);
        const float mean = red[0] * (1.0f / FW_FEAT);
        red[0] = 0.0f;
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < FW_SELF; k++) red[0] += (self[k] - mean) * (self[k] - mean);
        }
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) { float h[16]; read_h16(ub * 16, h);
#pragma unroll
            for (int q = 0; q < 16; q++) red[0] += (h[q] - mean) * (h[q] - mean); }
        row_reduce(red, 1);
        const float rstd = rsqrtf(red[0] * (1.0f / FW_FEAT) + 1e-5f);                  // nn.LayerNorm: biased variance, eps 1e-5
        const float *lw = wp + FW_OFF_LN, *lb = lw + 288;
        float *fo = (P.feat_out && ok) ? P.feat_out + (size_t)a * FW_FEAT : nullptr;
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) {
            float h[16], w[16], b[16]; read_h16(ub * 16, h); ld16(lw + ub * 16, w); ld16(lb + ub * 16, b);
#pragma unroll
            for (int q = 0; q < 16; q++) h[q] = (h[q] - mean) * rstd * w[q] + b[q];
            write_h16(ub * 16, h);
            if (fo) {
#pragma unroll
                for (int q = 0; q < 16; q++) fo[FW_SELF + ub * 16 + q] = h[q];
            }
        }
        char *t8 = As + 8 * FW_A_CHUNK;
        if (sub == 0) {
            float sn[FW_SELF];
#pragma unroll
            for (int k = 0; k < FW_SELF; k++) { sn[k] = (self[k] - mean) * rstd * __ldg(lw + 256 + k) + __ldg(lb + 256 + k); if (fo) fo[k] = sn[k]; }
            *swz(t8, row, 0) = make_float4(sn[0], sn[1], sn[2], sn[3]); *swz(t8, row, 1) = make_float4(sn[4], sn[5], sn[6], sn[7]);
            *swz(t8, row, 2) = make_float4(sn[8], sn[9], sn[10], sn[11]);
        }
        for (int k4 = 3 + sub; k4 < 8; k4 += FW_S) *swz(t8, row, k4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    sync_smem_to_mma();

    // =========================================================================================== actor / critic heads
    if (warp == 0) {                                   // layer 1 of both heads from the same A (features): columns 0..255 actor, 256..511 critic
        for (int c = 0; c < 9; c++) {
            const int ks = c < 8 ? 4 : 2;              // chunk 8 holds the 12 self values: K = 16 covers them
            int s = acquire(); mma(c, smem_u32(Bs + s * FW_STAGE), ks, ID256, tmem + 0, c > 0); release(s);
            s = acquire(); mma(c, smem_u32(Bs + s * FW_STAGE), ks, ID256, tmem + 256, c > 0); release(s);
        }
        group_done();
    }
    wait_done();
    for (int head = 0; head < 2; head++) {
        // epilogue of layer 1: bias + ReLU -> A chunks 0..7 (the features are no longer needed)
        const float *b1 = wp + FW_OFF_B1 + head * 256;
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) {
#ifdef UAISA_FWD_DBG
            if (P.dbg & 16) break;
#endif
            float h[16], b[16]; tmem_ld16(tmem + lane_base + (head ? 256 : 0) + ub * 16, h); ld16(b1 + ub * 16, b);
#pragma unroll
            for (int q = 0; q < 16; q++) h[q] = fmaxf(h[q] + b[q], 0.0f);
            write_h16(ub * 16, h);
        }
        sync_smem_to_mma();
        if (warp == 0) {                               // layer 2 -> columns 0..255 (the critic's layer-1 accumulator still sits in 256..511)
            for (int c = 0; c < 8; c++) { const int s = acquire(); mma(c, smem_u32(Bs + s * FW_STAGE), 4, ID256, tmem + 0, c > 0); release(s); }
            group_done();
        }
        wait_done();
        // epilogue of layer 2: bias + ReLU, then the tiny output layer on the CUDA cores (partial dot products per sub-thread, summed through shared memory)
        const float *b2 = wp + FW_OFF_B2 + head * 256, *w3 = wp + FW_OFF_W3 + (head ? 768 : 0);
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int ub = sub; ub < 16; ub += FW_S) {
#ifdef UAISA_FWD_DBG
            if (P.dbg & 16) break;
#endif
            float h[16], b[16], w[16]; tmem_ld16(tmem + lane_base + ub * 16, h); ld16(b2 + ub * 16, b);
#pragma unroll
            for (int q = 0; q < 16; q++) h[q] = fmaxf(h[q] + b[q], 0.0f);
            ld16(w3 + ub * 16, w);
#pragma unroll
            for (int q = 0; q < 16; q++) acc[0] = fmaf(h[q], w[q], acc[0]);
            if (!head) {
                ld16(w3 + 256 + ub * 16, w);
#pragma unroll
                for (int q = 0; q < 16; q++) acc[1] = fmaf(h[q], w[q], acc[1]);
                ld16(w3 + 512 + ub * 16, w);
#pragma unroll
                for (int q = 0; q < 16; q++) acc[2] = fmaf(h[q], w[q], acc[2]);
            }
        }
        fence_before_sync();
        row_reduce(acc, head ? 1 : 3);                 // (its barriers also order the TMEM reads above before the next head's layer 2 overwrites columns 0..255)
        fence_after_sync();
        if (ok && sub == 0) {
            const float *b3 = wp + FW_OFF_B3;
            if (!head) { P.mu[(size_t)a * 3] = tanhf(acc[0] + __ldg(b3)); P.mu[(size_t)a * 3 + 1] = tanhf(acc[1] + __ldg(b3 + 1)); P.mu[(size_t)a * 3 + 2] = tanhf(acc[2] + __ldg(b3 + 2)); }
            else P.v[a] = acc[0] + __ldg(b3 + 3);
        }
    }
    if (CL > 1) cluster_sync();                        // no peer may still signal my barriers (or expect my copies) when I leave
    if (warp == 0) tmem_dealloc(tmem, 512u);
}
template __global__ void uaisa_policy_forward_kernel<1>(const __grid_constant__ FwdParams);
template __global__ void uaisa_policy_forward_kernel<2>(const __grid_constant__ FwdParams);


// rnn_ac.step's sampling (policy_rnn_ac.py:57-69 with GaussianActor._distribution :206-228 and _log_prob_from_distribution): std = clamp(std_factor exp(log_std) + 1e-6,
// 1e-4, 10), a = mu + std eps, logp = sum_q( -(a - mu)^2 / (2 std^2) - log std - log sqrt(2 pi) ) -- the dozen element-wise torch kernels of one rollout step
// in ONE launch; every operation in the order (and rounding) of the torch expression it replaces.  eps comes from the caller (torch's generator).
__global__ void uaisa_policy_sample_kernel(const float *__restrict__ mu, const float *__restrict__ eps, const float *__restrict__ log_std, float std_factor, int A,
                                           float *__restrict__ a_out, float *__restrict__ logp) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    float lp = 0.0f;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float sd = fminf(fmaxf(__fadd_rn(__fmul_rn(std_factor, expf(__ldg(log_std + q))), 1e-6f), 1e-4f), 10.0f);
        const float m = mu[3 * a + q];
        const float x = __fadd_rn(m, __fmul_rn(sd, eps[3 * a + q]));
        const float d = __fsub_rn(x, m);
        float t = __fdiv_rn(-__fmul_rn(d, d), __fmul_rn(__fmul_rn(2.0f, sd), sd));
        t = __fsub_rn(__fsub_rn(t, logf(sd)), 0.9189385332046727f);
        lp = q == 0 ? t : __fadd_rn(lp, t);
        a_out[3 * a + q] = x;
    }
    logp[a] = lp;
}

}  // namespace uaisa

// ---- C ABI -----------------------------------------------------------------------------------------------------------
extern "C" {

int uaisa_policy_pack_floats(void) { return uaisa::FW_PACK_FLOATS; }

// params: 22 DEVICE pointers in the order  weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, the same four of the reverse direction, ln.weight, ln.bias,
// actor W1 b1 W2 b2 W3 b3, critic W1 b1 W2 b2 W3 b3  (torch layouts, fp32, contiguous).  wpack: [uaisa_policy_pack_floats()] floats, 16-byte aligned.
int uaisa_policy_pack(const float *const *params, float *wpack, void *stream) {
    if (!params || !wpack || (reinterpret_cast<uintptr_t>(wpack) & 15)) return -1;
    for (int i = 0; i < 22; i++) if (!params[i]) return -1;
    uaisa::PackSrc S{};
    for (int d = 0; d < 2; d++) { S.w_ih[d] = params[4 * d]; S.w_hh[d] = params[4 * d + 1]; S.b_ih[d] = params[4 * d + 2]; S.b_hh[d] = params[4 * d + 3]; }
    S.ln_w = params[8]; S.ln_b = params[9];
    for (int h = 0; h < 2; h++) {
        S.W1[h] = params[10 + 6 * h]; S.b1[h] = params[11 + 6 * h]; S.W2[h] = params[12 + 6 * h];
        S.b2[h] = params[13 + 6 * h]; S.W3[h] = params[14 + 6 * h]; S.b3[h] = params[15 + 6 * h];
    }
    uaisa::uaisa_policy_pack_kernel<<<(uaisa::FW_PACK_FLOATS + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, wpack);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int uaisa_policy_forward(const float *obs, const uint8_t *count, int32_t A, int32_t obs_dim, const float *wpack, float *mu, float *v, float *feat_out,
                         void *scratch, int32_t *stats, void *stream) {
    if (!obs || !count || !wpack || !mu || !v || !scratch || A < 1 || obs_dim < uaisa::FW_SELF || (obs_dim - uaisa::FW_SELF) % uaisa::FW_IN) return -1;
    static bool attr_set[64] = {};      // per device
    // UAISA_POLICY_CLUSTER=2: clusters of two CTAs that multicast every weight tile into both (A/B switch).  Measured NO faster (0.72 vs 0.69 ms at 262 144
    // agents): the stall is the latency of two 32 KB stages in flight per SM, not L2 bandwidth -- see DESIGN.md section 6.  Default: 1.
    static const int cl = std::getenv("UAISA_POLICY_CLUSTER") ? std::atoi(std::getenv("UAISA_POLICY_CLUSTER")) : 1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(uaisa::uaisa_policy_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, uaisa::FW_SMEM) != cudaSuccess) return -2;
        if (cudaFuncSetAttribute(uaisa::uaisa_policy_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, uaisa::FW_SMEM) != cudaSuccess) return -2;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    uaisa::FwdParams P{};
    P.obs = obs; P.count = count; P.A = A; P.obs_dim = obs_dim; P.nm = (obs_dim - uaisa::FW_SELF) / uaisa::FW_IN;
#ifdef UAISA_FWD_DBG
    P.dbg = std::getenv("UAISA_FWD_DBG") ? std::atoi(std::getenv("UAISA_FWD_DBG")) : 0;
#endif
    P.wp = wpack; P.mu = mu; P.v = v; P.feat_out = feat_out; P.scratch = static_cast<float4 *>(scratch); P.stats = stats;
    const int tiles = (A + uaisa::FW_M - 1) / uaisa::FW_M;
    if (cl == 2) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(((tiles + 1) / 2) * 2); cfg.blockDim = dim3(uaisa::FW_NT); cfg.dynamicSmemBytes = uaisa::FW_SMEM; cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, uaisa::uaisa_policy_forward_kernel<2>, P) != cudaSuccess) return -2;
    } else {
        uaisa::uaisa_policy_forward_kernel<1><<<tiles, uaisa::FW_NT, uaisa::FW_SMEM, static_cast<cudaStream_t>(stream)>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int uaisa_policy_sample(const float *mu, const float *eps, const float *log_std, float std_factor, int32_t A, float *a_out, float *logp, void *stream) {
    if (!mu || !eps || !log_std || !a_out || !logp || A < 1) return -1;
    uaisa::uaisa_policy_sample_kernel<<<(A + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(mu, eps, log_std, std_factor, A, a_out, logp);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"

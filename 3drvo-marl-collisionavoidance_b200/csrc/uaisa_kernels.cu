// uaisa_kernels.cu -- the env kernel (all modes) for sm_100a.  See uaisa_kernels.cuh for the design.
//
// Structure of one env step inside the kernel (one thread = one drone, one CTA = a block of whole envs):
//
//   zero-fill this step's obs/nbr output region (coalesced 16-byte stores; rows are mostly padding)
//   phase A   publish pre-move state -> pairwise pass (rvo reward) -> move (kinematics + waypoint FSM)
//   phase B   publish post-move state (fp64) -> pairwise pass (collision, top-nm rows) -> flags, reward,
//             scatter the non-zero part of the observation row straight to HBM
//   phase OBS (only envs touched by the trainer reset policy, or the stand-alone observe/reset modes)
//             publish -> pairwise pass with action 0 -> rewrite the observation row
//
// The three phases run through ONE copy of the dronestate / pair / row-writer code (a 3-trip loop), which
// keeps the kernel small enough for the instruction cache and the live register set below 168.
//
// Geometries (template <MAXR, MAXT, WL, SL>, chosen by uaisa_create):
//   WL  N divides 32: every env inside one warp, barriers are __syncwarp, reductions are ballots, output rows are cleared by a
//       TMA bulk store per warp (cp.async.bulk shared -> global)
//   CTA whole envs per CTA, __syncthreads; optional CTA-level compaction of the re-observe phase
//   CLM N > 128: one env per thread-block cluster, columns exchanged through distributed shared memory (or L2) around cluster barriers
//   SL  neighbour slices: helper warps walk a share of each ego's neighbour chunks and hand partial results to the owner
// Other pieces: per-drone route cache and map sizes in shared memory; the host-pack tail of uaisa_step_host_sparse (results
// written straight to mapped host memory, completion word raised by the last CTA).
#include "uaisa_kernels.cuh"

namespace uaisa {

// fp32 round trip: what a store to HBM followed by the next step's load would do
__device__ __forceinline__ double rt(double x) { return (double)(float)x; }

template <class CONSTS>
__device__ __forceinline__ void reset_drone(const Params &P, Drone &d, const CONSTS &C) {   // drone.reset, drone.py:270-291
    load3(C.wp(P), d.px, d.py, d.pz);
    d.px = rt(d.px); d.py = rt(d.py); d.pz = rt(d.pz);   // the state is fp32: a start point is stored rounded
    d.vx = d.vy = d.vz = 0.0; d.yaw = d.pitch = 0.0; d.real_len = 0.0; d.max_dev = 0.0;
    d.i = 1; d.flags = 0;   // extra_len is deliberately not cleared
}

// n 32-bit words of `v` to global, widest stores the destination alignment allows
__device__ __forceinline__ void fill_out(uint32_t *__restrict__ dst, int n, uint32_t v, int tid, int nthreads) {
    const uintptr_t da = reinterpret_cast<uintptr_t>(dst);
    int head = (int)(((16 - (da & 15)) & 15) >> 2);   // words until 16-byte alignment
    if (head > n) head = n;
    if (tid < head) dst[tid] = v;
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    const int n4 = (n - head) >> 2;
    const uint4 vv = make_uint4(v, v, v, v);
    for (int i = tid; i < n4; i += nthreads) d4[i] = vv;
    for (int i = head + (n4 << 2) + tid; i < n; i += nthreads) dst[i] = v;
}

// ---- TMA bulk stores (cp.async.bulk shared::cta -> global): the zero padding that makes up most of every
// observation row is streamed to HBM by the copy engine from a constant shared-memory tile, off the warps'
// critical path and out of the LSU store queue.
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// called by ONE lane: fill [dst, dst+bytes) from the constant tile (dst 16-byte aligned, bytes % 16 == 0)
__device__ __forceinline__ void bulk_fill(char *dst, const void *tile, int tile_bytes, int bytes) {
    for (int off = 0; off < bytes; off += tile_bytes) bulk_store(dst + off, tile, (uint32_t)min(tile_bytes, bytes - off));   // normally ONE op
}

// hardware barrier over all threads of the thread-block cluster, with release/acquire ordering at cluster scope
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// distributed shared memory: the address of `saddr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster, and stores to it
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_v4f32(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
enum Phase : int { PH_A = 0, PH_B = 1, PH_OBS = 2 };

// Optional per-region cycle accounting (build with -DUAISA_REGION_CLOCKS; lane 0 of every warp accumulates
// clock64() deltas into counters[8+region]).  A profiling aid only; compiled out of the product build.
enum Region : int { RG_PROLOGUE = 0, RG_FILL, RG_SYNC, RG_DRONESTATE, RG_PAIRS, RG_REWARD_A, RG_MOVE, RG_OBSROW, RG_TAIL_B, RG_AUTORESET, RG_EPILOGUE, RG_B_FLAGS, RG_B_COL, RG_B_REWARD, RG_B_OUT, RG_ACT, RG_COUNT };
#ifdef UAISA_REGION_CLOCKS
#define RG_DECL long long rg_acc[RG_COUNT]; for (int q_ = 0; q_ < RG_COUNT; q_++) rg_acc[q_] = 0; long long rg_t = clock64();
#define RG_MARK(r) do { const long long n_ = clock64(); rg_acc[r] += n_ - rg_t; rg_t = n_; } while (0)
#define RG_FLUSH() do { if ((threadIdx.x & 31) == 0) for (int q_ = 0; q_ < RG_COUNT; q_++) atomicAdd(P.counters + 8 + q_, (unsigned long long)rg_acc[q_]); } while (0)
#else
#define RG_DECL
#define RG_MARK(r) do { } while (0)
#define RG_FLUSH() do { } while (0)
#endif

#define UAISA_PRAGMA_(x) _Pragma(#x)
#define UAISA_PRAGMA(x) UAISA_PRAGMA_(x)
#ifdef UAISA_PAIR_UNROLL        // experiments: unroll factor of the stage-1b / stage-2a mask loops (profiles/variants.sh)
#define UAISA_PAIR_UNROLL_PRAGMA UAISA_PRAGMA(unroll UAISA_PAIR_UNROLL)
#else
#define UAISA_PAIR_UNROLL_PRAGMA
#endif
#ifndef UAISA_LB_THREADS
#define UAISA_LB_THREADS MAXT   // experiments: smaller CTAs with a different register cap (profiles/variants.sh)
#endif
#ifndef UAISA_F32_SMALL
#define UAISA_F32_SMALL 0   // experiments: fp32 range pre-gate also in the few-neighbour kernels (MAXR <= 4)
#endif
#ifndef UAISA_MINB128
#define UAISA_MINB128 3   // CTAs of 128 threads per SM the register allocator must leave room for (few-neighbour kernels, MAXR <= 4: 168 registers)
#endif
// The many-neighbour instantiations without helper slices (MAXR > 4: top-nm lists in shared memory) are compiled in a translation unit of their own
// (-DUAISA_TU_DENSE -maxrregcount=128, see the Makefile).  The register file is split over the four schedulers of an SM (16384 each), so a warp of
// 168-register threads leaves room for 3 warps per scheduler = 12 per SM -- and so does any count down to 129; 128 registers make it 16.  For these
// kernels 128 registers cost ~50 bytes of spills (12 % per warp at equal occupancy) but let configs[2]'s 2048 warps run as ONE wave of 64-thread
// CTAs instead of 1.15 waves that cost 2: 1.24e9 -> 1.58e9 agent-steps/s (profiles/wave_quant.py, profiles/variants.sh).  The sliced instantiations
// (configs[3]) measured 4 % slower at 128 and keep 168, like the few-neighbour kernels (200+ bytes of spills at any lower cap).
// (The cap has to come from -maxrregcount ALONE: next to a __launch_bounds__ the same cap spills 190 bytes instead of 16, __maxnreg__ 240.)
// WL ("warp local"): N divides 32, so every env lives inside one warp -- all CTA barriers become __syncwarp(),
// the per-env reductions of the reset policy become ballots, and each warp clears its own 32 output rows.
template <int MAXR, int MAXT, bool WL, bool SL>
#if defined(UAISA_TU_DENSE)   // register cap from -maxrregcount alone (with a __launch_bounds__ next to it the same cap spills 190 bytes instead of 16)
__global__ void uaisa_env_kernel_d(
#elif defined(UAISA_NO_LB)
__global__ void uaisa_env_kernel(
#else
__global__ void __launch_bounds__(UAISA_LB_THREADS, (MAXT <= 128 ? UAISA_MINB128 : (MAXT <= 192 ? 2 : 1))) uaisa_env_kernel(
#endif
    const __grid_constant__ Params P, const __grid_constant__ Launch L) {
#define ENV_SYNC() do { if (WL) __syncwarp(); else __syncthreads(); } while (0)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int threads = P.threads;
    // Cluster mode (P.cl = G > 1, N > 128): one env is spread over a thread-block cluster of G CTAs.  Every CTA keeps a full
    // copy of the env's exchange columns in its own shared memory; at each publish point the CTAs write their slice to a
    // double-buffered global (L2) exchange array, meet at the hardware cluster barrier and reload all N columns.
    const bool CLM = !WL && P.cl > 1;
    // SPLIT (WL && SL, few-neighbour kernels, small batches): the CTA carries a HELPER warp per owner warp.  PASS A (pre-move dronestate, pair pass,
    // rvo reward, and the deferred re-observation row) does not feed the move -- only the reward -- so the helper runs it on the pre-move state the
    // owner publishes while the owner moves and runs PASS B: the dependent chain of a step shrinks from A + MOVE + B to max(A, MOVE + B).  What a
    // latency-bound batch (one warp per scheduler or less) gains is exactly that; a saturating batch would only pay for the second warp.
    constexpr bool SPLIT = WL && SL;
    constexpr bool F32G = (MAXR > 4) || (UAISA_F32_SMALL != 0);                  // stage 1a runs in fp32 on float copies of the positions
    const int sst = CLM ? P.N : threads;                                         // column stride of the exchange tile
    // Shared memory is carved up by byte offsets from the one __shared__ array, never through integer round trips of a pointer, so
    // that every access below compiles to LDS/STS (a pointer rebuilt from a uintptr_t is generic: LD.E on the long scoreboard).
    const int S = SL ? P.slices : 1;                   // neighbour slices per ego: slice 0 = the owner threads, the others are helpers (pair loops only)
    unsigned off = 0;
    auto take = [&](unsigned bytes) { unsigned char *p_ = smem_raw + off; off += bytes; return p_; };
    auto align16 = [&]() { off = (off + 15u) & ~15u; };
    double *sx = reinterpret_cast<double *>(take(8u * 8u * sst));                                    // [8][sst] exchange
    int *senv = reinterpret_cast<int *>(take(4u * P.envs_per_cta));                                  // [envs_per_cta] reduction bits
    int *ssel = reinterpret_cast<int *>(take(4u * P.envs_per_cta));                                  // [envs_per_cta] obs selected
    align16();                                                                                      // constant tiles for the TMA fills (WL only)
    uint4 *tile0 = reinterpret_cast<uint4 *>(take(P.tile0_bytes));                                   // one warp's 32 observation rows of zeros
    uint4 *tile1 = reinterpret_cast<uint4 *>(take(P.tile1_bytes));                                   // 32 rows of nbr = -1
    unsigned char *stop = take((unsigned)(S * TopList<MAXR>::smem_bytes(threads)));                  // top-nm lists of the MAXR > 4 instantiations, one set per slice
    int *s_nt = reinterpret_cast<int *>(take(8u));                                                   // [2] (double-buffered counter of the re-observe compaction)
    align16();
    float4 *sxf = reinterpret_cast<float4 *>(take(16u * sst));                                       // [sst] fp32 copies of the positions
    double *spart_t = reinterpret_cast<double *>(take(8u * (S - 1) * threads));                      // [S-1][threads] helpers' partial tmin
    int *spart_f = reinterpret_cast<int *>(take(4u * (S - 1) * threads));                            // [S-1][threads] helpers' vo_any | col << 1 | cnt << 2
    int *s_act = reinterpret_cast<int *>(take(4u * threads));                                        // [threads] owner is active in the re-observe phase
    align16();
    // Per-drone route cache: current waypoint (3), previous waypoint (3), unit direction + length of the segment between them (4).
    // They are read four times per step and change only at a waypoint switch or a reset; from L1 each read is a ~100-cycle
    // long-scoreboard round trip that small batches (one warp per scheduler) cannot hide.  smap: map size of each env of the CTA.
    double *swp = reinterpret_cast<double *>(take(8u * 10u * threads));                              // [10][threads]
    double *smap = reinterpret_cast<double *>(take(8u * 3u * P.envs_per_cta));                       // [3][envs_per_cta]
    // Cylinders of each env of the CTA ([x, y, h, r] x M, M <= 16): every drone walks them once per step; from L1 each value is a long-scoreboard
    // round trip (4.4 % of configs[2]'s stall samples sat on these loads, profiles/r2/ncu_env_c3.txt).
    const int Mc = P.bld_cache ? P.M : 0;          // (host: some scenario has cylinders and M <= 16)
    double *sbld = reinterpret_cast<double *>(take(8u * 4u * (unsigned)Mc * P.envs_per_cta));        // [envs_per_cta][Mc][4]
    // CTA-level compaction of the re-observe phase: parked owner state + list of touched envs.  Reserved only when the geometry can compact, and
    // LAST in the carve-up: with the conditional size in the middle of it the few-neighbour kernels measured 5 % slower (same work, different
    // register allocation around the offsets that follow).
    double *spk_d = reinterpret_cast<double *>(take(0u));                                            // [4][threads] yaw pitch real_len max_dev
    int *spk_i = reinterpret_cast<int *>(spk_d + 4 * threads);                                       // [3][threads] wp index, flags, staged
    int *s_list = spk_i + 3 * threads;                                                               // [envs_per_cta]
    // SPLIT (never together with the compaction): the pre-move exchange tile the helpers read while the owners republish `sx`, the helpers' own
    // route cache, and the owner <-> helper mailboxes (rvo reward and deviation of PASS A; waypoint index and rows staged by the previous PASS B)
    double *sxA = sx, *swp_h = swp, *s_rA = nullptr, *s_dvA = nullptr; int *s_wpi = nullptr, *s_stg = nullptr;
    if constexpr (SPLIT) {
        sxA = reinterpret_cast<double *>(take(8u * 8u * threads)); swp_h = reinterpret_cast<double *>(take(8u * 10u * threads));
        s_rA = reinterpret_cast<double *>(take(8u * threads)); s_dvA = reinterpret_cast<double *>(take(8u * threads));
        s_wpi = reinterpret_cast<int *>(take(4u * threads)); s_stg = reinterpret_cast<int *>(take(4u * threads));
    }
    {
        const bool clm_ = !WL && P.cl > 1;
        const int e0_ = clm_ ? (int)(blockIdx.x / P.cl) : blockIdx.x * P.envs_per_cta;
        for (int i = threadIdx.x; i < P.envs_per_cta * Mc * 4; i += blockDim.x) {
            const int le_ = i / (Mc * 4), e_ = e0_ + le_;
            sbld[i] = e_ < P.E ? __ldg(P.bld + (size_t)__ldg(P.env2scn + e_) * P.M * 4 + (i - le_ * Mc * 4)) : 0.0;
        }
    }
    if (threadIdx.x == 0) { s_nt[0] = 0; s_nt[1] = 0; }
    const float nbr2_lo_f = (float)(P.nbr_range * P.nbr_range * (1.0 - 1e-4)), nbr2_hi_f = (float)(P.nbr_range * P.nbr_range * (1.0 + 1e-4));
    if (WL) {
        for (int i = threadIdx.x; i < P.tile0_bytes / 16; i += blockDim.x) tile0[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < P.tile1_bytes / 16; i += blockDim.x) tile1[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the copy engine
    }
    __syncthreads();

    const int tix = threadIdx.x, slice = (S > 1) ? tix / threads : 0, tid = tix - slice * threads, N = P.N, lane = tix & 31;
    const bool helper = slice > 0;                     // whole warps (threads % 32 == 0)
    const int crank = CLM ? (int)(blockIdx.x % P.cl) : 0;
    const unsigned env_lanes = WL ? (N >= 32 ? 0xffffffffu : (((1u << N) - 1u) << (lane - lane % N))) : 0u;
    const int le = CLM ? 0 : tid / N, k = CLM ? crank * threads + tid : tid - le * N;
    const int env0 = CLM ? (int)(blockIdx.x / P.cl) : blockIdx.x * P.envs_per_cta;
    const int env = env0 + le;
    const bool evalid = CLM ? (k < N) : ((le < P.envs_per_cta) && (env < P.E));   // my ego exists
    const bool valid = evalid && !helper;              // ... and I own it: all state, flags and outputs are the owner's business
    const int lbase = le * N;
    const int A = P.E * N;
    const int a = evalid ? env * N + k : 0;
    // The first step's action is requested before anything else: in the host round trip it comes over PCIe (mapped pinned memory, ~3 us) and that
    // latency then overlaps the prologue's dependent loads (scenario constants, state, route cache) instead of following them.
    float p32x = 0.f, p32y = 0.f, p32z = 0.f;
    if (L.mode == MODE_STEP && evalid) { const float *ap = L.action + (size_t)a * 3; p32x = ap[0]; p32y = ap[1]; p32z = ap[2]; }
    const int scol = CLM ? k : tid;                               // my column in the exchange tile
    const int ag0 = CLM ? env0 * N + crank * threads : env0 * N;  // first agent of this CTA (its agents are contiguous)
    const int nag_here = CLM ? max(0, min(threads, N - crank * threads)) : min(P.envs_per_cta, P.E - env0) * N;
    int pubn = 0;                                                 // publish counter (selects the global exchange buffer)
    const int cap = min(P.nm, MAXR);
    const bool is_step = L.mode == MODE_STEP;
    const bool do_auto = (L.mode == MODE_AUTORESET) || (is_step && L.auto_reset);
    const bool compact = is_step && do_auto && P.compact;   // re-observe through CTA-level compaction (uniform)

    RG_DECL
    if (L.hp_self && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); L.hp_hdr[1] = g_; }   // host-step profiling: first CTA's start
    ConstsT<(MAXR <= 4)> C; Drone d; int ep_len = 0;
#define FILL_WP(col) do { const double *w_ = C.wp(P) + 3 * d.i, *wp_ = C.wp(P) + 3 * max(d.i - 1, 0), *sg_ = C.seg(P) + 4 * d.i; \
        swp[0 * threads + (col)] = __ldg(w_); swp[1 * threads + (col)] = __ldg(w_ + 1); swp[2 * threads + (col)] = __ldg(w_ + 2); \
        swp[3 * threads + (col)] = __ldg(wp_); swp[4 * threads + (col)] = __ldg(wp_ + 1); swp[5 * threads + (col)] = __ldg(wp_ + 2); \
        swp[6 * threads + (col)] = __ldg(sg_); swp[7 * threads + (col)] = __ldg(sg_ + 1); swp[8 * threads + (col)] = __ldg(sg_ + 2); swp[9 * threads + (col)] = __ldg(sg_ + 3); } while (0)
    const double *bld = nullptr; int nb = 0, scn = 0;
    if (valid) {
        scn = __ldg(P.env2scn + env);
        const int sn = scn * N + k;
        C.set(P, sn); C.radius = __ldg(P.radius + sn); C.prio = __ldg(P.prio + sn);
        C.route_len = __ldg(P.route_len + sn); C.n = __ldg(P.npts + sn);
        bld = P.bld + (size_t)scn * P.M * 4; nb = __ldg(P.nbld + scn);
        const float4 py = P.pos_yaw[a], vp = P.vel_pitch[a];
        d.px = py.x; d.py = py.y; d.pz = py.z; d.yaw = py.w; d.vx = vp.x; d.vy = vp.y; d.vz = vp.z; d.pitch = vp.w;
        d.real_len = P.real_len[a]; d.max_dev = P.max_dev[a];
        const int wf = P.wpf[a]; d.i = wf & WP_MASK; d.flags = wf & (FLAG_ARRIVE | FLAG_DEST);
        if (do_auto) ep_len = P.ep_len[a];
        FILL_WP(tid);
        if (k == 0 || CLM) { const int le_ = CLM ? 0 : le; smap[0 * P.envs_per_cta + le_] = __ldg(P.map + 3 * scn); smap[1 * P.envs_per_cta + le_] = __ldg(P.map + 3 * scn + 1); smap[2 * P.envs_per_cta + le_] = __ldg(P.map + 3 * scn + 2); }
        if (!CLM) { sx[6 * sst + scol] = C.radius; sx[7 * sst + scol] = C.prio; }      // constant over the launch
        if constexpr (SPLIT) { sxA[6 * sst + scol] = C.radius; sxA[7 * sst + scol] = C.prio; }
    }
    int h_i = -1;                                      // SPLIT helper: waypoint index its route cache holds
    if (SPLIT && helper && evalid) { scn = __ldg(P.env2scn + env); C.set(P, scn * N + k); }
    if (CLM) {   // every CTA of the cluster needs the constants of ALL drones of the env
        const int s_ = __ldg(P.env2scn + env0);
        for (int j = tix; j < N; j += blockDim.x) { sx[6 * sst + j] = __ldg(P.radius + s_ * N + j); sx[7 * sst + j] = __ldg(P.prio + s_ * N + j); }
    }
    unsigned n_rng = 0, n_cone = 0, n_bad = 0;
    const int T = is_step ? L.T : 1;
#define HP_STAMP(slot) do { if (L.hp_self && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); L.hp_hdr[slot] = g_; } } while (0)
    HP_STAMP(3);   // host-step profiling: CTA 0's prologue done
    // Deferred re-observation (T-step launches): the re-observe phase of step t < T-1 would publish the state, run dronestate and the
    // stage-1 gates on exactly what PASS A of step t+1 publishes and gates again (same fp32-rounded state, same waypoint).  So the phase
    // is not run: the reset policy only marks the env (`pend`), and PASS A of the next step also evaluates the cones with action 0, builds
    // the top-nm list and writes the observation row into the PREVIOUS step's slot.  The last step of a launch re-observes as before.
    const bool defer_on = is_step && do_auto && P.defer_obs;
    bool pend = false;                                            // my env's row of the previous step is still to be rewritten
    int staged = -1;                                              // rows of mine already in HBM this step (-1: none, row not cleared)
    RG_MARK(RG_PROLOGUE);

    for (int t = 0; t < T; t++) {
#define OUT_SLOT(agent) ((size_t)t * A + (size_t)(agent))   // slot in the [T][A] outputs (recomputed at the use sites: no live 64-bit pointers)
        float a32x = 0.f, a32y = 0.f, a32z = 0.f;        // issue the action loads ahead of the burst of clearing stores
        if (t == 0) { a32x = p32x; a32y = p32y; a32z = p32z; }
        if (is_step && evalid) {
            const float *ap = L.action + OUT_SLOT(a) * 3;
            if (t > 0) { a32x = ap[0]; a32y = ap[1]; a32z = ap[2]; }
            if (t + 1 < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + (size_t)A * 3));   // next step's action
        }
        if (!WL && tix < P.envs_per_cta) { senv[tix] = 0; ssel[tix] = 0; }
        const bool clear_rows = is_step && !L.hp_self;   // (the host round trip reads back only the rows that exist: its staging buffers are not cleared)
        if (clear_rows) {   // every row is written this step: clear the output region once, coalesced 16-byte stores
            if (WL) {    // my warp's 32 rows
                const int w0 = tid - lane, nw = (SPLIT ? !helper : helper) ? 0 : min(32, nag_here - w0);      // (SPLIT: the helper warp clears its owner's rows)
                if (nw > 0) {
                    const size_t abase = (size_t)t * A + (size_t)env0 * N + w0;
                    if (L.obs) {
                        char *dst = reinterpret_cast<char *>(L.obs + abase * P.obs_dim); const int bytes = nw * P.obs_dim * 4;
                        if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)bytes) & 15) == 0) { if (lane == 0) bulk_fill(dst, tile0, P.tile0_bytes, bytes); }
                        else fill_out(reinterpret_cast<uint32_t *>(dst), bytes >> 2, 0u, lane, 32);
                    }
                    if (L.nbr) {
                        char *dst = reinterpret_cast<char *>(L.nbr + abase * P.nm); const int bytes = nw * P.nm * 4;
                        if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)bytes) & 15) == 0) { if (lane == 0) bulk_fill(dst, tile1, P.tile1_bytes, bytes); }
                        else fill_out(reinterpret_cast<uint32_t *>(dst), bytes >> 2, ~0u, lane, 32);
                    }
                    if (lane == 0) bulk_commit();
                    if constexpr (SPLIT) __syncwarp();       // (fallback stores of the other lanes are ordered before lane 0 raises the flag)
                }
            } else {
                const size_t abase = (size_t)t * A + (size_t)ag0;
                if (L.obs) fill_out(reinterpret_cast<uint32_t *>(L.obs + abase * P.obs_dim), nag_here * P.obs_dim, 0u, tix, blockDim.x);
                if (L.nbr) fill_out(reinterpret_cast<uint32_t *>(L.nbr + abase * P.nm), nag_here * P.nm, ~0u, tix, blockDim.x);
            }
        }
        RG_MARK(RG_FILL);
        const double ax = a32x, ay = a32y, az = a32z; double rA = 0;
#ifdef UAISA_REGION_CLOCKS
        if (ax + ay + az == 123456.0) n_bad++;   // force the loads to complete here
#endif
        RG_MARK(RG_ACT);
        bool st_done = false, st_fin = false;
        const int staged_prev = staged;                  // (what PASS B of the previous step left in ITS slot: a deferred re-observation clears the rest)
        staged = -1;
        // "working identity": normally the thread's own drone; in a compacted re-observe phase the drone it was handed
        int wk = k, wlbase = lbase, tt = tid, wa = a;
        bool remapped = false, worker = false;

        for (int ph = is_step ? PH_A : PH_OBS; ph <= PH_OBS; ph++) {
            if (SPLIT && helper && ph != PH_A) break;      // a helper only runs PASS A
            bool active = valid;
            RG_MARK(RG_TAIL_B);
            if (ph == PH_OBS) {
                // ------------------------------------------------ who gets (re-)observed
                bool sel = false;      // my env is (re-)observed
                bool rs = false;       // I am reset
                if (L.mode == MODE_RESET) {
                    rs = valid && (L.mask == nullptr || L.mask[a]);
                    if (WL) sel = (__ballot_sync(0xffffffffu, rs) & env_lanes) != 0;
                    else if (CLM) {
                        int *ge = P.genv + env0;
                        if (rs) atomicOr(ge, 1);
                        cluster_sync_all();
                        sel = valid && (__ldcg(ge) != 0);
                        cluster_sync_all();
                        if (crank == 0 && tix == 0) *ge = 0;
                    }
                    else { __syncthreads(); if (rs) ssel[le] = 1; __syncthreads(); sel = valid && ssel[le]; }   // benign race: all write 1
                } else if (L.mode == MODE_OBSERVE) {
                    sel = valid && (L.mask == nullptr || L.mask[env]);
                } else if (L.mode == MODE_AUTORESET) {
                    if (valid) { st_done = L.done_in[a] != 0; st_fin = L.finish_in[a] != 0; }
                }
                if (do_auto) {   // trainer reset policy, multi_ppo.py:221-281
                    if (valid) ep_len += 1;
                    const bool over = valid && ep_len > L.max_ep_len;
                    int bits;
                    if (WL) {
                        const unsigned bd = __ballot_sync(0xffffffffu, valid && st_done), bf = __ballot_sync(0xffffffffu, valid && st_fin);
                        const unsigned bn = __ballot_sync(0xffffffffu, valid && !st_fin), bo = __ballot_sync(0xffffffffu, over);
                        bits = ((bd & env_lanes) ? 1 : 0) | ((bf & env_lanes) ? 2 : 0) | ((bn & env_lanes) ? 4 : 0) | ((bo & env_lanes) ? 8 : 0);
                    } else if (CLM) {
                        int *ge = P.genv + env0;
                        if (valid) atomicOr(ge, (st_done ? 1 : 0) | (st_fin ? 2 : 4) | (over ? 8 : 0));
                        cluster_sync_all();
                        bits = valid ? __ldcg(ge) : 0;
                        cluster_sync_all();                                    // everybody has read the bits: re-arm the slot
                        if (crank == 0 && tix == 0) *ge = 0;
                    } else {
                        __syncthreads();
                        if (valid) atomicOr(&senv[le], (st_done ? 1 : 0) | (st_fin ? 2 : 4) | (over ? 8 : 0));
                        __syncthreads();
                        bits = valid ? senv[le] : 0;
                    }
                    if (valid) {
                        const bool all_fin = !(bits & 4), terminal = (bits & 2) || (bits & 8);
                        bool cut = false;
                        if (bits & 1) { if (st_done) { rs = true; ep_len = 0; } sel = true; }                   // :230-242
                        if (L.epoch_ended || all_fin) { rs = true; ep_len = 0; cut = true; sel = true; }        // :244-264
                        else if (terminal) { if (st_fin || ep_len > L.max_ep_len) { rs = true; ep_len = 0; } cut = true; sel = true; }   // :266-281
                        if (L.reset_mask) L.reset_mask[OUT_SLOT(a)] = rs;
                        if (L.path_cut) L.path_cut[OUT_SLOT(a)] = cut;
                    }
                } else if (is_step) {
                    break;       // plain step: nothing to re-observe (uniform over the CTA)
                }
                if (rs) { reset_drone(P, d, C); FILL_WP(tid); }
                if (rs && L.mode == MODE_RESET) P.ep_len[a] = 0;      // a reset starts a new episode (multi_ppo.py:168,238,261: ep_len_list[i] = 0)
                active = valid && sel && (L.obs || L.count || L.nbr);
                if (S > 1 && !helper) s_act[tid] = active;             // read by my helpers after the barriers below
                RG_MARK(RG_AUTORESET);
                pend = false;
                if (defer_on && t + 1 < T) { pend = active; continue; }   // (uniform) re-observed inside PASS A of the next step
            }
            // a warp none of whose envs is re-observed has nothing to publish or read (warp-local envs only)
            if (WL && !compact && ph == PH_OBS && !__any_sync(0xffffffffu, active)) continue;      // (the compacted variant meets at CTA barriers)
            // -------------------------------------------------------- publish my state, then read the others'
            if (WL && !SPLIT && clear_rows && ph == PH_B && lane == 0) bulk_wait_all();   // my warp's cleared rows have landed: rows may now be scattered
            const bool DSM = CLM && P.dsmem;
            if (DSM) cluster_sync_all(); else ENV_SYNC();   // everybody (of the cluster / the CTA) is done reading the previous phase's columns
            if (DSM) {
                // Distributed shared memory: every owner stores its six values (+ the fp32 copy) straight into the exchange tile of
                // each CTA of the cluster (st.shared::cluster through mapa); the cluster barrier below makes them visible.  No trip
                // through L2, no reload loop.
                if (valid) {
                    const uint32_t sx_a = (uint32_t)__cvta_generic_to_shared(sx) + 8u * (uint32_t)k, sxf_a = (uint32_t)__cvta_generic_to_shared(sxf) + 16u * (uint32_t)k;
                    const uint32_t fs = 8u * (uint32_t)sst;
                    for (int r = 0; r < P.cl; r++) {
                        const uint32_t b_ = mapa_shared(sx_a, (uint32_t)r);
                        st_cluster_f64(b_, d.px); st_cluster_f64(b_ + fs, d.py); st_cluster_f64(b_ + 2 * fs, d.pz);
                        st_cluster_f64(b_ + 3 * fs, d.vx); st_cluster_f64(b_ + 4 * fs, d.vy); st_cluster_f64(b_ + 5 * fs, d.vz);
                        if constexpr (F32G) st_cluster_v4f32(mapa_shared(sxf_a, (uint32_t)r), (float)d.px, (float)d.py, (float)d.pz, 0.f);
                    }
                }
            } else if (CLM) {
                double *gb = P.gx + ((size_t)(pubn & 1) * A + (size_t)env0 * N) * 6;      // this env's slice of the exchange buffer
                if (valid) { double *g = gb + (size_t)k * 6; g[0] = d.px; g[1] = d.py; g[2] = d.pz; g[3] = d.vx; g[4] = d.vy; g[5] = d.vz; }
                pubn++;
                cluster_sync_all();                        // every CTA of the env has written its slice (release/acquire at cluster scope)
                for (int j = tix; j < N; j += blockDim.x) {
                    const double *g = gb + (size_t)j * 6;  // L1 is not coherent across SMs: read through L2
                    const double x_ = __ldcg(g), y_ = __ldcg(g + 1), z_ = __ldcg(g + 2);
                    sx[0 * sst + j] = x_; sx[1 * sst + j] = y_; sx[2 * sst + j] = z_;
                    sx[3 * sst + j] = __ldcg(g + 3); sx[4 * sst + j] = __ldcg(g + 4); sx[5 * sst + j] = __ldcg(g + 5);
                    if constexpr (F32G) sxf[j] = make_float4((float)x_, (float)y_, (float)z_, 0.f);
                }
            } else if (valid) {
                double *sxw = (SPLIT && ph == PH_A) ? sxA : sx;
                sxw[0 * sst + scol] = d.px; sxw[1 * sst + scol] = d.py; sxw[2 * sst + scol] = d.pz;
                sxw[3 * sst + scol] = d.vx; sxw[4 * sst + scol] = d.vy; sxw[5 * sst + scol] = d.vz;
                if constexpr (F32G) sxf[scol] = make_float4((float)d.px, (float)d.py, (float)d.pz, 0.f);
                if constexpr (SPLIT) { if (ph == PH_A) { s_wpi[tid] = d.i; s_stg[tid] = staged_prev; } }
            }
            if (ph == PH_OBS && compact) {
                // Only the envs touched by the reset policy are re-observed (typically a quarter of them, spread over
                // nearly every warp).  Gather them: owners park their state in shared memory, the touched envs are listed,
                // and the first nt*N threads of the CTA each take over one drone of a listed env, so whole warps are
                // either fully busy or skip the phase.
                if (valid) {
                    spk_d[0 * threads + tid] = d.yaw; spk_d[1 * threads + tid] = d.pitch; spk_d[2 * threads + tid] = d.real_len; spk_d[3 * threads + tid] = d.max_dev;
                    spk_i[0 * threads + tid] = d.i; spk_i[1 * threads + tid] = d.flags; spk_i[2 * threads + tid] = staged;
                }
                if (active && k == 0) s_list[atomicAdd(&s_nt[t & 1], 1)] = le;
                __syncthreads();
                const int nt = s_nt[t & 1];
                if (tix == 0) s_nt[(t + 1) & 1] = 0;
                remapped = true;
                worker = active = tid < nt * N;
                if (worker) {
                    const int q = tid / N, e_l = s_list[q];
                    wk = tid - q * N; wlbase = e_l * N; tt = wlbase + wk;
                    const int env_t = env0 + e_l, sn_t = __ldg(P.env2scn + env_t) * N + wk;
                    d.px = sx[0 * threads + tt]; d.py = sx[1 * threads + tt]; d.pz = sx[2 * threads + tt];
                    d.vx = sx[3 * threads + tt]; d.vy = sx[4 * threads + tt]; d.vz = sx[5 * threads + tt];
                    d.i = spk_i[0 * threads + tt]; d.max_dev = spk_d[3 * threads + tt]; staged = spk_i[2 * threads + tt];
                    C.set(P, sn_t);
                    C.radius = sx[6 * threads + tt]; C.prio = sx[7 * threads + tt];
                    wa = env_t * N + wk;
                }
            } else if (SPLIT && ph == PH_A) {
                asm volatile("barrier.sync 0;" ::: "memory");    // SPLIT sync #1: the pre-move tile is complete, the helpers start PASS A on it
            } else {
                if (DSM) cluster_sync_all(); else ENV_SYNC();   // all columns are in place
            }
            RG_MARK(RG_SYNC);
            // helper threads (slices > 1) walk their share of the ego's neighbours whenever its owner is active
            const bool hact = S > 1 && helper && evalid && (ph != PH_OBS || s_act[tid] != 0);
            // (threads with nothing to do still meet the others at the merge barrier below: barrier.sync without .aligned
            // counts thread arrivals on barrier 0 wherever the instruction sits)
            if (!active && !hact) {
                if (S > 1 && !SPLIT) asm volatile("barrier.sync 0;" ::: "memory");
                if (SPLIT && ((helper && ph == PH_A) || (!helper && ph == PH_B))) asm volatile("barrier.sync 0;" ::: "memory");      // idle lanes still arrive at SPLIT sync #2
                continue;
            }
            const double *sxr = (SPLIT && ph == PH_A) ? sxA : sx;       // the tile this pass reads
            const bool own_a_skip = SPLIT && !helper && ph == PH_A;    // SPLIT owner in PASS A: publish, then straight to the move

            // drone.dronestate (drone.py:254-263): desired velocity + deviation, side effect on max_deviation
            double d0 = 0, d1 = 0, d2 = 0, dev = 0;
            int staged_a = staged_prev;                        // rows PASS B of the previous step left in its slot (SPLIT helper: from the owner's mailbox)
            if (own_a_skip) {
            } else if (SPLIT && helper) {   // the ego as its owner just published it; my own route cache follows the published waypoint index
                d.px = sxA[0 * sst + scol]; d.py = sxA[1 * sst + scol]; d.pz = sxA[2 * sst + scol];
                d.vx = sxA[3 * sst + scol]; d.vy = sxA[4 * sst + scol]; d.vz = sxA[5 * sst + scol];
                C.radius = sxA[6 * sst + scol]; C.prio = sxA[7 * sst + scol];
                staged_a = s_stg[tid];
                const int wi = s_wpi[tid];
                if (wi != h_i) {
                    h_i = wi;
                    const double *w_ = C.wp(P) + 3 * wi, *wp_ = C.wp(P) + 3 * max(wi - 1, 0), *sg_ = C.seg(P) + 4 * wi;
                    swp_h[0 * threads + tid] = __ldg(w_); swp_h[1 * threads + tid] = __ldg(w_ + 1); swp_h[2 * threads + tid] = __ldg(w_ + 2);
                    swp_h[3 * threads + tid] = __ldg(wp_); swp_h[4 * threads + tid] = __ldg(wp_ + 1); swp_h[5 * threads + tid] = __ldg(wp_ + 2);
                    swp_h[6 * threads + tid] = __ldg(sg_); swp_h[7 * threads + tid] = __ldg(sg_ + 1); swp_h[8 * threads + tid] = __ldg(sg_ + 2); swp_h[9 * threads + tid] = __ldg(sg_ + 3);
                }
                cal_des_vel_v(P, d.px, d.py, d.pz, swp_h[0 * threads + tid], swp_h[1 * threads + tid], swp_h[2 * threads + tid], d0, d1, d2);
                dev = deviation(swp_h[3 * threads + tid], swp_h[4 * threads + tid], swp_h[5 * threads + tid], swp_h[6 * threads + tid], swp_h[7 * threads + tid],
                                swp_h[8 * threads + tid], swp_h[9 * threads + tid], d.px, d.py, d.pz);
            } else if (!SL || !helper) {
                cal_des_vel_v(P, d.px, d.py, d.pz, swp[0 * threads + tt], swp[1 * threads + tt], swp[2 * threads + tt], d0, d1, d2);
                dev = deviation(swp[3 * threads + tt], swp[4 * threads + tt], swp[5 * threads + tt], swp[6 * threads + tt], swp[7 * threads + tt],
                                swp[8 * threads + tt], swp[9 * threads + tt], d.px, d.py, d.pz);
                if (dev > d.max_dev) d.max_dev = dev;
            } else {   // the ego as its owner just published it
                d.px = sx[0 * sst + scol]; d.py = sx[1 * sst + scol]; d.pz = sx[2 * sst + scol];
                d.vx = sx[3 * sst + scol]; d.vy = sx[4 * sst + scol]; d.vz = sx[5 * sst + scol];
                C.radius = sx[6 * sst + scol]; C.prio = sx[7 * sst + scol];
            }
            RG_MARK(RG_DRONESTATE);
            // pairwise pass over the drones of my env (rvo_inter.py:20-83)
            bool vo_any = false, col = false; double tmin = CUDART_INF;
            TopList<MAXR> top; top.init(stop + (unsigned)(slice * TopList<MAXR>::smem_bytes(threads)), threads, tid); top.clear();
            int jb_lo = 0, jb_hi = N;                          // my slice: whole 32-neighbour chunks
            if (S > 1 && !SPLIT) { const int nch = (N + 31) >> 5; jb_lo = ((slice * nch) / S) << 5; jb_hi = min(N, (((slice + 1) * nch) / S) << 5); }
            if (own_a_skip) jb_hi = jb_lo;                     // (the helper walks the neighbours of PASS A)
            // evaluations of the cones in this pass: ev 1 = the phase's own (the action; 0 when re-observing), ev 0 = the re-observation
            // deferred from the previous step (action 0, list only) -- both on the same stage-1 masks
            const bool ev0 = (ph == PH_A) && !own_a_skip && ((S > 1 && helper) ? (defer_on && t > 0 && s_act[tid] != 0) : pend);
            {
                double bx = (ph == PH_OBS) ? 0.0 : ax, by = (ph == PH_OBS) ? 0.0 : ay, bz = (ph == PH_OBS) ? 0.0 : az;
                if (np_dot3(bx, by, bz, bx, by, bz) < 1e-10 * (1.0 - 1e-14)) { bx = by = bz = 0.0; }   // norm(a) < 1e-5, rvo_inter.py:118-119
                // Stage 1 runs over a chunk of up to 32 neighbours with no divergent work and yields a mask of the
                // pairs whose cone must be evaluated; stage 2 walks that mask (rvo_inter.py:29,67 map both).
                for (int jb = jb_lo; jb < jb_hi; jb += 32) {
                    unsigned cand = 0, slow = 0, near = 0;
                    const int jn = min(32, jb_hi - jb);
                    if constexpr (F32G) {                         // fp32 pre-gate on float copies of the positions
                        const float fx = (float)d.px, fy = (float)d.py, fz = (float)d.pz;
                        if (jn == 32) {                          // full chunk: fixed trip count, fully unrolled -> 32 independent gates in flight
#pragma unroll
                            for (int jj = 0; jj < 32; jj++) {    // straight-line: "self" is gated like everybody else and masked out below
                                const int bits = pair_range_f32(fx, fy, fz, sxf[wlbase + jb + jj], nbr2_lo_f, nbr2_hi_f);
                                near |= (bits & PB_EVAL) ? (1u << jj) : 0u;
                                slow |= (bits & PB_SLOW) ? (1u << jj) : 0u;
                            }
                        } else {
#pragma unroll 8
                            for (int jj = 0; jj < jn; jj++) {    // stage 1a: who is within range (fp32, branch-free, overlapped)
                                const int bits = pair_range_f32(fx, fy, fz, sxf[wlbase + jb + jj], nbr2_lo_f, nbr2_hi_f);
                                near |= (bits & PB_EVAL) ? (1u << jj) : 0u;
                                slow |= (bits & PB_SLOW) ? (1u << jj) : 0u;
                            }
                        }
                        const unsigned notself = ((unsigned)(wk - jb) < 32u) ? ~(1u << (wk - jb)) : ~0u;
                        near &= notself; slow &= notself;
                    } else {                                      // a handful of neighbours: the fp64 gate directly
#pragma unroll 4
                        for (int jj = 0; jj < jn; jj++) {        // straight-line like the fp32 gates: +1 gate (self), no branch (+6..9 % per step)
                            const int bits = pair_range_fast(P, d.px, d.py, d.pz, sxr + wlbase + jb + jj, sst);
                            near |= (bits & PB_EVAL) ? (1u << jj) : 0u;
                            slow |= (bits & PB_SLOW) ? (1u << jj) : 0u;
                        }
                        { const unsigned notself = ((unsigned)(wk - jb) < 32u) ? ~(1u << (wk - jb)) : ~0u; near &= notself; slow &= notself; }
                    }
                    if ((near | slow) == 0u) continue;          // nobody in range in this chunk (the common case in sparse airspace): one branch instead of five empty loops
                    UAISA_PAIR_UNROLL_PRAGMA
                    for (unsigned m = near; m; m &= m - 1) {     // stage 1b: collision / moving-away gates of the near ones (fp64)
                        const int jj = __ffs(m) - 1;
                        const int bits = pair_near_fast(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, sxr + wlbase + jb + jj, sst);
                        n_rng += (bits & PB_SAME) ? 0u : 1u;
                        col |= (bits & PB_COL) != 0;
                        cand |= (bits & PB_FULL) ? (1u << jj) : 0u;
                        slow |= (bits & PB_SLOW) ? (1u << jj) : 0u;
                    }
                    for (unsigned m = slow; m; m &= m - 1) {     // a squared distance inside a guard band: exact comparisons
                        const int jj = __ffs(m) - 1;
                        const int bits = pair_gate(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, sxr + wlbase + jb + jj, sst);
                        n_rng += ((bits & PB_EVAL) && !((near >> jj) & 1u)) ? 1u : 0u;   // (pairs already counted in stage 1b are not counted twice)
                        col |= (bits & PB_COL) != 0;
                        cand |= (bits & PB_FULL) ? (1u << jj) : 0u;
                    }
                    if (cand == 0u) continue;
                    for (int ev = ev0 ? 0 : 1; ev < 2; ev++) {
                    const double ex_ = ev ? bx : 0.0, ey_ = ev ? by : 0.0, ez_ = ev ? bz : 0.0;
                    n_cone += __popc(cand);
                    // stage 2a: trig-free screen of every candidate (uniform work) -> inside / ambiguous masks
                    unsigned ins = 0, amb = 0;
                    UAISA_PAIR_UNROLL_PRAGMA
                    for (unsigned m = cand; m; m &= m - 1) {
                        const int jj = __ffs(m) - 1;
                        PairGeom g; pair_geom(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio, sxr + wlbase + jb + jj, sst, ex_, ey_, ez_, g);
                        const int c = cone_screen(g);
                        ins |= (c == 1) ? (1u << jj) : 0u; amb |= (c == 2) ? (1u << jj) : 0u;
                    }
                    // stage 2b: the few ambiguous pairs go through asin/acos with the reference's roundings
                    for (unsigned m = amb; m; m &= m - 1) {
                        const int jj = __ffs(m) - 1;
                        PairGeom g; pair_geom(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio, sxr + wlbase + jb + jj, sst, ex_, ey_, ez_, g);
                        double ah;
                        if (cone_exact(g, ah)) ins |= 1u << jj;
                    }
                    // stage 2c: pairs inside the cone -> expected collision time, vo_flag, row keys
                    for (unsigned m = ins; m; m &= m - 1) {
                        const int jj = __ffs(m) - 1, j = jb + jj;
                        const double *so = sxr + wlbase + j;
                        PairGeom g; pair_geom(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio, so, sst, ex_, ey_, ez_, g);
                        double tt, md, iet;
                        if (cone_time(P, d.vx, d.vy, d.vz, C.radius, so, sst, ex_, ey_, ez_, g, tt, md, iet)) {
                            if (ev) { vo_any = true; if (tt < tmin) tmin = tt; }
                            if (ph != PH_A || ev == 0) {
                                const double ah = py_round2_hundredths(rare_asin(g.R / sqrt(g.s2)));   // get_alpha, vel_obs3D.py:8-17
                                top.insert(cap, iet, md, j, (int)ah);
                            }
                        }
                    }
                    }
                }
            }
            RG_MARK(RG_PAIRS);
            if constexpr (SPLIT) {
                // The rows the helper warp cleared at the top of the step must have landed before its owner warp scatters into them: lane 0 of the
                // helper (the lane that issued the bulk stores) raises a per-warp-pair flag in shared memory, the owner's lanes poll it.  (A named
                // barrier with arrive / sync would do, but a NON-ALIGNED barrier.arrive executed by a diverged warp -- a partly filled last warp --
                // faulted with "illegal instruction" on sm_100a.)
                volatile int *fd = s_nt + (tid >> 5);
                if (clear_rows && helper) {
                    if (lane == 0) { bulk_wait_all(); __threadfence_block(); *fd = t + 1; }
                } else if (clear_rows && ph == PH_B) {
                    while (*fd < t + 1) { }
                    __threadfence_block();
                }
            }
            if (S > 1) {       // helpers hand their partial results to the owner: flags, earliest collision time, their top-nm lists
                if constexpr (MAXR > 4) {
                    if (helper) { spart_t[(slice - 1) * threads + tid] = tmin; spart_f[(slice - 1) * threads + tid] = (vo_any ? 1 : 0) | (col ? 2 : 0) | (top.cnt << 2); }
                    asm volatile("barrier.sync 0;" ::: "memory");
                    if (helper) continue;
                    for (int sl = 1; sl < S; sl++) {
                        const int f = spart_f[(sl - 1) * threads + tid]; const double pt = spart_t[(sl - 1) * threads + tid];
                        vo_any |= (f & 1) != 0; col |= (f & 2) != 0; if (pt < tmin) tmin = pt;
                        TopList<MAXR> oth; oth.init(stop + (unsigned)(sl * TopList<MAXR>::smem_bytes(threads)), threads, tid); oth.cnt = f >> 2;
                        oth.for_each([&](int, double o_iet, double o_md, int o_j, int o_ah) { top.insert(cap, o_iet, o_md, o_j, o_ah); });   // (iet, md, j) is a total order
                    }
                }
            }
            // ------------------------------------------------------------ observation row (ir_gym.py:203-229)
            if (ph != PH_A || ev0) {    // (PASS A: the re-observation deferred from the previous step, into THAT step's slot)
                bool bad = false;
                const size_t wo = (ph == PH_A) ? OUT_SLOT(wa) - (size_t)A : OUT_SLOT(wa);
                const int staged_in = (ph == PH_A) ? staged_a : staged;
                float *const wobs = L.obs ? L.obs + wo * P.obs_dim : nullptr; int *const wnbr = L.nbr ? L.nbr + wo * P.nm : nullptr;
                if (wobs) {
                    const double self[OBS_SELF] = {d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio, d0, d1, d2, dev};
                    float fs[OBS_SELF];
#pragma unroll
                    for (int q = 0; q < OBS_SELF; q++) fs[q] = np_round2_f32(self[q]);
                    float chk = fs[0];
#pragma unroll
                    for (int q = 1; q < OBS_SELF; q++) chk += fs[q];            // NaN or Inf iff some value is (finite values of this size cannot overflow)
                    bad |= (chk * 0.f != 0.f);
                    if ((P.obs_dim & 1) == 0) {
#pragma unroll
                        for (int q = 0; q < OBS_SELF; q += 2) reinterpret_cast<float2 *>(wobs)[q >> 1] = make_float2(fs[q], fs[q + 1]);
                    } else {
#pragma unroll
                        for (int q = 0; q < OBS_SELF; q++) wobs[q] = fs[q];
                    }
                }
                // slot r (least urgent first, most urgent last) = top[cnt-1-r]
                top.for_each([&](int sl, double t_iet, double t_md, int t_j, int t_ah) {
                    const int r = top.cnt - 1 - sl;
                    const double *so = sxr + wlbase + t_j;
                    if (wobs) {
                        const double mx = so[0], my = so[sst], mz = so[2 * sst];
                        const double mvx = so[3 * sst], mvy = so[4 * sst], mvz = so[5 * sst], prb = so[7 * sst];
                        const double pr = (C.prio == prb) ? 0.5 : fast_div(C.prio, C.prio + prb);
                        const double v[ROW] = {pr * (2.0 * d.px + (d.vx + mvx) * 1.0), pr * (2.0 * d.py + (d.vy + mvy) * 1.0),
                                               pr * (2.0 * d.pz + (d.vz + mvz) * 1.0), mx - d.px, my - d.py, mz - d.pz,
                                               (double)t_ah * 0.01, t_md, t_iet};
                        float *rr = wobs + OBS_SELF + ROW * r;
#pragma unroll
                        for (int q = 0; q < ROW; q++) { const float w = np_round2_f32(v[q]); bad |= (w * 0.f != 0.f); rr[q] = w; }
                    }
                    if (wnbr) wnbr[r] = t_j;
                });
                // rows beyond cnt: untouched zero fill after a step's clear, stale after a re-observe, or
                // never cleared in the stand-alone observe/reset/autoreset modes
                const int stale_end = (staged_in < 0) ? (is_step ? 0 : P.nm) : staged_in;
                for (int r = top.cnt; r < stale_end; r++) {
                    if (wobs) { float *rr = wobs + OBS_SELF + ROW * r;
#pragma unroll
                        for (int q = 0; q < ROW; q++) rr[q] = 0.0f; }
                    if (wnbr) wnbr[r] = -1;
                }
                if (ph != PH_A) staged = top.cnt;
                n_bad += bad ? 1u : 0u;
                if (L.count) L.count[wo] = (uint8_t)top.cnt;
            }
            RG_MARK(RG_OBSROW);
            if (ph == PH_A) {
                if (!own_a_skip) {
                // ---------------------------------------------------- ir_gym.rvo_reward_cal (ir_gym.py:64-133)
                const double vnum = 0.2 * fast_sqrt(np_dot3(ax, ay, az, ax, ay, az)), vden = fast_sqrt(np_dot3(d0, d1, d2, d0, d1, d2));
                const double vel_pen = (vden != 0.0) ? fast_div(vnum, vden) : vnum / vden;   // IEEE inf/nan when des = 0 (ir_gym.py:88)
                const double ap_ = angle_punish(d0, d1, d2, ax, ay, az);
                double safety = 0;
                if (vo_any) { double urg = 0; if (tmin < 2) urg = -8.0 * rare_exp(-tmin / 0.5); safety = -2.5 + urg; }
                rA = np_round3(ap_ + vel_pen + safety);
                }
                RG_MARK(RG_REWARD_A);
                if (SPLIT && helper) {      // hand the reward and the deviation of PASS A to the owner (SPLIT sync #2) and wait for the next step
                    s_rA[tid] = rA; s_dvA[tid] = dev;
                    asm volatile("barrier.sync 0;" ::: "memory");
                    break;
                }
                // ---------------------------------------------------- MOVE: drone.move_forward (drone.py:96-129,431-490)
                const double speed = fast_sqrt(np_dot3(d.vx, d.vy, d.vz, d.vx, d.vy, d.vz));
                const double acc = np_clip(ax * P.max_acc, -P.max_acc, P.max_acc);
                const double dyaw = np_clip(ay * P.max_ang, -P.max_ang, P.max_ang), dpitch = np_clip(az * P.max_ang, -P.max_ang, P.max_ang);
                double nv = speed + acc * 1.0; nv = (0.0 > nv) ? 0.0 : nv;
                d.yaw = py_mod360(d.yaw + dyaw); d.pitch = np_clip(d.pitch + dpitch, -90.0, 90.0);
                const double D2R = CUDART_PI / 180.0;
                double sy, cy, sp, cp; sincos_small(d.yaw * D2R, sy, cy); sincos_small(d.pitch * D2R, sp, cp);
                double nvx = nv * cp * cy, nvy = nv * cp * sy, nvz = nv * sp;
                if (d.flags & FLAG_DEST) { nvx = nvy = nvz = 0.0; }   // stop is always on (env_base.py:142)
                const double nx = d.px + nvx * P.dt, ny = d.py + nvy * P.dt, nz = d.pz + nvz * P.dt;
                { const double ux = nx - d.px, uy = ny - d.py, uz = nz - d.pz; d.real_len = d.real_len + fast_sqrt(np_dot3(ux, uy, uz, ux, uy, uz)); }
                d.px = nx; d.py = ny; d.pz = nz; d.vx = nvx; d.vy = nvy; d.vz = nvz;
                if (within_goal_v(P, d.px, d.py, d.pz, swp[0 * threads + tt], swp[1 * threads + tt], swp[2 * threads + tt])) {           // drone.py:116-129
                    const bool at_dest = within_goal(P, d.px, d.py, d.pz, C.wp(P) + 3 * (C.n - 1));
                    if (at_dest) P.extra_len[a] = (float)(d.real_len - C.route_len);      // (stored right away: one live double less)
                    if (!at_dest && d.i < C.n - 1) { d.i += 1; d.flags &= ~FLAG_ARRIVE; FILL_WP(tt); }
                }
                RG_MARK(RG_MOVE);
                continue;
            }

            if (ph == PH_OBS) { d.max_dev = rt(d.max_dev); continue; }

            // ------------------------------------------------------------ phase B only: flags, collisions, reward
            bool arr_now = false, dest_now = false;
            if (within_goal_v(P, d.px, d.py, d.pz, swp[0 * threads + tt], swp[1 * threads + tt], swp[2 * threads + tt]) && !(d.flags & FLAG_ARRIVE)) { d.flags |= FLAG_ARRIVE; arr_now = true; }
            if (d.flags & FLAG_ARRIVE) {
                const bool at_dest = within_goal(P, d.px, d.py, d.pz, C.wp(P) + 3 * (C.n - 1));
                if (at_dest) P.extra_len[a] = (float)(d.real_len - C.route_len);
                if (at_dest && !(d.flags & FLAG_DEST)) { d.flags |= FLAG_DEST; dest_now = true; }
            }
            RG_MARK(RG_B_FLAGS);
            const double *sb_ = sbld + (CLM ? 0 : le) * Mc * 4;
            for (int m = 0; m < nb; m++) {   // rvo_inter.py:99-105 filter, :198-209 check
                double bx, by, bh, br;
                if (Mc) { bx = sb_[4 * m]; by = sb_[4 * m + 1]; bh = sb_[4 * m + 2]; br = sb_[4 * m + 3]; }
                else { bx = __ldg(bld + 4 * m); by = __ldg(bld + 4 * m + 1); bh = __ldg(bld + 4 * m + 2); br = __ldg(bld + 4 * m + 3); }
                if (bh > d.pz - P.bld_zgate && d.pz <= bh) {
                    const double ex = d.px - bx, ey = d.py - by;
                    if (sqrt_le(fma(ey, ey, ex * ex), P.bld_range) && sqrt_le(ex * ex + ey * ey, C.radius + br)) col = true;
                }
            }
            {   // drone.drone_out_map (drone.py:213-225)
                const int le_ = CLM ? 0 : le; const double mapx = smap[0 * P.envs_per_cta + le_], mapy = smap[1 * P.envs_per_cta + le_], mapz = smap[2 * P.envs_per_cta + le_];
                if (d.px < 0 || d.px > mapx || d.py < 0 || d.py > mapy || d.pz < 0 || d.pz > mapz) col = true;
            }
            RG_MARK(RG_B_COL);
            double rB;   // ir_gym.mov_reward (ir_gym.py:256-311)
            if (col) rB = -50.0;
            else {
                double rew = 0;
                if (arr_now) rew += 3.0 * rare_pow(0.95, (double)((C.n - 1) - d.i));
                if (dest_now) rew += 20.0;
                const double dev_pen = -1.5 * fast_div(2.0, 1.0 + exp(-(dev * 10.0 - 5.0) / 0.3));
                double ex_pen = 0;
                const double exlen = d.real_len - C.route_len + 4.0;
                if (exlen > 0) { ex_pen = -0.3 * rare_log(exlen + 1.0 + 1e-6); if (ex_pen < -6.0 || ex_pen != ex_pen) ex_pen = -6.0; }
                rB = np_round3(rew + dev_pen + ex_pen);
            }
            RG_MARK(RG_B_REWARD);
            st_done = col; st_fin = (d.flags & FLAG_DEST) != 0;
            if constexpr (SPLIT) {          // SPLIT sync #2: the helper's PASS A results (max() commutes: the deviation of PASS A folds in here)
                asm volatile("barrier.sync 0;" ::: "memory");
                rA = s_rA[tid];
                const double dva = s_dvA[tid];
                if (dva > d.max_dev) d.max_dev = dva;
            }
            { const size_t o = OUT_SLOT(a);
              if (L.reward) L.reward[o] = (float)(rA + rB);
              if (L.done) L.done[o] = col;
              if (L.info) L.info[o] = (d.flags & FLAG_ARRIVE) != 0;
              if (L.finish) L.finish[o] = st_fin; }
            // what the next step will load from HBM
            d.px = rt(d.px); d.py = rt(d.py); d.pz = rt(d.pz); d.vx = rt(d.vx); d.vy = rt(d.vy); d.vz = rt(d.vz);
            d.yaw = rt(d.yaw); d.pitch = rt(d.pitch); d.real_len = rt(d.real_len); d.max_dev = rt(d.max_dev);
            RG_MARK(RG_B_OUT);
        }
        if (remapped) {   // hand max_deviation back to the owners and take my own drone up again
            if (worker) spk_d[3 * threads + tt] = d.max_dev;
            __syncthreads();
            if (valid) {
                d.px = sx[0 * threads + tid]; d.py = sx[1 * threads + tid]; d.pz = sx[2 * threads + tid];
                d.vx = sx[3 * threads + tid]; d.vy = sx[4 * threads + tid]; d.vz = sx[5 * threads + tid];
                d.yaw = spk_d[0 * threads + tid]; d.pitch = spk_d[1 * threads + tid]; d.real_len = spk_d[2 * threads + tid]; d.max_dev = spk_d[3 * threads + tid];
                d.i = spk_i[0 * threads + tid]; d.flags = spk_i[1 * threads + tid];
                const int sn = scn * N + k;
                C.set(P, sn);
                C.radius = sx[6 * threads + tid]; C.prio = sx[7 * threads + tid];
            }
        }
        RG_MARK(RG_TAIL_B);
        ENV_SYNC();   // the next step republishes sx and re-arms senv/ssel
        RG_MARK(RG_SYNC);
    }

    if (valid) {
        P.pos_yaw[a] = make_float4((float)d.px, (float)d.py, (float)d.pz, (float)d.yaw);
        P.vel_pitch[a] = make_float4((float)d.vx, (float)d.vy, (float)d.vz, (float)d.pitch);
        P.real_len[a] = (float)d.real_len; P.max_dev[a] = (float)d.max_dev;
        P.wpf[a] = d.i | d.flags;
        if (do_auto) P.ep_len[a] = ep_len;
    }
    HP_STAMP(4);       // (host-step profiling: CTA 0's step done)
    if (L.hp_self) {   // ---- host-pack tail: self part + existing rows of my agent -> mapped host memory, then the completion protocol
        // Writes over PCIe want full 128-byte lines, so every warp stages its 32 agents' data in a 2 KB slice of the (now idle)
        // exchange tile and streams it out with consecutive lanes on consecutive 16-byte words.
        __syncthreads();                                   // all warps are done with sx; re-observe compaction: rows written by other threads
        const int warp = tid >> 5;
        uint4 *stg = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(sx) + warp * 2048);      // sx has 64 B per thread
        float *stg_f = reinterpret_cast<float *>(stg);
        const int a_w = ag0 + (tid - lane);                 // first agent of this warp (a CTA's agents are contiguous)
        const int nv_w = helper ? 0 : max(0, min(32, nag_here - (tid - lane)));
        int kk = 0;
        const float *o = L.obs + (size_t)a * P.obs_dim;
        if (valid) {
            kk = __ldcg(L.count + a);
#pragma unroll
            for (int q = 0; q < OBS_SELF; q++) stg_f[lane * OBS_SELF + q] = __ldcg(o + q);
            L.hp_count[a] = (uint8_t)kk;
        }
        __syncwarp();
        {
            uint4 *hs = reinterpret_cast<uint4 *>(L.hp_self + (size_t)a_w * OBS_SELF);
            for (int i = lane; i < nv_w * (OBS_SELF / 4); i += 32) hs[i] = stg[i];
        }
        __syncwarp();
        int tot = kk, kmax = kk;                           // one reservation per warp
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, off); kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off)); }
        int base = 0;
        if (lane == 0 && tot) base = (int)(unsigned)atomicAdd(L.hp_ctr, (unsigned long long)tot);       // low word: rows reserved so far
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int r = 0; r < kmax; r++) {                   // round r: slot r of every agent that has one, packed and streamed together
            const unsigned have = __ballot_sync(0xffffffffu, kk > r);
            const int n = __popc(have);
            if (kk > r) {
                const int p = __popc(have & ((1u << lane) - 1u));
                float v[ROW];
#pragma unroll
                for (int q = 0; q < ROW; q++) v[q] = __ldcg(o + OBS_SELF + ROW * r + q);
                const int nb_ = L.nbr ? __ldcg(L.nbr + (size_t)a * P.nm + r) : -1;
                stg[3 * p + 0] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));   // UaisaRowRec
                stg[3 * p + 1] = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
                stg[3 * p + 2] = make_uint4(__float_as_uint(v[8]), (unsigned)a, (unsigned)r, (unsigned)nb_);
            }
            __syncwarp();
            const int room = max(0, min(n, L.hp_cap - base));
            uint4 *hr = reinterpret_cast<uint4 *>(static_cast<char *>(L.hp_rows) + (size_t)base * 48);
            for (int i = lane; i < 3 * room; i += 32) hr[i] = stg[i];
            __syncwarp();
            base += n;
        }
        // Completion (the pattern of a cooperative-groups grid barrier): the CTA barrier orders every thread's host writes -- these and the step's
        // reward / flags -- before thread 0's fence.  That fence only needs DEVICE scope: it makes the CTA's writes performed with respect to the
        // GPU before the CTA's arrival on the counter; the one CTA that observes all the others arrived then issues the SYSTEM fence, which is
        // cumulative over everything that happened before the arrivals it observed, and only then raises the ready word.  (A system fence per
        // CTA, as before, cost 4.6 us each on the critical path -- it waits for the PCIe writes -- i.e. two of them in series: 17.4 -> 13 us
        // from the first CTA's start to the publication.  UAISA_HOST_FENCE_SYS=1 restores it.)
        __syncthreads();
        HP_STAMP(5);
        if (tix == 0) {
            if (L.hp_sysfence) __threadfence_system(); else __threadfence();
            HP_STAMP(6);
            const unsigned long long old_ = atomicAdd(L.hp_ctr, 1ull << 32);      // high word: CTAs finished (their reservations precede their arrival)
            if ((unsigned)(old_ >> 32) == gridDim.x - 1) {
                __threadfence_system();
                *L.hp_ctr = 0ull;                          // re-armed for the next launch
                { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); L.hp_hdr[2] = g_; }                   // (profiling: the moment of publication)
                *reinterpret_cast<volatile unsigned long long *>(L.hp_hdr) = ((unsigned long long)(unsigned)old_ << 32) | 1ull;      // {ready, n_rows}
            }
        }
    }
    RG_MARK(RG_EPILOGUE);
    RG_FLUSH();
    const bool any_bad = __any_sync(0xffffffffu, n_bad != 0);
    if (P.count_pairs || any_bad) {
        for (int off = 16; off > 0; off >>= 1) {
            n_rng += __shfl_down_sync(0xffffffffu, n_rng, off); n_cone += __shfl_down_sync(0xffffffffu, n_cone, off);
            n_bad += __shfl_down_sync(0xffffffffu, n_bad, off);
        }
        if ((tid & 31) == 0) {
            if (P.count_pairs && n_rng) atomicAdd(P.counters + 0, (unsigned long long)n_rng);
            if (P.count_pairs && n_cone) atomicAdd(P.counters + 1, (unsigned long long)n_cone);
            if (n_bad) atomicAdd(P.counters + 2, (unsigned long long)n_bad);
        }
    }
}

#if !defined(UAISA_TU_DENSE) && !defined(UAISA_TU_SPLIT) && !defined(UAISA_TU_WIDE)   // (the small kernels live in the default translation unit only)
// cal_des_list (ir_gym.py:44-46)
__global__ void uaisa_cal_des_kernel(const __grid_constant__ Params P, float *des) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.E * P.N) return;
    const int env = a / P.N, k = a - env * P.N, s = __ldg(P.env2scn + env), sn = s * P.N + k;
    const float4 py = P.pos_yaw[a];
    const int i = P.wpf[a] & WP_MASK;
    double d0, d1, d2;
    cal_des_vel(P, py.x, py.y, py.z, P.wp + ((size_t)sn * P.W + i) * 3, d0, d1, d2);
    des[3 * a] = (float)d0; des[3 * a + 1] = (float)d1; des[3 * a + 2] = (float)d2;
}

// ---- K7: classic RVO velocity selection ------------------------------------------------------------------
// reciprocal_vel_obs.cal_vel -> config_vo -> vel_candidate -> vo_out2 -> vel_select -> penalty (uaisa_env/vel_obs/reciprocal_vel_obs.py:21-166).
// PARITY UNPINNED: that code is dead in the reference and raises on its first call; the behaviour implemented here is the corrected restatement
// oracle/uaisa_oracle.c:uo_rvo_select (fixes F1-F9 documented there and in DESIGN.md), against which the candidate choice is tested bit for bit.
// One thread per drone.  The neighbours of the env are walked ONCE: per neighbour the cone (apex PAA, axis, half-angle alpha) is built and every
// candidate velocity of the 0.5-spaced grid within +-acceler of the current velocity is classified against it (inside <=> alpha > beta) and its
// expected collision time folded into the candidate's minimum; the selection then runs over the candidates in the reference's loop order
// (vx outer, vz inner; first minimum wins): nearest to des_vel among the candidates outside every cone, else the one minimising
// 1/tc_min + |v - des_vel|.  ctrl (optional): the chosen velocity in the env's own controls [acc, dyaw/90, dpitch/90] (see the oracle's note).
constexpr int K7_MAXC = 5;                 // grid points per axis (2 * acceler / 0.5 + 1, acceler <= 1)
__device__ __forceinline__ double k7_exp_tim(double pax, double pay, double paz, double pbx, double pby, double pbz, double vax, double vay, double vaz,
                                             double vbx, double vby, double vbz, double ra, double rb) {       // vel_obs3D.cal_exp_tim, vel_obs3D.py:104-143
    const double rx = pax - pbx, ry = pay - pby, rz = paz - pbz, vx = vax - vbx, vy = vay - vby, vz = vaz - vbz, r = ra + rb;
    const double a = vx * vx + vy * vy + vz * vz, b = 2 * rx * vx + 2 * ry * vy + 2 * rz * vz, c = rx * rx + ry * ry + rz * rz - r * r;
    if (c <= 0) return 0.0;
    const double temp = b * b - 4 * a * c;
    if (temp <= 0) return CUDART_INF;
    const double sq = sqrt(temp), t1 = (-b + sq) / (2 * a), t2 = (-b - sq) / (2 * a);
    const double t3 = t1 >= 0 ? t1 : CUDART_INF, t4 = t2 >= 0 ? t2 : CUDART_INF;
    return t3 < t4 ? t3 : t4;
}
__global__ void __launch_bounds__(128) uaisa_rvo_select_kernel(const __grid_constant__ Params P, double vmax, double acceler, float *vel_out, float *ctrl_out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.E * P.N) return;
    const int N = P.N, env = a / N, k = a - env * N, scn = __ldg(P.env2scn + env), sn = scn * N + k;
    const float4 py = P.pos_yaw[a], vp = P.vel_pitch[a];
    const double pax = py.x, pay = py.y, paz = py.z, vax = vp.x, vay = vp.y, vaz = vp.z;
    const double ra = __ldg(P.radius + sn), pra = __ldg(P.prio + sn);
    double des0, des1, des2;
    cal_des_vel(P, pax, pay, paz, P.wp + ((size_t)sn * P.W + (P.wpf[a] & WP_MASK)) * 3, des0, des1, des2);        // agent_state[8:11]
    double grid[3][K7_MAXC]; int ng[3];
    const double va3[3] = {vax, vay, vaz};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {                                       // vel_candidate, reciprocal_vel_obs.py:85-96 (F9: end point included)
        const double lo = np_clip(va3[ax] - acceler, -vmax, vmax), hi = np_clip(va3[ax] + acceler, -vmax, vmax);
        int n = 0;
        while (n < K7_MAXC && lo + n * 0.5 <= hi + 1e-9) { grid[ax][n] = lo + n * 0.5; n++; }
        ng[ax] = n;
    }
    const int nc = ng[0] * ng[1] * ng[2];
    double tcmin[K7_MAXC * K7_MAXC * K7_MAXC];
    unsigned inside[4] = {0u, 0u, 0u, 0u}, slow[4] = {0u, 0u, 0u, 0u};
    for (int c = 0; c < nc; c++) {
        const int iz = c % ng[2], iy = (c / ng[2]) % ng[1], ix = c / (ng[2] * ng[1]);
        const double vx = grid[0][ix], vy = grid[1][iy], vz = grid[2][iz];
        tcmin[c] = CUDART_INF;
        if (sqrt(vx * vx + vy * vy + vz * vz) < 0.3) slow[c >> 5] |= 1u << (c & 31);                               // :98
    }
    for (int j = 0; j < N; j++) {
        if (j == k) continue;                                                                                        // F8
        const int o = env * N + j;
        const float4 qy = P.pos_yaw[o], qv = P.vel_pitch[o];
        const double pbx = qy.x, pby = qy.y, pbz = qy.z, vbx = qv.x, vby = qv.y, vbz = qv.z;
        const double fx = pax - pbx, fy = pay - pby, fz = paz - pbz;
        if (!(sqrt(np_dot3(fx, fy, fz, fx, fy, fz)) <= 10.0)) continue;                                              // preprocess :38-44
        const double rb = __ldg(P.radius + scn * N + j), prb = __ldg(P.prio + scn * N + j);
        const double rx = pbx - pax, ry = pby - pay, rz = pbz - paz;                                                 // get_rvo_array
        const double nab = sqrt(np_dot3(rx, ry, rz, rx, ry, rz)), R = ra + rb;
        const double q = R / nab;
        const double alpha_h = (q <= 1.0) ? py_round2_hundredths(asin(q)) : py_round2_hundredths(CUDART_PI / 2);     // get_alpha (F3), in hundredths
        const double pr = pra / (pra + prb);
        const double PAAx = pr * (2 * pax + (vax + vbx) * 1), PAAy = pr * (2 * pay + (vay + vby) * 1), PAAz = pr * (2 * paz + (vaz + vbz) * 1);
        for (int c = 0; c < nc; c++) {
            if (slow[c >> 5] & (1u << (c & 31))) continue;
            const int iz = c % ng[2], iy = (c / ng[2]) % ng[1], ix = c / (ng[2] * ng[1]);
            const double vx = grid[0][ix], vy = grid[1][iy], vz = grid[2][iz];
            const double Bx = (pax + vx * P.delta_t) - PAAx, By = (pay + vy * P.delta_t) - PAAy, Bz = (paz + vz * P.delta_t) - PAAz;   // vo_out2 :111-123
            const double AB = nab * sqrt(np_dot3(Bx, By, Bz, Bx, By, Bz));
            const double cosang = (AB != 0.0) ? np_dot3(rx, ry, rz, Bx, By, Bz) / AB : 0.0;                          // get_beta: unclamped arccos
            const double beta_h = rint(acos(cosang) * 100.0);
            if (alpha_h > beta_h) inside[c >> 5] |= 1u << (c & 31);                                                  // :125-126 (NaN beta -> outside)
            const double tc = k7_exp_tim(pax, pay, paz, pbx, pby, pbz, vx, vy, vz, vbx, vby, vbz, ra, rb);             // penalty :151-153 (F6)
            if (tc < tcmin[c]) tcmin[c] = tc;
        }
    }
    bool have_out = false, have_in = false;
    double best_out = 0, best_in = 0, ox = 0, oy = 0, oz = 0, ix_ = 0, iy_ = 0, iz_ = 0;
    for (int c = 0; c < nc; c++) {                                                                                   // vel_select :131-138
        if (slow[c >> 5] & (1u << (c & 31))) continue;
        const int iz = c % ng[2], iy = (c / ng[2]) % ng[1], ix = c / (ng[2] * ng[1]);
        const double vx = grid[0][ix], vy = grid[1][iy], vz = grid[2][iz];
        const double dx = des0 - vx, dy = des1 - vy, dz = des2 - vz;
        if (!(inside[c >> 5] & (1u << (c & 31)))) {
            const double dd = sqrt(dx * dx + dy * dy + dz * dz);                                                     // distance(v, vel_des)
            if (!have_out || dd < best_out) { have_out = true; best_out = dd; ox = vx; oy = vy; oz = vz; }
        } else {
            const double ex = vx - des0, ey = vy - des1, ez = vz - des2;                                             // distance(vel_des, v)
            const double tc_inv = (tcmin[c] == 0.0) ? CUDART_INF : 1.0 / tcmin[c];
            const double pen = 1.0 * tc_inv + sqrt(ex * ex + ey * ey + ez * ez);
            if (!have_in || pen < best_in) { have_in = true; best_in = pen; ix_ = vx; iy_ = vy; iz_ = vz; }
        }
    }
    double sx_ = vax, sy_ = vay, sz_ = vaz;                                                                          // F7
    if (have_out) { sx_ = ox; sy_ = oy; sz_ = oz; } else if (have_in) { sx_ = ix_; sy_ = iy_; sz_ = iz_; }
    vel_out[3 * a] = (float)sx_; vel_out[3 * a + 1] = (float)sy_; vel_out[3 * a + 2] = (float)sz_;
    if (ctrl_out) {
        const double sp_new = sqrt(sx_ * sx_ + sy_ * sy_ + sz_ * sz_), sp = sqrt(vax * vax + vay * vay + vaz * vaz);
        double dyaw = 0, dpitch = 0;
        if (sp_new > 0) {
            const double yaw_t = atan2(sy_, sx_) * (180.0 / CUDART_PI), pitch_t = atan2(sz_, sqrt(sx_ * sx_ + sy_ * sy_)) * (180.0 / CUDART_PI);
            double m = fmod(yaw_t - (double)py.w + 180.0, 360.0);
            if (m != 0.0) { if (m < 0.0) m += 360.0; } else m = 0.0;                                                 // Python float %
            dyaw = m - 180.0; dpitch = pitch_t - (double)vp.w;
        }
        ctrl_out[3 * a] = (float)np_clip(sp_new - sp, -1.0, 1.0); ctrl_out[3 * a + 1] = (float)np_clip(dyaw / 90.0, -1.0, 1.0);
        ctrl_out[3 * a + 2] = (float)np_clip(dpitch / 90.0, -1.0, 1.0);
    }
}

// ---- rollout-side kernels (SURVEY.md section 8a rows a-P2) ---------------------------------------------------
// Trainer action transform, multi_ppo.py:197-205: a_inc = np.round(a_inc, 2) on the float32 sample,
// abs_action = np.round(acceler * a_inc + cur_vel, 2) in float64 (cur_vel = drone.vel, read from the env state).
__global__ void uaisa_action_transform_kernel(const __grid_constant__ Params P, const float *a_inc, float acceler, float *a_inc_rounded, float *action) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.E * P.N) return;
    const float4 vp = P.vel_pitch[a];
    const float v[3] = {vp.x, vp.y, vp.z};
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float ar = __fdiv_rn(rintf(__fmul_rn(a_inc[3 * a + q], 100.0f)), 100.0f);   // np.round on a float32 array
        if (a_inc_rounded) a_inc_rounded[3 * a + q] = ar;
        const double ab = (double)__fmul_rn(acceler, ar) + (double)v[q];
        action[3 * a + q] = np_round2_f32(ab);
    }
}

// GAE-lambda over [T][A] buffers, multi_PPObuf.finish_path with last_val = 0 (multi_ppo.py:68-77): a path ends after
// step t where cut[t][a] != 0 (and at t = T-1); one thread per agent scans its column backwards.
__global__ void uaisa_gae_kernel(const float *rew, const float *val, const uint8_t *cut, int T, int A, float gamma, float lam, float *adv, float *ret) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    float next_v = 0.f, next_adv = 0.f, next_ret = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        const size_t o = (size_t)t * A + a;
        if (t == T - 1 || cut[o]) { next_v = 0.f; next_adv = 0.f; next_ret = 0.f; }
        const float r = rew[o], v = val[o];
        const float delta = r + gamma * next_v - v;
        next_adv = delta + gamma * lam * next_adv;
        next_ret = r + gamma * next_ret;
        next_v = v;
        adv[o] = next_adv; ret[o] = next_ret;
    }
}

// Pack kernel of uaisa_step_host_sparse: self part of every observation + the existing neighbour rows as records.
__global__ void uaisa_pack_kernel(int A, int obs_dim, int nm, const float *obs, const uint8_t *count, const int *nbr, float *self_obs,
                                  void *rows_v, int *n_rows, int cap) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    const float *o = obs + (size_t)a * obs_dim;
#pragma unroll
    for (int q = 0; q < OBS_SELF; q++) self_obs[(size_t)a * OBS_SELF + q] = o[q];
    const int k = count[a];
    struct Rec { float v[ROW]; int agent, slot, nbr; };
    Rec *rows = reinterpret_cast<Rec *>(rows_v);
    for (int r = 0; r < k; r++) {
        const int idx = atomicAdd(n_rows, 1);
        if (idx < cap) {
            Rec rec;
#pragma unroll
            for (int q = 0; q < ROW; q++) rec.v[q] = o[OBS_SELF + ROW * r + q];
            rec.agent = a; rec.slot = r; rec.nbr = nbr[(size_t)a * nm + r];
            rows[idx] = rec;
        }
    }
}

#endif  // default translation unit

#define UAISA_INST(R, T, W, S) template __global__ void uaisa_env_kernel<R, T, W, S>(const __grid_constant__ Params, const __grid_constant__ Launch);
#if defined(UAISA_TU_SPLIT)
UAISA_INST(4, 128, true, true)          // SPLIT: owner + helper warps (small batches of few-neighbour envs)
#elif defined(UAISA_TU_WIDE)
UAISA_INST(10, 192, false, true) UAISA_INST(16, 192, false, true)      // neighbour slices in CTAs of up to 192 threads
#elif defined(UAISA_TU_DENSE)
#define UAISA_INST_D(R, T, W, S) template __global__ void uaisa_env_kernel_d<R, T, W, S>(const __grid_constant__ Params, const __grid_constant__ Launch);
UAISA_INST_D(10, 128, true, false) UAISA_INST_D(16, 128, true, false)
UAISA_INST_D(10, 128, false, false) UAISA_INST_D(16, 128, false, false)
#undef UAISA_INST_D
#else
// (The default translation unit keeps ALL instantiations, also the ones whose 128-register twins in the other unit are the ones launched.)
UAISA_INST(4, 128, true, false) UAISA_INST(10, 128, true, false) UAISA_INST(16, 128, true, false)
UAISA_INST(4, 128, false, false) UAISA_INST(10, 128, false, false) UAISA_INST(16, 128, false, false)
UAISA_INST(10, 128, false, true) UAISA_INST(16, 128, false, true)      // with neighbour slices (helper threads)
#endif
#undef UAISA_INST

}  // namespace uaisa

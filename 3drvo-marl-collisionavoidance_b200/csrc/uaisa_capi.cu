// uaisa_capi.cu -- the C ABI of libuaisa_b200.so (declared in include/uaisa_b200.h).
// Host-side only: handle management, scenario upload, launch geometry, error reporting.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/uaisa_b200.h"
#include "uaisa_kernels.cuh"

using namespace uaisa;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return UAISA_ECUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

using KernelFn = void (*)(const Params, const Launch);

}  // namespace

struct UaisaHandle {
    UaisaConfig cfg{};
    Params P{};
    int device = 0;
    int A = 0;
    int grid = 0, threads = 0;
    size_t smem = 0;
    KernelFn kernel = nullptr;
    bool scenarios_set = false;
    int64_t launches = 0;
    std::vector<void *> allocs;
    // lazily allocated device I/O buffers for uaisa_step_host
    float *h_action = nullptr, *h_obs = nullptr, *h_reward = nullptr;
    uint8_t *h_count = nullptr, *h_flags = nullptr;   // flags: done | info | finish, A each
    int32_t *h_nbr = nullptr;
    cudaStream_t h_stream = nullptr;
    // uaisa_step_host_sparse: device block + pinned host mirror + one captured graph
    char *sp_dev = nullptr, *sp_host = nullptr, *sp_host_dev = nullptr; float *sp_hact = nullptr, *sp_hact_dev = nullptr; bool sp_zero_copy = false, sp_zc_action = false, sp_inflight = false;
    size_t sp_off_reward = 0, sp_off_self = 0, sp_off_flags = 0, sp_off_rows = 0, sp_bytes = 0, sp_phase1 = 0;
    int sp_cap = 0, sp_r0 = 0, sp_auto = -1, sp_maxlen = -1;
    cudaGraphExec_t sp_graph = nullptr;
};

namespace {

// every entry point that launches or copies selects the handle's device (and restores the caller's): a handle created on cuda:1 works
// whatever device is current in the calling thread
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) ok = false; else if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false; if (prev == dev) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
int dalloc(UaisaHandle *h, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    h->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return UAISA_OK;
}

// The polled host step returns when its kernel has PUBLISHED the results, a few microseconds before the kernel retires (state is
// already written; the pair / non-finite counters are not).  Every other entry point that touches the handle's device memory on
// another stream first lets that kernel finish.
int settle(UaisaHandle *h, void *stream) {
    if (h->sp_inflight && static_cast<cudaStream_t>(stream) != h->h_stream) { CK(cudaStreamSynchronize(h->h_stream)); h->sp_inflight = false; }
    return UAISA_OK;
}

int launch(UaisaHandle *h, const Launch &L, void *stream) {
    if (!h->scenarios_set) return fail(UAISA_ESTATE, "uaisa_set_scenarios has not been called");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    if (h->P.cl > 1) {     // thread-block cluster launch: the G CTAs of an env are co-scheduled and share the hardware cluster barrier
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(h->grid); cfg.blockDim = dim3(h->threads * h->P.slices); cfg.dynamicSmemBytes = h->smem; cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = h->P.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, h->kernel, h->P, L));
    } else {
        h->kernel<<<h->grid, h->threads * h->P.slices, h->smem, static_cast<cudaStream_t>(stream)>>>(h->P, L);
    }
    h->launches++;
    CK(cudaGetLastError());
    return UAISA_OK;
}

}  // namespace

extern "C" {

const char *uaisa_last_error(void) { return g_err.c_str(); }
int uaisa_version(void) { return 100; }

int uaisa_default_config(UaisaConfig *c) {
    if (!c) return fail(UAISA_EINVAL, "cfg is NULL");
    std::memset(c, 0, sizeof(*c));
    c->n_envs = 1; c->n_drones = 1; c->max_waypoints = 2; c->max_buildings = 1; c->n_scenarios = 1; c->nm = 10;
    c->env_train = 1; c->threads_per_block = 0; c->count_pairs = 0;
    c->goal_threshold = 0.4; c->max_acc = 1.0; c->max_angle_change = 90.0; c->dt = 1.0; c->vel_max = 1.0;
    c->neighbor_range = 10.0; c->building_range = 5.0; c->building_zgate = 2.0; c->ctime_threshold = 2.0;
    c->delta_t = 1.0; c->exp_radius = 0.2;
    return UAISA_OK;
}

int uaisa_create(const UaisaConfig *cfg, UaisaHandle **out) {
    if (!cfg || !out) return fail(UAISA_EINVAL, "NULL argument");
    *out = nullptr;
    if (cfg->n_envs < 1 || cfg->n_drones < 1 || cfg->n_drones > UAISA_MAX_DRONES) return fail(UAISA_EINVAL, "n_envs >= 1 and 1 <= n_drones <= 1024 required");
    if (cfg->max_waypoints < 2 || cfg->max_waypoints > 0xFFFF) return fail(UAISA_EINVAL, "2 <= max_waypoints <= 65535 required");
    if (cfg->max_buildings < 1 || cfg->n_scenarios < 1) return fail(UAISA_EINVAL, "max_buildings >= 1 and n_scenarios >= 1 required");
    if (cfg->nm < 0 || cfg->nm > UAISA_MAX_NM) return fail(UAISA_EINVAL, "0 <= nm <= 16 required");
    if ((long long)cfg->n_envs * cfg->n_drones > 0x7fffffffLL / 128) return fail(UAISA_EINVAL, "too many agents");
    int ndev = 0, cur_dev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(UAISA_ECUDA, "no CUDA device");
    CK(cudaGetDevice(&cur_dev));              // (every CUDA query that can fail early runs before the handle exists: nothing to free on that path)
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cur_dev));
    if (prop.major != 10) return fail(UAISA_ECUDA, "libuaisa_b200 is built for sm_100a (B200) only");
    UaisaHandle *h = new (std::nothrow) UaisaHandle();
    if (!h) return fail(UAISA_ENOMEM, "out of host memory");
    h->cfg = *cfg;
    h->device = cur_dev;
    const int E = cfg->n_envs, N = cfg->n_drones;
    h->A = E * N;
    Params &P = h->P;
    P.E = E; P.N = N; P.W = cfg->max_waypoints; P.M = cfg->max_buildings; P.nm = cfg->nm;
    P.obs_dim = UAISA_OBS_SELF + UAISA_ROW * cfg->nm;
    P.env_train = cfg->env_train; P.count_pairs = cfg->count_pairs; P.compact = 0;
    P.goal_thr = cfg->goal_threshold; P.max_acc = cfg->max_acc; P.max_ang = cfg->max_angle_change; P.dt = cfg->dt;
    P.vel_max = cfg->vel_max; P.nbr_range = cfg->neighbor_range; P.bld_range = cfg->building_range;
    P.bld_zgate = cfg->building_zgate; P.ctime = cfg->ctime_threshold; P.delta_t = cfg->delta_t; P.exp_radius = cfg->exp_radius;
    P.goal2_lo = P.goal_thr * P.goal_thr * (1.0 - 1e-14); P.goal2_hi = P.goal_thr * P.goal_thr * (1.0 + 1e-14);
    P.nbr2_lo = P.nbr_range * P.nbr_range * (1.0 - 1e-14); P.nbr2_hi = P.nbr_range * P.nbr_range * (1.0 + 1e-14);
    // launch geometry: whole envs per CTA (or a cluster of CTAs per env); shrink the CTA until the grid covers the 148 SMs twice
    int threads = cfg->threads_per_block, slices = 1;
    const bool cluster = N > 128;                 // an env that does not fit one 128-thread CTA is spread over a thread-block cluster
    const bool wl = (N <= 32) && (32 % N == 0);   // every env inside one warp
    const int need = cfg->nm < N - 1 ? cfg->nm : N - 1;
    const int maxr = need <= 4 ? 4 : need <= 10 ? 10 : 16;
    const int minthreads = cluster ? 32 : ((N + 31) / 32) * 32;
    const int nch = (N + 31) / 32;                // 32-neighbour chunks of the pair loop
    P.tile0_bytes = ((32 * P.obs_dim * 4) + 15) & ~15; P.tile1_bytes = ((32 * (P.nm ? P.nm : 1) * 4) + 15) & ~15;
    if (maxr > 4 && P.tile0_bytes > 2048) P.tile0_bytes = 2048;     // many-neighbour kernels: shared memory decides the resident CTAs; the fill is then a few bulk stores instead of one
    const bool force_compact = std::getenv("UAISA_FORCE_COMPACT") != nullptr, no_compact = std::getenv("UAISA_NO_COMPACT") != nullptr;
    // CTA-level compaction of the re-observe phase pays only with several warps per CTA and more than a handful of
    // drones per env (measured: N=16 +7 %, N=4 +2 % at 128 threads, -10 % with single-warp CTAs)
    auto compact_for = [&](int th, int s) { return !cluster && s == 1 && th >= 128 && (N > 4 || force_compact) && !no_compact; };
#ifndef UAISA_REGS_DENSE
#define UAISA_REGS_DENSE 128
#endif
    auto regs_for = [&](int s) { return (maxr > 4 && s == 1) ? UAISA_REGS_DENSE : 168; };   // register caps of the instantiations (uaisa_kernels.cu, Makefile)
    auto grid_for = [&](int th) { return cluster ? (long long)E * ((N + th - 1) / th) : (long long)(E + (th / N) - 1) / (th / N); };
    auto smem_for = [&](int th, int s) {
        const size_t epc = cluster ? 1 : th / N, cols = cluster ? N : th;
        return sizeof(double) * 8 * cols + sizeof(int) * 2 * epc + 16 + P.tile0_bytes + P.tile1_bytes   // exchange tile, env bits, TMA fill tiles
               + (maxr > 4 ? (size_t)th * s * maxr * 24 : 0)                                       // + shared top-nm lists (one set per slice)
               + (compact_for(th, s) ? (size_t)th * (4 * 8 + 3 * 4) + sizeof(int) * epc : 0) + sizeof(int) * 2 + 32   // + re-observe compaction
               + cols * 16 + 16                                                                    // + fp32 position copies
               + (size_t)(s - 1) * th * 12 + (size_t)th * 4 + 16                                     // + helpers' partial results, owners' activity
               + (size_t)th * 80 + epc * 24 + 16                                                   // + per-drone route cache, map sizes
               + (size_t)(P.M <= 16 ? P.M : 0) * 32 * epc                                          // + cylinders of the CTA's envs (used when a scenario has any)
               + ((wl && maxr <= 4 && s == 2) ? (size_t)th * (64 + 80 + 16 + 8) + 16 : 0);          // + owner / helper split: pre-move tile, helpers' route cache, mailboxes
    };
    const bool auto_threads = threads <= 0;
    if (auto_threads) {
        threads = 128;
        while (threads / 2 >= minthreads && threads / 2 >= 32 && (!cluster || (N + threads / 2 - 1) / (threads / 2) <= 8) &&
               grid_for(threads) < 2 * prop.multiProcessorCount) threads /= 2;
    }
    // Cost model per (egos per CTA, slices): waves of the grid over the resident CTA slots x (32-neighbour chunks per slice + 1 chunk-equivalent
    // of per-ego work), stretched by the warps sharing an SM.
    auto est_for = [&](int th, int s) -> double {
        const int block = th * s;
        long long per_sm = 4 * (16384 / ((long long)regs_for(s) * 32)) / (block / 32), by_smem = (long long)prop.sharedMemPerMultiprocessor / (long long)(smem_for(th, s) + 1024);
        if (by_smem < per_sm) per_sm = by_smem;
        if (per_sm > 32) per_sm = 32;
        if (per_sm < 1) return 1e300;
        const long long cap = per_sm * prop.multiProcessorCount, g = grid_for(th), waves = (g + cap - 1) / cap;
        const double wpsm = (double)(g < cap ? g : cap) * (block / 32) / prop.multiProcessorCount;
        const int csz = cluster ? (N + th - 1) / th : 1;      // CTAs per env: every publish stores into, and waits for, all of them (64 x 3 measured 3 % ahead of 32 x 3 on configs[3])
        return (double)waves * ((nch + s - 1) / s + 1.0) * (1.0 + 0.04 * wpsm) * (1.0 + 0.01 * csz);
    };
    if (auto_threads && !cluster && grid_for(threads) >= 2 * prop.multiProcessorCount) {
        // Wave quantisation: a T-step launch keeps every CTA resident for all T steps, so a grid of 1.15 waves costs 2 (configs[2]: 2048 warps on
        // 1776 warp slots ran at 1.24e9 agent-steps/s, the 1776 warps of 3552 envs at 1.79e9, profiles/wave_quant.py).  Smaller CTAs pack more
        // warps next to the shared-memory limit; take them when the model predicts >= 3 % less time.
        double best = est_for(threads, 1); int bt = threads;
        for (int th = threads - 32; th >= 32; th -= 32) {
            if (th < minthreads || th / N < 1) continue;
            const double e_ = est_for(th, 1);
            if (e_ < 0.97 * best) { best = e_; bt = th; }
        }
        threads = bt;
    }
    if (auto_threads && !wl && maxr > 4 && nch > 1) {
        // Many neighbours per ego and few egos (dense airspace): add neighbour slices -- helper threads that walk a share of each
        // ego's neighbours and hand partial results to the owner through shared memory.  The unsliced geometry stays unless slicing is
        // predicted >= 10 % faster.
        double best = 0.9 * est_for(threads, 1); int bt = threads, bs = 1;
        for (int th = 96; th >= 32; th -= 32) {       // (egos per CTA; 96 x 2 = 192 threads is the largest sliced CTA)
            if (th < minthreads || (cluster && (N + th - 1) / th > 8)) continue;
            for (int s = 2; s * th <= 192 && s <= nch; s++) { const double e_ = est_for(th, s); if (e_ < best) { best = e_; bt = th; bs = s; } }
        }
        threads = bt; slices = bs;
    }
    // Owner / helper split of the few-neighbour warp-local kernels (slices = 2 there): a helper warp per owner warp runs PASS A while the owner moves and
    // runs PASS B.  It shortens the dependent chain of a step, which is all a small batch is bound by; a batch that fills the machine would only pay
    // for the extra warps.  Taken when owners + helpers fit the resident warp slots of that instantiation (compiled without a register cap: 8 per SM).
    const bool can_split = wl && maxr <= 4;
    if (can_split && auto_threads && cfg->neighbor_slices == 0 && !std::getenv("UAISA_NO_SPLIT") &&
        (long long)E * N <= 4LL * 32 * prop.multiProcessorCount) { slices = 2; if (threads > 64) threads = 64; }
    if (cfg->neighbor_slices > 0) {
        slices = cfg->neighbor_slices;
        if (can_split && slices == 2) { if (threads > 64) { if (!auto_threads) { delete h; return fail(UAISA_EINVAL, "the owner / helper split needs threads_per_block <= 64"); } threads = 64; } }
        else if (slices > 1 && (wl || maxr <= 4 || slices * threads > 192 || slices > nch)) { delete h; return fail(UAISA_EINVAL, "neighbor_slices > 1 needs n_drones > 32, nm > 4 and slices * threads_per_block <= 192 (or = 2 for the owner / helper split of warp-local envs with nm <= 4 rows)"); }
    }
    if (cluster) {
        if (threads > 128 || threads % 32 || (N + threads - 1) / threads > 8) { delete h; return fail(UAISA_EINVAL, "n_drones > 128 needs threads_per_block in {32,64,96,128} with at most 8 CTAs per env"); }
        P.cl = (N + threads - 1) / threads;
    } else if (threads < N || threads > 128 || threads % 32) { delete h; return fail(UAISA_EINVAL, "threads_per_block must be a multiple of 32 in [n_drones, 128]"); }
    P.dsmem = (cluster && !std::getenv("UAISA_CLUSTER_L2")) ? 1 : 0;   // exchange through distributed shared memory unless asked for the L2 path
    h->threads = threads; P.threads = threads; P.slices = slices; P.envs_per_cta = cluster ? 1 : threads / N;
    P.defer_obs = std::getenv("UAISA_NO_DEFER") ? 0 : 1;   // T-step launches: re-observation of step t < T-1 rides on PASS A of step t+1 (A/B switch)
    P.compact = compact_for(threads, slices) ? 1 : 0;
    h->grid = cluster ? E * P.cl : (E + P.envs_per_cta - 1) / P.envs_per_cta;
    h->smem = smem_for(threads, slices);
    if (std::getenv("UAISA_VERBOSE")) std::fprintf(stderr, "[uaisa] E=%d N=%d: %s, %d egos x %d slices per CTA, %d CTAs%s, %zu B smem\n", E, N, wl ? "warp-local" : "CTA", threads, slices, h->grid, cluster ? " (clustered)" : "", h->smem);
    if (wl && slices == 2) h->kernel = uaisa_env_kernel<4, 128, true, true>;
    else if (wl) h->kernel = need <= 4 ? uaisa_env_kernel<4, 128, true, false> : need <= 10 ? uaisa_env_kernel_d<10, 128, true, false> : uaisa_env_kernel_d<16, 128, true, false>;
    else if (slices * threads > 128) h->kernel = need <= 10 ? uaisa_env_kernel<10, 192, false, true> : uaisa_env_kernel<16, 192, false, true>;
    else if (slices > 1 || (maxr > 4 && std::getenv("UAISA_FORCE_SL"))) h->kernel = need <= 10 ? uaisa_env_kernel<10, 128, false, true> : uaisa_env_kernel<16, 128, false, true>;
    else h->kernel = need <= 4 ? uaisa_env_kernel<4, 128, false, false> : need <= 10 ? uaisa_env_kernel_d<10, 128, false, false> : uaisa_env_kernel_d<16, 128, false, false>;
    if (h->smem > (size_t)prop.sharedMemPerBlockOptin) { delete h; return fail(UAISA_EINVAL, "shared memory per CTA exceeds the device limit; lower threads_per_block"); }
    cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void *>(h->kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaFuncSetAttribute"); }
    const size_t A = h->A, S = cfg->n_scenarios, SN = S * N;
    int rc = UAISA_OK;
    double *wp, *sg, *rl, *rad, *pr, *bld, *map; int *npts, *nbld, *e2s;
    if ((rc = dalloc(h, &wp, SN * P.W * 3)) || (rc = dalloc(h, &sg, SN * P.W * 4)) || (rc = dalloc(h, &npts, SN)) || (rc = dalloc(h, &rl, SN)) || (rc = dalloc(h, &rad, SN)) ||
        (rc = dalloc(h, &pr, SN)) || (rc = dalloc(h, &bld, S * P.M * 4)) || (rc = dalloc(h, &nbld, S)) || (rc = dalloc(h, &map, S * 3)) ||
        (rc = dalloc(h, &e2s, (size_t)E)) || (rc = dalloc(h, &P.pos_yaw, A)) || (rc = dalloc(h, &P.vel_pitch, A)) ||
        (rc = dalloc(h, &P.real_len, A)) || (rc = dalloc(h, &P.max_dev, A)) || (rc = dalloc(h, &P.extra_len, A)) ||
        (rc = dalloc(h, &P.wpf, A)) || (rc = dalloc(h, &P.ep_len, A)) || (rc = dalloc(h, &P.counters, 32)) ||
        (cluster && ((rc = dalloc(h, &P.gx, 2 * A * 6)) || (rc = dalloc(h, &P.genv, 2 * (size_t)E))))) {
        uaisa_destroy(h);
        return rc;
    }
    P.wp = wp; P.seg = sg; P.npts = npts; P.route_len = rl; P.radius = rad; P.prio = pr; P.bld = bld; P.nbld = nbld; P.map = map; P.env2scn = e2s;
    if (cluster) cudaMemset(P.genv, 0, sizeof(int) * 2 * E);
    cudaMemset(P.extra_len, 0, sizeof(float) * A); cudaMemset(P.ep_len, 0, sizeof(int) * A); cudaMemset(P.counters, 0, 32 * 8);
    *out = h;
    return UAISA_OK;
}

int uaisa_destroy(UaisaHandle *h) {
    if (!h) return UAISA_OK;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (void *p : h->allocs) cudaFree(p);
    if (h->sp_graph) cudaGraphExecDestroy(h->sp_graph);
    if (h->sp_host) cudaFreeHost(h->sp_host);
    if (h->sp_hact) cudaFreeHost(h->sp_hact);
    if (h->h_stream) cudaStreamDestroy(h->h_stream);
    delete h;
    return UAISA_OK;
}

int uaisa_set_scenarios(UaisaHandle *h, const double *waypoints, const int32_t *n_points, const double *radius, const double *priority,
                        const double *buildings, const int32_t *n_buildings, const double *map_size, const int32_t *env2scenario) {
    if (!h || !waypoints || !n_points || !radius || !priority || !buildings || !n_buildings || !map_size || !env2scenario)
        return fail(UAISA_EINVAL, "NULL argument");
    const Params &P = h->P;
    const size_t S = h->cfg.n_scenarios, N = P.N, SN = S * N, W = P.W;
    std::vector<double> rl(SN), sg(SN * W * 4, 0.0);
    for (size_t q = 0; q < SN; q++) {
        if (n_points[q] < 2 || n_points[q] > (int)W) return fail(UAISA_EINVAL, "n_points must be in [2, max_waypoints] (drone.py:46-48 reads waypoints[1])");
        double tot = 0.0;   // drone.calculate_total_length (drone.py:409-429)
        const double *w = waypoints + q * W * 3;
        for (int i = 0; i + 1 < n_points[q]; i++) {
            double dx = w[3 * (i + 1)] - w[3 * i], dy = w[3 * (i + 1) + 1] - w[3 * i + 1], dz = w[3 * (i + 1) + 2] - w[3 * i + 2];
            tot += std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        rl[q] = tot;
        for (size_t i = 1; i < W; i++) {   // segment wp[i-1] -> wp[i] as drone.calculate_deviation normalises it (drone.py:374-388)
            double dx = w[3 * i] - w[3 * (i - 1)], dy = w[3 * i + 1] - w[3 * (i - 1) + 1], dz = w[3 * i + 2] - w[3 * (i - 1) + 2];
            double mag = std::sqrt(dx * dx + dy * dy + dz * dz);
            double *g = sg.data() + (q * W + i) * 4;
            g[3] = mag;
            if (mag != 0.0) { g[0] = dx / mag; g[1] = dy / mag; g[2] = dz / mag; }
        }
    }
    int any_bld = 0;
    for (size_t s = 0; s < S; s++) { if (n_buildings[s] < 0 || n_buildings[s] > P.M) return fail(UAISA_EINVAL, "n_buildings out of range"); any_bld |= n_buildings[s] > 0; }
    h->P.bld_cache = (any_bld && P.M <= 16) ? 1 : 0;
    for (int e = 0; e < P.E; e++) if (env2scenario[e] < 0 || env2scenario[e] >= (int)S) return fail(UAISA_EINVAL, "env2scenario out of range");
    DeviceGuard g(h->device);
    CK(cudaMemcpy(const_cast<double *>(P.wp), waypoints, sizeof(double) * SN * W * 3, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.seg), sg.data(), sizeof(double) * SN * W * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int *>(P.npts), n_points, sizeof(int) * SN, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.route_len), rl.data(), sizeof(double) * SN, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.radius), radius, sizeof(double) * SN, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.prio), priority, sizeof(double) * SN, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.bld), buildings, sizeof(double) * S * P.M * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int *>(P.nbld), n_buildings, sizeof(int) * S, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(P.map), map_size, sizeof(double) * S * 3, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int *>(P.env2scn), env2scenario, sizeof(int) * P.E, cudaMemcpyHostToDevice));
    // a defined (if meaningless) state for the first MODE_RESET load
    CK(cudaMemset(P.pos_yaw, 0, sizeof(float4) * h->A)); CK(cudaMemset(P.vel_pitch, 0, sizeof(float4) * h->A));
    CK(cudaMemset(P.real_len, 0, sizeof(float) * h->A)); CK(cudaMemset(P.max_dev, 0, sizeof(float) * h->A));
    { std::vector<int> one((size_t)h->A, 1); CK(cudaMemcpy(P.wpf, one.data(), sizeof(int) * h->A, cudaMemcpyHostToDevice)); }
    h->scenarios_set = true;
    int rc = uaisa_reset(h, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    return UAISA_OK;
}

int uaisa_step(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward, uint8_t *done,
               uint8_t *info, uint8_t *finish, void *stream) {
    return uaisa_rollout(h, 1, action, obs, count, nbr, reward, done, info, finish, 0, 0, stream);
}

int uaisa_rollout(UaisaHandle *h, int32_t T, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward,
                  uint8_t *done, uint8_t *info, uint8_t *finish, int32_t auto_reset, int32_t max_ep_len, void *stream) {
    if (!h || !action) return fail(UAISA_EINVAL, "NULL handle or action");
    if (T < 1) return fail(UAISA_EINVAL, "T must be >= 1");
    Launch L{};
    L.mode = MODE_STEP; L.T = T; L.auto_reset = auto_reset ? 1 : 0; L.max_ep_len = max_ep_len; L.epoch_ended = 0;
    L.action = action; L.obs = obs; L.count = count; L.nbr = nbr; L.reward = reward; L.done = done; L.info = info; L.finish = finish;
    return launch(h, L, stream);
}

// One env step + the trainer reset policy + the re-observation of the touched envs in ONE launch (what uaisa_step followed by uaisa_autoreset
// computes, bit for bit): the step of multi_ppo.training_loop (multi_ppo.py:210-281) as the PPO rollout issues it.
int uaisa_step_autoreset(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward, uint8_t *done, uint8_t *info,
                         uint8_t *finish, int32_t max_ep_len, int32_t epoch_ended, uint8_t *reset_mask, uint8_t *path_cut, void *stream) {
    if (!h || !action) return fail(UAISA_EINVAL, "NULL handle or action");
    Launch L{};
    L.mode = MODE_STEP; L.T = 1; L.auto_reset = 1; L.max_ep_len = max_ep_len; L.epoch_ended = epoch_ended ? 1 : 0;
    L.action = action; L.obs = obs; L.count = count; L.nbr = nbr; L.reward = reward; L.done = done; L.info = info; L.finish = finish;
    L.reset_mask = reset_mask; L.path_cut = path_cut;
    return launch(h, L, stream);
}

int uaisa_reset(UaisaHandle *h, const uint8_t *agent_mask, float *obs, uint8_t *count, int32_t *nbr, void *stream) {
    if (!h) return fail(UAISA_EINVAL, "NULL handle");
    Launch L{};
    L.mode = MODE_RESET; L.T = 1; L.mask = agent_mask; L.obs = obs; L.count = count; L.nbr = nbr;
    return launch(h, L, stream);
}

int uaisa_observe(UaisaHandle *h, const uint8_t *env_mask, float *obs, uint8_t *count, int32_t *nbr, void *stream) {
    if (!h) return fail(UAISA_EINVAL, "NULL handle");
    Launch L{};
    L.mode = MODE_OBSERVE; L.T = 1; L.mask = env_mask; L.obs = obs; L.count = count; L.nbr = nbr;
    return launch(h, L, stream);
}

int uaisa_autoreset(UaisaHandle *h, const uint8_t *done, const uint8_t *finish, int32_t max_ep_len, int32_t epoch_ended, float *obs,
                    uint8_t *count, int32_t *nbr, uint8_t *reset_mask, uint8_t *path_cut, void *stream) {
    if (!h || !done || !finish) return fail(UAISA_EINVAL, "NULL handle / done / finish");
    Launch L{};
    L.mode = MODE_AUTORESET; L.T = 1; L.max_ep_len = max_ep_len; L.epoch_ended = epoch_ended ? 1 : 0;
    L.done_in = done; L.finish_in = finish; L.obs = obs; L.count = count; L.nbr = nbr; L.reset_mask = reset_mask; L.path_cut = path_cut;
    return launch(h, L, stream);
}

static int copy_state(UaisaHandle *h, const UaisaStateView *v, bool to_handle, void *stream) {
    if (!h || !v) return fail(UAISA_EINVAL, "NULL argument");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    const Params &P = h->P;
    const size_t A = h->A;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CP(field, hp, bytes) do { if (v->field) { if (to_handle) CK(cudaMemcpyAsync(hp, v->field, bytes, cudaMemcpyDefault, s)); else CK(cudaMemcpyAsync(v->field, hp, bytes, cudaMemcpyDefault, s)); } } while (0)
    CP(pos_yaw, P.pos_yaw, sizeof(float4) * A); CP(vel_pitch, P.vel_pitch, sizeof(float4) * A);
    CP(real_len, P.real_len, sizeof(float) * A); CP(max_dev, P.max_dev, sizeof(float) * A);
    CP(extra_len, P.extra_len, sizeof(float) * A); CP(wp_flags, P.wpf, sizeof(int) * A);
#undef CP
    return UAISA_OK;
}
int uaisa_get_state(UaisaHandle *h, const UaisaStateView *dst, void *stream) { return copy_state(h, dst, false, stream); }
int uaisa_set_state(UaisaHandle *h, const UaisaStateView *src, void *stream) { return copy_state(h, src, true, stream); }

int uaisa_cal_des(UaisaHandle *h, float *des, void *stream) {
    if (!h || !des) return fail(UAISA_EINVAL, "NULL argument");
    if (!h->scenarios_set) return fail(UAISA_ESTATE, "uaisa_set_scenarios has not been called");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    uaisa_cal_des_kernel<<<(h->A + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->P, des);
    h->launches++;
    CK(cudaGetLastError());
    return UAISA_OK;
}

int uaisa_rvo_select(UaisaHandle *h, double vmax, double acceler, float *vel_out, float *ctrl_out, void *stream) {
    if (!h || !vel_out) return fail(UAISA_EINVAL, "NULL argument");
    if (!h->scenarios_set) return fail(UAISA_ESTATE, "uaisa_set_scenarios has not been called");
    if (!(vmax > 0.0) || !(acceler > 0.0) || acceler > 1.0) return fail(UAISA_EINVAL, "vmax > 0 and 0 < acceler <= 1 required (at most 5 grid points per axis)");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    uaisa_rvo_select_kernel<<<(h->A + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->P, vmax, acceler, vel_out, ctrl_out);
    h->launches++;
    CK(cudaGetLastError());
    return UAISA_OK;
}

int uaisa_action_transform(UaisaHandle *h, const float *a_inc, float acceler, float *a_inc_rounded, float *action, void *stream) {
    if (!h || !a_inc || !action) return fail(UAISA_EINVAL, "NULL argument");
    if (!h->scenarios_set) return fail(UAISA_ESTATE, "uaisa_set_scenarios has not been called");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    uaisa_action_transform_kernel<<<(h->A + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->P, a_inc, acceler, a_inc_rounded, action);
    h->launches++;
    CK(cudaGetLastError());
    return UAISA_OK;
}

int uaisa_gae(const float *rew, const float *val, const uint8_t *cut, int32_t T, int32_t A, float gamma, float lam, float *adv, float *ret, void *stream) {
    if (!rew || !val || !cut || !adv || !ret || T < 1 || A < 1) return fail(UAISA_EINVAL, "bad argument");
    uaisa_gae_kernel<<<(A + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(rew, val, cut, T, A, gamma, lam, adv, ret);
    CK(cudaGetLastError());
    return UAISA_OK;
}

int uaisa_indicators(UaisaHandle *h, float *max_dev, float *extra_len, void *stream) {
    if (!h) return fail(UAISA_EINVAL, "NULL handle");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (max_dev) CK(cudaMemcpyAsync(max_dev, h->P.max_dev, sizeof(float) * h->A, cudaMemcpyDefault, s));
    if (extra_len) CK(cudaMemcpyAsync(extra_len, h->P.extra_len, sizeof(float) * h->A, cudaMemcpyDefault, s));
    return UAISA_OK;
}

int uaisa_get_counters(UaisaHandle *h, int64_t *in_range_pairs, int64_t *cone_evals, int64_t *nonfinite_obs, void *stream) {
    if (!h) return fail(UAISA_EINVAL, "NULL handle");
    DeviceGuard g(h->device);
    { int rc_ = settle(h, stream); if (rc_) return rc_; }
    unsigned long long c[4] = {0, 0, 0, 0};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(c, h->P.counters, sizeof(c), cudaMemcpyDeviceToHost, s));
    CK(cudaMemsetAsync(h->P.counters, 0, sizeof(c), s));
    CK(cudaStreamSynchronize(s));
    if (in_range_pairs) *in_range_pairs = (int64_t)c[0];
    if (cone_evals) *cone_evals = (int64_t)c[1];
    if (nonfinite_obs) *nonfinite_obs = (int64_t)c[2];
    return UAISA_OK;
}

int64_t uaisa_launch_count(const UaisaHandle *h) { return h ? h->launches : 0; }

int uaisa_get_region_clocks(UaisaHandle *h, int64_t *out16, void *stream) {
    if (!h || !out16) return fail(UAISA_EINVAL, "NULL argument");
    DeviceGuard g(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(out16, h->P.counters + 8, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemsetAsync(h->P.counters + 8, 0, 16 * sizeof(int64_t), s));
    CK(cudaStreamSynchronize(s));
    return UAISA_OK;
}

int uaisa_step_host(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward, uint8_t *done,
                    uint8_t *info, uint8_t *finish, int32_t auto_reset, int32_t max_ep_len) {
    if (!h || !action) return fail(UAISA_EINVAL, "NULL handle or action");
    DeviceGuard g(h->device);
    const size_t A = h->A; const Params &P = h->P;
    if (!h->h_action) {
        int rc;
        if ((rc = dalloc(h, &h->h_action, A * 3)) || (rc = dalloc(h, &h->h_obs, A * P.obs_dim)) || (rc = dalloc(h, &h->h_reward, A)) ||
            (rc = dalloc(h, &h->h_count, A)) || (rc = dalloc(h, &h->h_flags, A * 3)) || (rc = dalloc(h, &h->h_nbr, A * (P.nm ? P.nm : 1))))
            return rc;
        CK(cudaStreamCreateWithFlags(&h->h_stream, cudaStreamNonBlocking));
    }
    cudaStream_t s = h->h_stream;
    CK(cudaMemcpyAsync(h->h_action, action, sizeof(float) * A * 3, cudaMemcpyHostToDevice, s));
    int rc = uaisa_rollout(h, 1, h->h_action, obs ? h->h_obs : nullptr, count ? h->h_count : nullptr, nbr ? h->h_nbr : nullptr,
                           reward ? h->h_reward : nullptr, done ? h->h_flags : nullptr, info ? h->h_flags + A : nullptr,
                           finish ? h->h_flags + 2 * A : nullptr, auto_reset, max_ep_len, s);
    if (rc) return rc;
    if (obs) CK(cudaMemcpyAsync(obs, h->h_obs, sizeof(float) * A * P.obs_dim, cudaMemcpyDeviceToHost, s));
    if (count) CK(cudaMemcpyAsync(count, h->h_count, A, cudaMemcpyDeviceToHost, s));
    if (nbr) CK(cudaMemcpyAsync(nbr, h->h_nbr, sizeof(int32_t) * A * P.nm, cudaMemcpyDeviceToHost, s));
    if (reward) CK(cudaMemcpyAsync(reward, h->h_reward, sizeof(float) * A, cudaMemcpyDeviceToHost, s));
    if (done) CK(cudaMemcpyAsync(done, h->h_flags, A, cudaMemcpyDeviceToHost, s));
    if (info) CK(cudaMemcpyAsync(info, h->h_flags + A, A, cudaMemcpyDeviceToHost, s));
    if (finish) CK(cudaMemcpyAsync(finish, h->h_flags + 2 * A, A, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return UAISA_OK;
}

static int sparse_setup(UaisaHandle *h) {
    const size_t A = h->A; const Params &P = h->P;
    int rc;
    if (!h->h_action) {
        if ((rc = dalloc(h, &h->h_action, A * 3)) || (rc = dalloc(h, &h->h_obs, A * P.obs_dim)) || (rc = dalloc(h, &h->h_reward, A)) ||
            (rc = dalloc(h, &h->h_count, A)) || (rc = dalloc(h, &h->h_flags, A * 3)) || (rc = dalloc(h, &h->h_nbr, A * (P.nm ? P.nm : 1))))
            return rc;
        CK(cudaStreamCreateWithFlags(&h->h_stream, cudaStreamNonBlocking));
    }
    auto up16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t max_rows = A * (size_t)(P.nm ? P.nm : 1);
    h->sp_cap = (int)(max_rows * sizeof(UaisaRowRec) <= (size_t(256) << 20) ? max_rows : (size_t(256) << 20) / sizeof(UaisaRowRec));
    h->sp_r0 = (int)(A / 8 > 256 ? A / 8 : 256); if (h->sp_r0 > h->sp_cap) h->sp_r0 = h->sp_cap;
    h->sp_off_reward = 64;      // header: ready word, seven profiling timestamps (GPU %globaltimer: first CTA start, publication, CTA 0 after its prologue / step / result writes / fence)
    h->sp_off_self = up16(h->sp_off_reward + 4 * A);
    h->sp_off_flags = up16(h->sp_off_self + 4 * A * UAISA_OBS_SELF);      // done | info | finish | count
    h->sp_off_rows = up16(h->sp_off_flags + 4 * A);
    h->sp_phase1 = h->sp_off_rows + sizeof(UaisaRowRec) * (size_t)h->sp_r0;
    h->sp_bytes = h->sp_off_rows + sizeof(UaisaRowRec) * (size_t)h->sp_cap;
    if ((rc = dalloc(h, &h->sp_dev, h->sp_bytes))) return rc;
    CK(cudaHostAlloc(reinterpret_cast<void **>(&h->sp_host), h->sp_bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->sp_host_dev), h->sp_host, 0));
    CK(cudaMemset(h->sp_dev, 0, 16));
    CK(cudaHostAlloc(reinterpret_cast<void **>(&h->sp_hact), sizeof(float) * A * 3, cudaHostAllocMapped));   // read by the kernel straight over PCIe
    CK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->sp_hact_dev), h->sp_hact, 0));
    return UAISA_OK;
}

int uaisa_step_host_sparse(UaisaHandle *h, const float *action, int32_t auto_reset, int32_t max_ep_len, UaisaHostStep *out) {
    if (!h || !action || !out) return fail(UAISA_EINVAL, "NULL argument");
    if (!h->scenarios_set) return fail(UAISA_ESTATE, "uaisa_set_scenarios has not been called");
    DeviceGuard g(h->device);
    const size_t A = h->A; const Params &P = h->P;
    if (!h->sp_dev) { int rc = sparse_setup(h); if (rc) return rc; }
    cudaStream_t s = h->h_stream;
    if (!h->sp_graph || h->sp_auto != auto_reset || h->sp_maxlen != max_ep_len) {
        if (h->sp_graph) { cudaGraphExecDestroy(h->sp_graph); h->sp_graph = nullptr; }
        CK(cudaStreamSynchronize(s));
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        // The graph holds ONE kernel node: the env kernel writes reward / flags and (in its host-pack tail) the self parts and existing
        // rows straight into the mapped result block, and its last CTA raises the ready word the host polls -- no copy-engine D2H,
        // memset or pack nodes and no stream synchronisation on the critical path.  UAISA_HOST_COPY_ENGINE selects the older
        // H2D action / env kernel / pack kernel / D2H block graph (kept as a fallback and for A/B runs).
        const bool zero_copy = std::getenv("UAISA_HOST_COPY_ENGINE") == nullptr;
        // small action blocks are read over PCIe by the kernel itself; large ones are copied by the copy engine straight from the caller's
        // buffer (outside the graph: its address changes from call to call), without a staging memcpy on the host
        const bool zc_action = zero_copy && A * 3 * sizeof(float) <= (size_t(1) << 20);
        h->sp_zero_copy = zero_copy; h->sp_zc_action = zc_action;
        uint8_t *fl = reinterpret_cast<uint8_t *>(h->sp_dev + h->sp_off_flags);
        Launch L{};
        L.mode = MODE_STEP; L.T = 1; L.auto_reset = auto_reset ? 1 : 0; L.max_ep_len = max_ep_len;
        L.obs = h->h_obs; L.count = fl + 3 * A; L.nbr = h->h_nbr;
        if (zero_copy) {
            uint8_t *hfl = reinterpret_cast<uint8_t *>(h->sp_host_dev + h->sp_off_flags);
            L.action = zc_action ? h->sp_hact_dev : h->h_action;
            L.reward = reinterpret_cast<float *>(h->sp_host_dev + h->sp_off_reward); L.done = hfl; L.info = hfl + A; L.finish = hfl + 2 * A;
            L.hp_self = reinterpret_cast<float *>(h->sp_host_dev + h->sp_off_self); L.hp_count = hfl + 3 * A;
            L.hp_rows = h->sp_host_dev + h->sp_off_rows; L.hp_cap = h->sp_cap;
            L.hp_ctr = reinterpret_cast<unsigned long long *>(h->sp_dev); L.hp_hdr = reinterpret_cast<unsigned long long *>(h->sp_host_dev);
            L.hp_sysfence = std::getenv("UAISA_HOST_FENCE_SYS") ? 1 : 0;
        } else {
            cudaMemcpyAsync(h->h_action, h->sp_hact, sizeof(float) * A * 3, cudaMemcpyHostToDevice, s);
            cudaMemsetAsync(h->sp_dev, 0, 16, s);
            L.action = h->h_action;
            L.reward = reinterpret_cast<float *>(h->sp_dev + h->sp_off_reward); L.done = fl; L.info = fl + A; L.finish = fl + 2 * A;
        }
        { int rc_ = launch(h, L, s); if (rc_) { cudaGraph_t g_ = nullptr; cudaStreamEndCapture(s, &g_); if (g_) cudaGraphDestroy(g_); return rc_; } h->launches--; }
        if (!zero_copy) {
            uaisa_pack_kernel<<<(unsigned)((A + 127) / 128), 128, 0, s>>>((int)A, P.obs_dim, P.nm, h->h_obs, fl + 3 * A, h->h_nbr,
                                                                        reinterpret_cast<float *>(h->sp_dev + h->sp_off_self),
                                                                        h->sp_dev + h->sp_off_rows, reinterpret_cast<int *>(h->sp_dev), h->sp_cap);
            cudaMemcpyAsync(h->sp_host, h->sp_dev, h->sp_phase1, cudaMemcpyDeviceToHost, s);
        }
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
        ce = cudaGraphInstantiate(&h->sp_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
        h->sp_auto = auto_reset; h->sp_maxlen = max_ep_len;
    }
    static const bool prof = std::getenv("UAISA_HOST_PROF") != nullptr;
    static double acc[3] = {0, 0, 0}; static long nprof = 0;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = prof ? now() : 0;
    if (h->sp_zero_copy && !h->sp_zc_action) CK(cudaMemcpyAsync(h->h_action, action, sizeof(float) * A * 3, cudaMemcpyHostToDevice, s));
    else std::memcpy(h->sp_hact, action, sizeof(float) * A * 3);
    const double t1 = prof ? now() : 0;
    volatile unsigned long long *hdr = reinterpret_cast<volatile unsigned long long *>(h->sp_host);
    if (h->sp_zero_copy) *hdr = 0;
    CK(cudaGraphLaunch(h->sp_graph, s));
    const double t2 = prof ? now() : 0;
    int n_rows;
    if (h->sp_zero_copy) {
        h->launches += 1;
        unsigned long long w = 0;
        for (unsigned spin = 1;; spin++) {       // poll the ready word; look at the stream now and then so that a failed launch cannot hang us
            w = __atomic_load_n(reinterpret_cast<unsigned long long *>(h->sp_host), __ATOMIC_ACQUIRE);
            if (w & 1ull) break;
            if ((spin & 0x3fff) == 0) {
                cudaError_t q = cudaStreamQuery(s);
                if (q == cudaSuccess) { w = __atomic_load_n(reinterpret_cast<unsigned long long *>(h->sp_host), __ATOMIC_ACQUIRE); if (w & 1ull) break; return fail(UAISA_ECUDA, "host step finished without publishing its result"); }
                if (q != cudaErrorNotReady) return cuda_fail(q, "uaisa_step_host_sparse");
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        n_rows = (int)(w >> 32);
        h->sp_inflight = true;        // published, not necessarily retired (see settle())
    } else {
        h->launches += 2;
        CK(cudaStreamSynchronize(s));
        n_rows = *reinterpret_cast<int *>(h->sp_host);
    }
    if (prof) {
        const double t3 = now();
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2;
        static double kacc = 0;
        static double sacc[4] = {0, 0, 0, 0};
        if (h->sp_zero_copy) { const unsigned long long *hd = reinterpret_cast<const unsigned long long *>(h->sp_host); kacc += (double)(hd[2] - hd[1]) * 1e-3;
            for (int q = 0; q < 4; q++) sacc[q] += (double)((long long)hd[3 + q] - (long long)hd[1]) * 1e-3; }
        if (nprof % 200 == 199 && h->sp_zero_copy) { std::fprintf(stderr, "[uaisa host step] CTA 0 after first CTA start: prologue done %.1f us, step done %.1f us, results written %.1f us, fence done %.1f us\n", sacc[0] / 200, sacc[1] / 200, sacc[2] / 200, sacc[3] / 200); sacc[0] = sacc[1] = sacc[2] = sacc[3] = 0; }
        if (++nprof % 200 == 0) { std::fprintf(stderr, "[uaisa host step] memcpy %.1f us, graph launch %.1f us, wait %.1f us of which first CTA start -> publication %.1f us (avg of 200)\n", acc[0] / 200, acc[1] / 200, acc[2] / 200, kacc / 200); acc[0] = acc[1] = acc[2] = 0; kacc = 0; }
    }
    if (n_rows > h->sp_cap) return fail(UAISA_EINVAL, "neighbour-row capacity of the host block exceeded");
    if (!h->sp_zero_copy && n_rows > h->sp_r0) {   // rare: more rows than the first-phase window
        CK(cudaMemcpyAsync(h->sp_host + h->sp_phase1, h->sp_dev + h->sp_phase1, sizeof(UaisaRowRec) * (size_t)(n_rows - h->sp_r0), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    const uint8_t *fl = reinterpret_cast<const uint8_t *>(h->sp_host + h->sp_off_flags);
    out->reward = reinterpret_cast<const float *>(h->sp_host + h->sp_off_reward);
    out->self_obs = reinterpret_cast<const float *>(h->sp_host + h->sp_off_self);
    out->done = fl; out->info = fl + A; out->finish = fl + 2 * A; out->count = fl + 3 * A;
    out->rows = reinterpret_cast<const UaisaRowRec *>(h->sp_host + h->sp_off_rows);
    out->n_rows = n_rows; out->reserved = 0;
    return UAISA_OK;
}

}  // extern "C"

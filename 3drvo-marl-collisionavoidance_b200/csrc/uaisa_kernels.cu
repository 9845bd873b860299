// uaisa_kernels.cu -- the env kernel (all modes) for sm_100a.  See uaisa_kernels.cuh for the design.
#include "uaisa_kernels.cuh"

namespace uaisa {

// fp32 round trip: what a store to HBM followed by the next step's load would do
__device__ __forceinline__ double rt(double x) { return (double)(float)x; }

__device__ __forceinline__ void reset_drone(Drone &d, const Consts &C) {   // drone.reset, drone.py:270-291
    load3(C.wp, d.px, d.py, d.pz);
    d.px = rt(d.px); d.py = rt(d.py); d.pz = rt(d.pz);   // the state is fp32: a start point is stored rounded
    d.vx = d.vy = d.vz = 0.0; d.yaw = d.pitch = 0.0; d.real_len = 0.0; d.max_dev = 0.0;
    d.i = 1; d.flags = 0;   // extra_len is deliberately not cleared
}

__device__ __forceinline__ void publish(double *sx, int threads, int tid, const Drone &d, const Consts &C) {
    sx[0 * threads + tid] = d.px; sx[1 * threads + tid] = d.py; sx[2 * threads + tid] = d.pz;
    sx[3 * threads + tid] = d.vx; sx[4 * threads + tid] = d.vy; sx[5 * threads + tid] = d.vz;
    sx[6 * threads + tid] = C.radius; sx[7 * threads + tid] = C.prio;
}

// drone.dronestate (drone.py:254-263): desired velocity + deviation, side effect on max_deviation
__device__ __forceinline__ void dronestate(const Params &P, Drone &d, const Consts &C, double &d0, double &d1, double &d2, double &dev) {
    const double *cur = C.wp + 3 * d.i, *prev = C.wp + 3 * (d.i - 1);
    cal_des_vel(P, d.px, d.py, d.pz, cur, d0, d1, d2);
    dev = deviation(prev, cur, d.px, d.py, d.pz);
    if (dev > d.max_dev) d.max_dev = dev;
}

// Pairwise pass over the drones of my env.  PASS_B additionally keeps the top-nm list and the collision flag.
template <int MAXR, bool PASS_B>
__device__ __forceinline__ void pair_pass(const Params &P, const double *sx, int threads, int lbase, int N, int k,
                                          const Drone &d, const Consts &C, double ax, double ay, double az,
                                          bool &vo_any, double &t_min, bool &col, TopList<MAXR> &top, int cap,
                                          unsigned &n_rng, unsigned &n_cone) {
    vo_any = false; t_min = CUDART_INF;
    if (np_norm3(ax, ay, az) < 1e-5) { ax = ay = az = 0.0; }   // rvo_inter.py:118-119
    for (int j = 0; j < N; j++) {
        if (j == k) continue;
        const int o = lbase + j;
        double t, ah, md, iet;
        int bits = vo_pair(P, d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio,
                           sx[0 * threads + o], sx[1 * threads + o], sx[2 * threads + o], sx[3 * threads + o],
                           sx[4 * threads + o], sx[5 * threads + o], sx[6 * threads + o], sx[7 * threads + o],
                           ax, ay, az, t, ah, md, iet);
        n_rng += (bits & PB_EVAL) ? 1u : 0u; n_cone += (bits & PB_FULL) ? 1u : 0u;
        if (bits & PB_VO) {
            vo_any = true; if (t < t_min) t_min = t;
            if (PASS_B) top.insert(cap, iet, md, j, (int)ah);
        }
        if (PASS_B && (bits & PB_COL)) col = true;
    }
}

// Observation assembly into the staging row (ir_gym.py:203-229): 12 self values + rows, np.round(., 2).
// The staging area is zeroed (nbr: -1) cooperatively at the top of a step; `staged` = rows this thread has
// written since, so a re-observe after an auto-reset only clears what is stale.
template <int MAXR>
__device__ __forceinline__ bool stage_obs(const Params &P, float *row, int *nbr, const double *sx, int threads, int lbase,
                                          const Drone &d, const Consts &C, double d0, double d1, double d2, double dev,
                                          const TopList<MAXR> &top, int &staged) {
    const double self[OBS_SELF] = {d.px, d.py, d.pz, d.vx, d.vy, d.vz, C.radius, C.prio, d0, d1, d2, dev};
    bool bad = false;
#pragma unroll
    for (int q = 0; q < OBS_SELF; q++) row[q] = np_round2_f32(self[q], bad);
    // slot r (least urgent first, most urgent last) = top[cnt-1-r]
#pragma unroll
    for (int s = 0; s < MAXR; s++) {
        if (s < top.cnt) {
            const int r = top.cnt - 1 - s, o = lbase + top.j[s];
            const double mx = sx[0 * threads + o], my = sx[1 * threads + o], mz = sx[2 * threads + o];
            const double mvx = sx[3 * threads + o], mvy = sx[4 * threads + o], mvz = sx[5 * threads + o], prb = sx[7 * threads + o];
            const double pr = fast_div(C.prio, C.prio + prb);
            const double v[ROW] = {pr * (2.0 * d.px + (d.vx + mvx) * 1.0), pr * (2.0 * d.py + (d.vy + mvy) * 1.0),
                                   pr * (2.0 * d.pz + (d.vz + mvz) * 1.0), mx - d.px, my - d.py, mz - d.pz,
                                   (double)top.ah[s] * 0.01, top.md[s], top.iet[s]};
            float *rr = row + OBS_SELF + ROW * r;
#pragma unroll
            for (int q = 0; q < ROW; q++) rr[q] = np_round2_f32(v[q], bad);
            nbr[r] = top.j[s];
        }
    }
    for (int r = top.cnt; r < staged; r++) {
        float *rr = row + OBS_SELF + ROW * r;
#pragma unroll
        for (int q = 0; q < ROW; q++) rr[q] = 0.0f;
        nbr[r] = -1;
    }
    staged = top.cnt;
    return bad;
}

// n 32-bit words smem -> global, widest vectors the destination alignment allows (src is 16-byte aligned)
__device__ __forceinline__ void copy_out(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int n, int tid, int nthreads) {
    const uintptr_t da = reinterpret_cast<uintptr_t>(dst);
    if ((da & 15) == 0) {
        const int n4 = n >> 2;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src); uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (int i = tid; i < n4; i += nthreads) d4[i] = s4[i];
        for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
    } else if ((da & 7) == 0) {
        const int n2 = n >> 1;
        const uint2 *s2 = reinterpret_cast<const uint2 *>(src); uint2 *d2 = reinterpret_cast<uint2 *>(dst);
        for (int i = tid; i < n2; i += nthreads) d2[i] = s2[i];
        for (int i = (n2 << 1) + tid; i < n; i += nthreads) dst[i] = src[i];
    } else {
        for (int i = tid; i < n; i += nthreads) dst[i] = src[i];
    }
}

template <int MAXR, int MAXT>
__global__ void __launch_bounds__(MAXT) uaisa_env_kernel(const __grid_constant__ Params P, const __grid_constant__ Launch L) {
    extern __shared__ double smem_d[];
    const int threads = P.threads;
    double *sx = smem_d;                                                         // [8][threads] exchange
    const int obs_words = (threads * P.obs_dim + 3) & ~3, nbr_words = (threads * P.nm + 3) & ~3;
    float *sobs = reinterpret_cast<float *>(sx + 8 * threads);                   // [threads][obs_dim] dense rows
    int *snbr = reinterpret_cast<int *>(sobs + obs_words);                       // [threads][nm]
    int *senv = snbr + nbr_words;                                                // [envs_per_cta] reduction bits
    int *ssel = senv + P.envs_per_cta;                                           // [envs_per_cta] obs selected

    const int tid = threadIdx.x, N = P.N;
    const int le = tid / N, k = tid - le * N;
    const int env0 = blockIdx.x * P.envs_per_cta;
    const int env = env0 + le;
    const bool valid = (le < P.envs_per_cta) && (env < P.E);
    const int lbase = le * N;
    const int A = P.E * N;
    const int a = valid ? env * N + k : 0;
    const int nenv_here = min(P.envs_per_cta, P.E - env0);
    const int nag_here = nenv_here * N;          // agents of this CTA: contiguous range starting at env0*N
    const int cap = min(P.nm, MAXR);
    float *myrow = sobs + (size_t)tid * P.obs_dim;
    int *mynbr = snbr + (size_t)tid * P.nm;

    Consts C; Drone d; int ep_len = 0;
    const double *bld = nullptr; int nb = 0; double mapx = 0, mapy = 0, mapz = 0;
    double extra_len_new = 0.0; bool extra_len_dirty = false;
    if (valid) {
        const int s = __ldg(P.env2scn + env), sn = s * N + k;
        C.wp = P.wp + (size_t)sn * P.W * 3; C.radius = __ldg(P.radius + sn); C.prio = __ldg(P.prio + sn);
        C.route_len = __ldg(P.route_len + sn); C.n = __ldg(P.npts + sn);
        bld = P.bld + (size_t)s * P.M * 4; nb = __ldg(P.nbld + s);
        mapx = __ldg(P.map + 3 * s); mapy = __ldg(P.map + 3 * s + 1); mapz = __ldg(P.map + 3 * s + 2);
        const float4 py = P.pos_yaw[a], vp = P.vel_pitch[a];
        d.px = py.x; d.py = py.y; d.pz = py.z; d.yaw = py.w; d.vx = vp.x; d.vy = vp.y; d.vz = vp.z; d.pitch = vp.w;
        d.real_len = P.real_len[a]; d.max_dev = P.max_dev[a];
        const int wf = P.wpf[a]; d.i = wf & WP_MASK; d.flags = wf & (FLAG_ARRIVE | FLAG_DEST);
        if (L.auto_reset || L.mode == MODE_AUTORESET) ep_len = P.ep_len[a];
    }
    unsigned n_rng = 0, n_cone = 0, n_bad = 0;
    const int T = (L.mode == MODE_STEP) ? L.T : 1;

    for (int t = 0; t < T; t++) {
        bool st_done = false, st_fin = false;
        if (tid < P.envs_per_cta) { senv[tid] = 0; ssel[tid] = (L.mode == MODE_STEP) ? 1 : 0; }
        {   // clear the staging area: rows are mostly zero padding (nm slots), neighbour ids default to -1
            uint4 *z = reinterpret_cast<uint4 *>(sobs);
            for (int i = tid; i < (obs_words >> 2); i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
            uint4 *m = reinterpret_cast<uint4 *>(snbr);
            for (int i = tid; i < (nbr_words >> 2); i += blockDim.x) m[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        int staged = 0;
        __syncthreads();
        if (L.mode == MODE_STEP) {
            // ---------------------------------------------------------------- PASS A (pre-move)
            double ax = 0, ay = 0, az = 0, rA = 0;
            if (valid) {
                const float *ap = L.action + ((size_t)t * A + a) * 3;
                ax = ap[0]; ay = ap[1]; az = ap[2];
                publish(sx, threads, tid, d, C);
            }
            __syncthreads();
            if (valid) {
                double d0, d1, d2, dev; dronestate(P, d, C, d0, d1, d2, dev);
                bool vo, col = false; double tmin; TopList<MAXR> dummy; dummy.cnt = 0;
                pair_pass<MAXR, false>(P, sx, threads, lbase, N, k, d, C, ax, ay, az, vo, tmin, col, dummy, 0, n_rng, n_cone);
                // ir_gym.rvo_reward_cal (ir_gym.py:64-133)
                double e0 = np_round3(d0), e1 = np_round3(d1), e2 = np_round3(d2);
                double vel_pen = 0.2 * np_norm3(ax, ay, az) / np_norm3(e0, e1, e2);
                double ang = angle_between(e0, e1, e2, ax, ay, az), ap_;
                const double PI = CUDART_PI;
                if (-PI / 18 < ang && ang < PI / 18) ap_ = 3; else if (-PI / 6 < ang && ang < PI / 6) ap_ = 1;
                else if (-PI / 3 < ang && ang < PI / 3) ap_ = 0.5; else if (-PI / 2 < ang && ang < PI / 2) ap_ = 0; else ap_ = -4;
                double safety = 0;
                if (vo) { double urg = 0; if (tmin < 2) urg = -8.0 * exp(-tmin / 0.5); safety = -2.5 + urg; }
                rA = np_round3(ap_ + vel_pen + safety);
            }
            __syncthreads();
            // ---------------------------------------------------------------- MOVE
            if (valid) {
                double speed = np_norm3(d.vx, d.vy, d.vz);
                double acc = np_clip(ax * P.max_acc, -P.max_acc, P.max_acc);
                double dyaw = np_clip(ay * P.max_ang, -P.max_ang, P.max_ang), dpitch = np_clip(az * P.max_ang, -P.max_ang, P.max_ang);
                double nv = speed + acc * 1.0; nv = (0.0 > nv) ? 0.0 : nv;
                d.yaw = py_mod360(d.yaw + dyaw); d.pitch = np_clip(d.pitch + dpitch, -90.0, 90.0);
                const double D2R = CUDART_PI / 180.0;
                double sy, cy, sp, cp; sincos(d.yaw * D2R, &sy, &cy); sincos(d.pitch * D2R, &sp, &cp);
                double nvx = nv * cp * cy, nvy = nv * cp * sy, nvz = nv * sp;
                if (d.flags & FLAG_DEST) { nvx = nvy = nvz = 0.0; }   // stop is always on (env_base.py:142)
                double nx = d.px + nvx * P.dt, ny = d.py + nvy * P.dt, nz = d.pz + nvz * P.dt;
                d.real_len = d.real_len + np_norm3(nx - d.px, ny - d.py, nz - d.pz);
                d.px = nx; d.py = ny; d.pz = nz; d.vx = nvx; d.vy = nvy; d.vz = nvz;
                if (within_goal(P, d.px, d.py, d.pz, C.wp + 3 * d.i)) {           // drone.py:116-129
                    bool at_dest = within_goal(P, d.px, d.py, d.pz, C.wp + 3 * (C.n - 1));
                    if (at_dest) { extra_len_new = d.real_len - C.route_len; extra_len_dirty = true; }
                    if (!at_dest && d.i < C.n - 1) { d.i += 1; d.flags &= ~FLAG_ARRIVE; }
                }
                publish(sx, threads, tid, d, C);
            }
            __syncthreads();
            // ---------------------------------------------------------------- PASS B (post-move)
            if (valid) {
                double d0, d1, d2, dev; dronestate(P, d, C, d0, d1, d2, dev);
                bool arr_now = false, dest_now = false;
                if (within_goal(P, d.px, d.py, d.pz, C.wp + 3 * d.i) && !(d.flags & FLAG_ARRIVE)) { d.flags |= FLAG_ARRIVE; arr_now = true; }
                if (d.flags & FLAG_ARRIVE) {
                    bool at_dest = within_goal(P, d.px, d.py, d.pz, C.wp + 3 * (C.n - 1));
                    if (at_dest) { extra_len_new = d.real_len - C.route_len; extra_len_dirty = true; }
                    if (at_dest && !(d.flags & FLAG_DEST)) { d.flags |= FLAG_DEST; dest_now = true; }
                }
                double exlen = d.real_len - C.route_len + 4.0;
                bool col = false;
                for (int m = 0; m < nb; m++) {   // rvo_inter.py:99-105 filter, :198-209 check
                    double bx = __ldg(bld + 4 * m), by = __ldg(bld + 4 * m + 1), bh = __ldg(bld + 4 * m + 2), br = __ldg(bld + 4 * m + 3);
                    if (bh > d.pz - P.bld_zgate && np_norm2(d.px - bx, d.py - by) <= P.bld_range && d.pz <= bh) {
                        double ex = d.px - bx, ey = d.py - by;
                        if (sqrt(ex * ex + ey * ey) <= C.radius + br) col = true;
                    }
                }
                bool vo; double tmin; TopList<MAXR> top; top.clear();
                pair_pass<MAXR, true>(P, sx, threads, lbase, N, k, d, C, ax, ay, az, vo, tmin, col, top, cap, n_rng, n_cone);
                if (d.px < 0 || d.px > mapx || d.py < 0 || d.py > mapy || d.pz < 0 || d.pz > mapz) col = true;   // drone.py:213-225
                n_bad += stage_obs<MAXR>(P, myrow, mynbr, sx, threads, lbase, d, C, d0, d1, d2, dev, top, staged) ? 1u : 0u;
                // ir_gym.mov_reward (ir_gym.py:256-311)
                double rB;
                if (col) rB = -50.0;
                else {
                    double rew = 0;
                    if (arr_now) rew += 3.0 * pow(0.95, (double)((C.n - 1) - d.i));
                    if (dest_now) rew += 20.0;
                    double dev_pen = -1.5 * (2.0 / (1.0 + exp(-(dev * 10.0 - 5.0) / 0.3)));
                    double ex_pen = 0;
                    if (exlen > 0) { ex_pen = -0.3 * log(exlen + 1.0 + 1e-6); if (ex_pen < -6.0 || ex_pen != ex_pen) ex_pen = -6.0; }
                    rB = np_round3(rew + dev_pen + ex_pen);
                }
                st_done = col; st_fin = (d.flags & FLAG_DEST) != 0;
                const size_t o = (size_t)t * A + a;
                if (L.reward) L.reward[o] = (float)(rA + rB);
                if (L.done) L.done[o] = col;
                if (L.info) L.info[o] = (d.flags & FLAG_ARRIVE) != 0;
                if (L.finish) L.finish[o] = st_fin;
                if (L.count) L.count[o] = (uint8_t)top.cnt;
                // what the next step will load from HBM
                d.px = rt(d.px); d.py = rt(d.py); d.pz = rt(d.pz); d.vx = rt(d.vx); d.vy = rt(d.vy); d.vz = rt(d.vz);
                d.yaw = rt(d.yaw); d.pitch = rt(d.pitch); d.real_len = rt(d.real_len); d.max_dev = rt(d.max_dev);
            }
        } else if (L.mode == MODE_RESET) {
            if (valid && (L.mask == nullptr || L.mask[a])) { reset_drone(d, C); ssel[le] = 1; }   // benign race: all write 1
        } else if (L.mode == MODE_OBSERVE) {
            if (valid && k == 0) ssel[le] = (L.mask == nullptr || L.mask[env]) ? 1 : 0;
        } else {   // MODE_AUTORESET: flags come from the caller's buffers
            if (valid) { st_done = L.done_in[a] != 0; st_fin = L.finish_in[a] != 0; }
        }

        // -------------------------------------------------------------------- trainer reset policy
        const bool do_auto = (L.mode == MODE_AUTORESET) || (L.mode == MODE_STEP && L.auto_reset);
        if (do_auto) {
            __syncthreads();   // senv zeroed at the top of the iteration
            if (valid) {
                ep_len += 1;
                int bits = (st_done ? 1 : 0) | (st_fin ? 2 : 4) | (ep_len > L.max_ep_len ? 8 : 0);
                atomicOr(&senv[le], bits);
            }
            __syncthreads();
            if (valid) {
                const int bits = senv[le];
                const bool all_fin = !(bits & 4), terminal = (bits & 2) || (bits & 8);
                bool rs = false, cut = false, ob = false;
                if (bits & 1) { if (st_done) { rs = true; ep_len = 0; } ob = true; }                   // multi_ppo.py:230-242
                if (L.epoch_ended || all_fin) { rs = true; ep_len = 0; cut = true; ob = true; }        // :244-264
                else if (terminal) { if (st_fin || ep_len > L.max_ep_len) { rs = true; ep_len = 0; } cut = true; ob = true; }   // :266-281
                if (rs) reset_drone(d, C);
                if (ob) ssel[le] = (L.mode == MODE_STEP) ? 2 : 1;   // 2 = re-observe after a step
                const size_t o = (size_t)t * A + a;
                if (L.reset_mask) L.reset_mask[o] = rs;
                if (L.path_cut) L.path_cut[o] = cut;
            }
        }

        // -------------------------------------------------------------------- observe (act = 0)
        if (L.mode != MODE_STEP || do_auto) {
            __syncthreads();
            if (valid) publish(sx, threads, tid, d, C);
            __syncthreads();
            const int want = (L.mode == MODE_STEP) ? 2 : 1;
            if (valid && ssel[le] == want && (L.obs || L.count || L.nbr)) {
                double d0, d1, d2, dev; dronestate(P, d, C, d0, d1, d2, dev);
                bool vo, col = false; double tmin; TopList<MAXR> top; top.clear();
                pair_pass<MAXR, true>(P, sx, threads, lbase, N, k, d, C, 0.0, 0.0, 0.0, vo, tmin, col, top, cap, n_rng, n_cone);
                n_bad += stage_obs<MAXR>(P, myrow, mynbr, sx, threads, lbase, d, C, d0, d1, d2, dev, top, staged) ? 1u : 0u;
                if (L.count) L.count[(size_t)t * A + a] = (uint8_t)top.cnt;
                d.max_dev = rt(d.max_dev);
            }
        }

        // -------------------------------------------------------------------- coalesced flush of the staged rows
        __syncthreads();
        {
            const size_t abase = (size_t)t * A + (size_t)env0 * N;
            if (L.mode == MODE_STEP) {   // every env of the CTA is written: one flat, vectorised copy
                if (L.obs) copy_out(reinterpret_cast<uint32_t *>(L.obs + abase * P.obs_dim), reinterpret_cast<const uint32_t *>(sobs), nag_here * P.obs_dim, tid, blockDim.x);
                if (L.nbr) copy_out(reinterpret_cast<uint32_t *>(L.nbr + abase * P.nm), reinterpret_cast<const uint32_t *>(snbr), nag_here * P.nm, tid, blockDim.x);
            } else {                     // only the selected envs: one warp per row
                const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
                for (int la = warp; la < nag_here; la += nwarps) {
                    if (!ssel[la / N]) continue;
                    if (L.obs) for (int q = lane; q < P.obs_dim; q += 32) L.obs[(abase + la) * P.obs_dim + q] = sobs[(size_t)la * P.obs_dim + q];
                    if (L.nbr) for (int q = lane; q < P.nm; q += 32) L.nbr[(abase + la) * P.nm + q] = snbr[(size_t)la * P.nm + q];
                }
            }
        }
        __syncthreads();
    }

    if (valid) {
        P.pos_yaw[a] = make_float4((float)d.px, (float)d.py, (float)d.pz, (float)d.yaw);
        P.vel_pitch[a] = make_float4((float)d.vx, (float)d.vy, (float)d.vz, (float)d.pitch);
        P.real_len[a] = (float)d.real_len; P.max_dev[a] = (float)d.max_dev;
        P.wpf[a] = d.i | d.flags;
        if (extra_len_dirty) P.extra_len[a] = (float)extra_len_new;
        if (L.auto_reset || L.mode == MODE_AUTORESET) P.ep_len[a] = ep_len;
    }
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

template __global__ void uaisa_env_kernel<4, 256>(const __grid_constant__ Params, const __grid_constant__ Launch);
template __global__ void uaisa_env_kernel<10, 256>(const __grid_constant__ Params, const __grid_constant__ Launch);
template __global__ void uaisa_env_kernel<16, 256>(const __grid_constant__ Params, const __grid_constant__ Launch);
template __global__ void uaisa_env_kernel<16, 1024>(const __grid_constant__ Params, const __grid_constant__ Launch);

}  // namespace uaisa

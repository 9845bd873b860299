// uaisa_kernels.cuh -- device side of libuaisa_b200.so (sm_100a only).
//
// One kernel template runs every mode of the env hot path; one thread owns one drone, one CTA owns a
// block of whole environments, so all coupling (the pairwise RVO passes) stays inside the CTA:
//
//   PASS A  (pre-move states)   rvo reward            ir_gym.py:50-133, rvo_inter.py:63-83
//   MOVE                        kinematics + FSM      drone.py:96-129,431-490 (env_base.py:135-144)
//   PASS B  (post-move states)  collision / obs / reward   ir_gym.py:136-311, rvo_inter.py:20-61
//   [AUTO-RESET + OBSERVE]      trainer reset policy  multi_ppo.py:221-281, ir_gym.py:372-383
//
// Numerics: mutable state lives in HBM as fp32 SoA (float4 pos+yaw, float4 vel+pitch, ...).  Inside a
// step everything is evaluated in fp64 from those fp32 values -- every discrete output of the reference
// (flags, neighbour order, the 2-dp/3-dp rounded observations and rewards) hangs off comparisons and
// decimal roundings of fp64 quantities, and an fp32 evaluation flips them at 1e-5..1e-3 per value
// (SURVEY.md section 7).  The post-move state is exchanged between the threads of an env in fp64 through
// shared memory and only rounded to fp32 when it is written back, so one step started from the same
// fp32 state reproduces the fp64 reference to the last ulp of libm.  The file is compiled with
// -fmad=false: the only fused operations are the explicit fma() chains that mirror numpy's dot/norm.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace uaisa {

enum Mode : int { MODE_STEP = 0, MODE_OBSERVE = 1, MODE_RESET = 2, MODE_AUTORESET = 3 };

constexpr int OBS_SELF = 12;
constexpr int ROW = 9;
constexpr int FLAG_ARRIVE = 1 << 16;
constexpr int FLAG_DEST = 1 << 17;
constexpr int WP_MASK = 0xFFFF;

struct Params {
    // geometry
    int E, N, W, M, nm, obs_dim;
    int envs_per_cta, threads;
    int tile0_bytes, tile1_bytes;   // constant smem tiles behind the TMA fills (zeros / 0xFF), multiples of 16
    int env_train, count_pairs, compact;
    int defer_obs;                   // T-step launches with the reset policy: the re-observation of step t < T-1 is evaluated inside PASS A of step t+1
    int bld_cache;                   // cylinders of the CTA's envs are cached in shared memory (set by uaisa_set_scenarios: some scenario has cylinders, M <= 16)
    int slices;                      // neighbour slices per ego (helper threads; 1 = off).  blockDim.x = threads * slices
    int dsmem;                       // cluster mode: exchange the columns through distributed shared memory (1) or through L2 (0)
    int cl;                          // CTAs per env (thread-block cluster size) when an env does not fit one CTA; 0 = one CTA holds whole envs
    double *gx; int *genv;           // cluster mode: global exchange buffers [2][A][6] (p, v in fp64) and [2][E] (reset-policy bits)
    // hyper-parameters (reference constructor defaults, see include/uaisa_b200.h)
    double goal_thr, max_acc, max_ang, dt, vel_max, nbr_range, bld_range, bld_zgate, ctime, delta_t, exp_radius;
    double goal2_lo, goal2_hi, nbr2_lo, nbr2_hi;   // T^2*(1 -+ 1e-14) guard bands for sqrt(x) <= T (host-computed)
    // scenario constants (device, fp64, shared by envs -> L2/L1 resident)
    const double *wp; const double *seg; const int *npts; const double *route_len; const double *radius; const double *prio;
    const double *bld; const int *nbld; const double *map; const int *env2scn;
    // mutable state (device, fp32 SoA)
    float4 *pos_yaw; float4 *vel_pitch; float *real_len; float *max_dev; float *extra_len; int *wpf; int *ep_len;
    unsigned long long *counters;  // [0] in-range pairs, [1] cone evaluations, [2] non-finite observations
};

struct Launch {
    int mode, T, auto_reset, max_ep_len, epoch_ended;
    const float *action;         // [T][A][3]
    const uint8_t *mask;         // MODE_RESET: agent mask; MODE_OBSERVE: env mask; (NULL = all)
    const uint8_t *done_in, *finish_in;  // MODE_AUTORESET
    float *obs; uint8_t *count; int *nbr; float *reward; uint8_t *done, *info, *finish;
    uint8_t *reset_mask, *path_cut;
    // host-pack tail of uaisa_step_host_sparse (T == 1, NULL otherwise): the CTAs write the ragged result straight into mapped
    // pinned HOST memory and the last one to finish publishes {1, n_rows} in *hp_hdr, which the host polls.
    float *hp_self;              // HOST [A][12]
    uint8_t *hp_count;           // HOST [A]
    void *hp_rows;               // HOST [hp_cap] UaisaRowRec
    int hp_cap;
    unsigned long long *hp_ctr;  // DEVICE: low word rows reserved, high word CTAs finished (left at 0 by the last CTA)
    unsigned long long *hp_hdr;  // HOST: low word = ready flag, high word = n_rows
    int hp_sysfence;             // 1: every CTA issues a system-scope fence before its arrival (the older completion protocol, A/B switch)
};

// ---- numpy / python arithmetic primitives (see oracle/uaisa_oracle.c for the derivation) ----------
__device__ __forceinline__ double np_dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return fma(a2, b2, fma(a1, b1, a0 * b0));
}
__device__ __forceinline__ double np_norm3(double a0, double a1, double a2) { return sqrt(np_dot3(a0, a1, a2, a0, a1, a2)); }
__device__ __forceinline__ double np_norm2(double a0, double a1) { return sqrt(fma(a1, a1, a0 * a0)); }
__device__ __forceinline__ double np_clip(double x, double lo, double hi) { return (x != x) ? x : fmin(fmax(x, lo), hi); }   // np.clip propagates NaN

// ---- division without the libdevice slow path ---------------------------------------------------------
// a / c for a compile-time constant c with y = RN(1/c): q0 = a*y, r = a - c*q0 (exact in the FMA),
// q = q0 + r*y is the correctly rounded quotient (Markstein); verified exhaustively on the host against
// IEEE division for every integer |a| <= 2^26 and 2e8 random numerators for c = 100 and c = 1000
// (DESIGN.md "exact rounding").  a must be finite.
__device__ __forceinline__ double div_const(double a, double c, double y) {
    double q0 = a * y, r = fma(-c, q0, a);
    return fma(r, y, q0);
}
// General a / b for finite, normal, non-zero b: MUFU.RCP64H seed, two Newton steps, Markstein correction.
// Correctly rounded except in rare last-bit cases; used only where an ulp cannot matter (the value is later
// rounded to 2-3 decimals or compared with a threshold; a flip needs a 1e-16-wide coincidence).
__device__ __forceinline__ double fast_rcp(double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0); y = fma(y, e, y);
    e = fma(-b, y, 1.0); y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ double fast_div(double a, double b) {
    double y = fast_rcp(b), q = a * y, r = fma(-b, q, a);
    return fma(r, y, q);
}
// 1/sqrt(x) and sqrt(x) for normal x > 0 without libdevice's special-case path: MUFU.RSQ64H seed (2^-22),
// two Newton steps; sqrt adds a Markstein correction (correctly rounded except rare last-bit cases).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0); y = fma(0.5 * y, e, y);
    e = fma(-(x * y), y, 1.0); y = fma(0.5 * y, e, y);
    return y;
}
__device__ __forceinline__ double fast_sqrt(double x) {   // x >= 0 (0 -> 0)
    const double y = fast_rsqrt(x);
    double s = x * y;
    s = fma(fma(-s, s, x), 0.5 * y, s);
    return x > 0.0 ? s : 0.0;
}
// np.round(x, 2) / np.round(x, 3) = rint(x*10^k)/10^k, bit-exact for finite x
__device__ __forceinline__ double np_round2(double x) { double k = rint(x * 100.0); return isfinite(k) ? div_const(k, 100.0, 0.01) : k; }
__device__ __forceinline__ double np_round3(double x) { double k = rint(x * 1000.0); return isfinite(k) ? div_const(k, 1000.0, 0.001) : k; }
// (float)np.round(x, 2): the fp32 image of rint(x*100)/100 equals the fp32 image of rint(x*100)*0.01
// (k/100 is never within 2^-30 of an fp32 rounding tie unless it is exact; checked exhaustively, ibid.)
__device__ __forceinline__ float np_round2_f32(double x) { return (float)(rint(x * 100.0) * 0.01); }
// Rare libm paths.  Inlined, each is 60-170 instructions in the middle of the hot instruction stream that every warp has to branch over (the
// kernel's second largest stall reason is `no_instruction`); moving them out of line (-DUAISA_RARE_NOINLINE) was MEASURED 3-4 % slower on the
// sparse workloads (call ABI: more spills around the call sites) and 1 % faster on configs[2] -- so they stay inline.
#ifdef UAISA_RARE_NOINLINE
#define UAISA_RARE_ATTR __noinline__
#else
#define UAISA_RARE_ATTR __forceinline__
#endif
static __device__ UAISA_RARE_ATTR double rare_exp(double x) { return exp(x); }
static __device__ UAISA_RARE_ATTR double rare_log(double x) { return log(x); }
static __device__ UAISA_RARE_ATTR double rare_pow(double x, double y) { return pow(x, y); }
static __device__ UAISA_RARE_ATTR double rare_fmod(double x, double y) { return fmod(x, y); }
static __device__ UAISA_RARE_ATTR double rare_asin(double x) { return asin(x); }
static __device__ UAISA_RARE_ATTR double rare_acos(double x) { return acos(x); }
// sqrt(x) <= T without the square root unless x is within 1e-14 of T^2 (NaN -> false)
static __device__ __noinline__ bool sqrt_le_exact(double x, double T) { return sqrt(x) <= T; }   // rare: kept out of line (i-cache)
__device__ __forceinline__ bool sqrt_le_band(double x, double T, double lo, double hi) {
    const bool below = x < lo;
    if (below || x > hi) return below;
    return sqrt_le_exact(x, T);
}
__device__ __forceinline__ bool sqrt_le(double x, double T) {
    const double T2 = T * T;
    return sqrt_le_band(x, T, T2 * (1.0 - 1e-14), T2 * (1.0 + 1e-14));
}
// sin/cos of t = deg * (pi/180) for |t| <~ 8 (yaw in [0,360], pitch in [-90,90]): two-term Cody-Waite
// reduction by pi/2 and the fdlibm __kernel_sin/__kernel_cos minimax polynomials (< 1 ulp), no slow path.
__device__ __forceinline__ void sincos_small(double t, double &s, double &c) {
    const double k = rint(t * 6.36619772367581382433e-01);          // 2/pi
    double r = fma(-k, 1.57079632679489655800e+00, t);             // pi/2 hi
    r = fma(-k, 6.12323399573676603587e-17, r);                    // pi/2 lo
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06); ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03); ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07); pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03); pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)k & 3;
    const double ss = (q & 1) ? cr : sr, cc = (q & 1) ? sr : cr;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}
// builtin round(x, 2) of a Python float, returned in hundredths (exact decimal rounding, ties to even)
__device__ __forceinline__ double py_round2_hundredths(double x) {
    double hi = x * 100.0, lo = fma(x, 100.0, -hi), r = rint(hi), d = hi - r;
    if (d == 0.5 && lo > 0.0) r += 1.0; else if (d == -0.5 && lo < 0.0) r -= 1.0;
    return r;
}
__device__ __forceinline__ double py_mod360(double x) {  // drone.py:457, Python float %
    // CPython float_rem: m = fmod(x, 360); if (m != 0 && m < 0) m += 360; else if (m == 0) m = +0.0.
    // yaw + dyaw lives in (-360, 720) (yaw in [0, 360], |dyaw| <= 90), where fmod is one exact subtraction (x - 360 is exact
    // for x in [360, 720)) or the identity -- branch-free selects; anything else (a state loaded by the caller) takes libm's fmod.
    if (x > -360.0 && x < 720.0) {
        const double m = (x >= 360.0) ? x - 360.0 : x;     // fmod(x, 360), sign of x
        return (m < 0.0) ? m + 360.0 : m + 0.0;            // (-0.0 + 0.0 = +0.0; a tiny negative m rounds up to 360.0 exactly like CPython)
    }
    double m = rare_fmod(x, 360.0);
    if (m != 0.0) { if (m < 0.0) m += 360.0; } else m = 0.0;
    return m;
}

// ---- per-thread drone ---------------------------------------------------------------------------
struct Drone {
    double px, py, pz, vx, vy, vz, yaw, pitch, real_len, max_dev;
    int i, flags;   // waypoint index, FLAG_ARRIVE | FLAG_DEST
};

// per-drone scenario constants.  PTR: keep the two waypoint-table pointers in registers (small-N kernels, where the waypoint
// loads sit on the critical path) or rebuild them from a 32-bit row index at each use (many-neighbour kernels, where the
// 3 registers buy more independent range gates in flight).
template <bool PTR>
struct ConstsT;
template <>
struct ConstsT<true> {
    const double *wp_, *seg_;
    double radius, prio, route_len;
    int n;              // n_points
    __device__ __forceinline__ void set(const Params &P, int sn) { wp_ = P.wp + (size_t)sn * P.W * 3; seg_ = P.seg + (size_t)sn * P.W * 4; }
    __device__ __forceinline__ const double *wp(const Params &) const { return wp_; }     // [W][3]
    __device__ __forceinline__ const double *seg(const Params &) const { return seg_; }   // [W][4]: unit direction and length of the segment wp[i-1] -> wp[i]
};
template <>
struct ConstsT<false> {
    int snw;            // (scenario * N + drone) * W
    double radius, prio, route_len;
    int n;
    __device__ __forceinline__ void set(const Params &P, int sn) { snw = sn * P.W; }
    __device__ __forceinline__ const double *wp(const Params &P) const { return P.wp + (size_t)snw * 3; }
    __device__ __forceinline__ const double *seg(const Params &P) const { return P.seg + (size_t)snw * 4; }
};

__device__ __forceinline__ void load3(const double *p, double &x, double &y, double &z) { x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2); }

__device__ __forceinline__ bool within_goal_v(const Params &P, double px, double py, double pz, double qx, double qy, double qz) {
    const double fx = px - qx, fy = py - qy, fz = pz - qz;
    return sqrt_le_band(np_dot3(fx, fy, fz, fx, fy, fz), P.goal_thr, P.goal2_lo, P.goal2_hi);   // norm(p - q) <= 0.4, drone.py:172-192
}
__device__ __forceinline__ bool within_goal(const Params &P, double px, double py, double pz, const double *q) {
    double qx, qy, qz; load3(q, qx, qy, qz);
    const double fx = px - qx, fy = py - qy, fz = pz - qz;
    return sqrt_le_band(np_dot3(fx, fy, fz, fx, fy, fz), P.goal_thr, P.goal2_lo, P.goal2_hi);   // norm(p - q) <= 0.4, drone.py:172-192
}

// drone.calculate_deviation (drone.py:366-406).  `st` = previous waypoint; `sg` = (dx/|d|, dy/|d|, dz/|d|, |d|) of
// the segment previous -> current, computed once on the host with the reference's own operations.
__device__ __forceinline__ double deviation(double x1, double y1, double z1, double dxh, double dyh, double dzh, double mag, double x0, double y0, double z0) {
    if (mag == 0.0) return 0.0;
    const double px = x0 - x1, py = y0 - y1, pz = z0 - z1;
    const double t = px * dxh + py * dyh + pz * dzh;
    const double qx = x1 + t * dxh, qy = y1 + t * dyh, qz = z1 + t * dzh;
    const double ex = x0 - qx, ey = y0 - qy, ez = z0 - qz;
    return fast_sqrt(ex * ex + ey * ey + ez * ez);
}

// drone.cal_des_vel (drone.py:199-210, 319-352).  cos(atan2(fy,fx))*cos(atan2(fz,h)) == fx/|f| etc. up to a few
// ulp (<< the 1e-3 rounding quantum); h == 0 gives (0,0,+-1) on both routes.
__device__ __forceinline__ void cal_des_vel_v(const Params &P, double px, double py, double pz, double cx, double cy, double cz,
                                              double &d0, double &d1, double &d2) {
    const double fx = cx - px, fy = cy - py, fz = cz - pz;
    const double s2 = np_dot3(fx, fy, fz, fx, fy, fz);
    if (!sqrt_le_band(s2, P.goal_thr, P.goal2_lo, P.goal2_hi) && s2 == s2) {      // dis > goal_threshold
        const double id = P.vel_max * fast_rsqrt(s2);
        d0 = np_round3(fx * id); d1 = np_round3(fy * id); d2 = np_round3(fz * id);
    } else {
        d0 = d1 = d2 = 0.0;
    }
}

__device__ __forceinline__ void cal_des_vel(const Params &P, double px, double py, double pz, const double *cur, double &d0, double &d1, double &d2) {
    double cx, cy, cz; load3(cur, cx, cy, cz);
    cal_des_vel_v(P, px, py, pz, cx, cy, cz, d0, d1, d2);
}

// vel_obs3D.cal_vo_exp_tim (vel_obs3D.py:145-182)
__device__ __forceinline__ double cal_vo_exp_tim(double rx, double ry, double rz, double ovx, double ovy, double ovz, double ra, double rb) {
    double r = ra + rb;
    double vx = -ovx, vy = -ovy, vz = -ovz;
    double a = vx * vx + vy * vy + vz * vz;
    double b = 2.0 * rx * vx + 2.0 * ry * vy + 2.0 * rz * vz;
    double c = rx * rx + ry * ry + rz * rz - r * r;
    if (c <= 0.0) return 0.0;
    double temp = b * b - 4.0 * a * c;
    if (temp <= 0.0) return CUDART_INF;
    double sq = sqrt(temp), ia = fast_rcp(2.0 * a);
    double t1 = (-b + sq) * ia, t2 = (-b - sq) * ia;
    if (t1 < 0.0 && t2 < 0.0) return -1.0;
    double t3 = t1 >= 0.0 ? t1 : CUDART_INF, t4 = t2 >= 0.0 ? t2 : CUDART_INF;
    return t4 < t3 ? t4 : t3;
}

enum PairBits : int { PB_EVAL = 1, PB_COL = 2, PB_FULL = 4, PB_VO = 8 };

// Stage 1 of a pair (branch-free up to the rare exact-sqrt fallbacks): rvo_inter.preprocess neighbour test
// (rvo_inter.py:89-97), the collision test and the "moving away" gate of config_vo_circle2
// (rvo_inter.py:131-161).  `so` = the other drone's column of the exchange array (stride `st` doubles
// between the fields px py pz vx vy vz r prio).  Returns PB_EVAL if in range, |PB_COL on collision,
// |PB_FULL if the cone has to be evaluated (stage 2).
__device__ __forceinline__ int pair_gate(const Params &P, double x, double y, double z, double vx, double vy, double vz,
                                         double r, const double *so, int st) {
    const double mx = so[0], my = so[st], mz = so[2 * st], mr = so[6 * st];
    const double rel_x = mx - x, rel_y = my - y, rel_z = mz - z;      // == -(self - other), squares identical
    const double s2 = np_dot3(rel_x, rel_y, rel_z, rel_x, rel_y, rel_z);   // np.linalg.norm argument
    const bool same = (x == mx) & (y == my) & (z == mz);              // how the reference excludes "self"
    if (same || !sqrt_le_band(s2, P.nbr_range, P.nbr2_lo, P.nbr2_hi)) return 0;
    const double d2 = rel_y * rel_y + rel_x * rel_x + rel_z * rel_z;      // dis_mr^2, rvo_inter.py:135
    const double Rc = P.env_train ? r + mr : r - P.exp_radius + mr;
    if (sqrt_le(d2, Rc)) return PB_EVAL | PB_COL;
    const double dotp = vx * rel_x + vy * rel_y + vz * rel_z;
    return (dotp <= 0.0) ? PB_EVAL : (PB_EVAL | PB_FULL);
}

// Branch-free forms of pair_gate for the hot loop: straight-line arithmetic + predicate logic, so the gates of
// several neighbours overlap (no branch has to resolve between them).  PB_SLOW marks the (1e-14-wide) cases where a
// squared distance falls inside a guard band; the caller re-runs those few pairs through the exact pair_gate.
constexpr int PB_SLOW = 16, PB_SAME = 32;
// stage 1a: neighbour-range test only (rvo_inter.py:89-97) -> PB_EVAL (in range), PB_SLOW, or 0
__device__ __forceinline__ int pair_range_fast(const Params &P, double x, double y, double z, const double *so, int st) {
    const double mx = so[0], my = so[st], mz = so[2 * st];
    const double rel_x = mx - x, rel_y = my - y, rel_z = mz - z;
    const double s2 = np_dot3(rel_x, rel_y, rel_z, rel_x, rel_y, rel_z);
    const bool same = (x == mx) & (y == my) & (z == mz);
    const bool in = s2 < P.nbr2_lo, amb = !(in | (s2 > P.nbr2_hi));
    return same ? 0 : (amb ? PB_SLOW : (in ? PB_EVAL : 0));
}
// stage 1a in fp32 on float copies of the positions: ~3x cheaper than the fp64 gate and enough to settle every pair whose
// squared distance is not within 1e-4 (relative) of range^2 -- fp32 evaluation error is ~1e-6.  The rest (PB_SLOW) and the
// exact-equality "self" exclusion are decided in fp64 afterwards.
__device__ __forceinline__ int pair_range_f32(float x, float y, float z, const float4 &m, float lo, float hi) {
    const float rx = m.x - x, ry = m.y - y, rz = m.z - z;
    const float s2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
    return (s2 < lo) ? PB_EVAL : ((s2 > hi) ? 0 : PB_SLOW);      // NaN -> PB_SLOW -> exact path
}
// stage 1b, for a pair known to be in range: collision test + "moving away" gate (rvo_inter.py:131-161)
__device__ __forceinline__ int pair_near_fast(const Params &P, double x, double y, double z, double vx, double vy, double vz,
                                              double r, const double *so, int st) {
    const double mx = so[0], my = so[st], mz = so[2 * st], mr = so[6 * st];
    const double rel_x = mx - x, rel_y = my - y, rel_z = mz - z;
    const double d2 = rel_y * rel_y + rel_x * rel_x + rel_z * rel_z;
    const double Rc = P.env_train ? r + mr : r - P.exp_radius + mr, c2 = Rc * Rc;
    const double dotp = vx * rel_x + vy * rel_y + vz * rel_z;
    const bool same = (x == mx) & (y == my) & (z == mz);     // how the reference excludes "self" (rvo_inter.py:92)
    const bool col = d2 < c2 * (1.0 - 1e-14), amb = !(col | (d2 > c2 * (1.0 + 1e-14)));
    return same ? PB_SAME : (amb ? PB_SLOW : (col ? PB_COL : ((dotp > 0.0) ? PB_FULL : 0)));
}

// Stage 2 of a pair: the cone (rvo_inter.py:165-196, vel_obs3D.py:8-66,145-182), itself split in three so that the
// lanes of a warp never wait on each other's rare paths:
//   cone_screen  trig-free classification: certainly outside / certainly inside / ambiguous
//   cone_exact   asin / acos with the reference's roundings, only for ambiguous pairs
//   cone_time    expected collision time + the row's sort keys, only for pairs inside the cone
// (ax,ay,az) is the action with the |a|<1e-5 -> 0 substitution already applied.
struct PairGeom {
    double rel_x, rel_y, rel_z, s2, b2, dab, R;
};
__device__ __forceinline__ void pair_geom(const Params &P, double x, double y, double z, double vx, double vy, double vz,
                                          double r, double pra, const double *so, int st, double ax, double ay, double az, PairGeom &g) {
    const double mx = so[0], my = so[st], mz = so[2 * st], mr = so[6 * st];
    const double mvx = so[3 * st], mvy = so[4 * st], mvz = so[5 * st], prb = so[7 * st];
    g.rel_x = mx - x; g.rel_y = my - y; g.rel_z = mz - z;
    g.s2 = np_dot3(g.rel_x, g.rel_y, g.rel_z, g.rel_x, g.rel_y, g.rel_z);
    g.R = r + mr;
    const double pr = (pra == prb) ? 0.5 : fast_div(pra, pra + prb);      // get_PAA, vel_obs3D.py:19-32 (a/(a+a) == 0.5 exactly)
    const double PAAx = pr * (2.0 * x + (vx + mvx) * 1.0), PAAy = pr * (2.0 * y + (vy + mvy) * 1.0), PAAz = pr * (2.0 * z + (vz + mvz) * 1.0);
    const double Bx = x + 2.0 * ax * P.delta_t - PAAx, By = y + 2.0 * ay * P.delta_t - PAAy, Bz = z + 2.0 * az * P.delta_t - PAAz;
    g.b2 = np_dot3(Bx, By, Bz, Bx, By, Bz); g.dab = np_dot3(g.rel_x, g.rel_y, g.rel_z, Bx, By, Bz);
}
// inside <=> round(100 alpha) > round(100 beta), alpha = asin(R/|rel|), beta = acos(c).  beta > alpha + 0.02 rad is
// certainly outside and beta < alpha - 0.02 certainly inside (each rounding moves a value by <= 0.005 rad); both are
// decided on cosines without trig or square roots: with s = sin a, c = cos beta,
//   beta > a + .02  <=>  c < cos(a + .02)  <=>  c + s S < C sqrt(1 - s^2)      (S = sin .02, C = cos .02)
//   beta < a - .02  <=>  c - s S > C sqrt(1 - s^2)
// Returns 0 outside, 1 inside, 2 ambiguous (the +-0.02 band, degenerate |B| = 0, and near-collinear vectors where the
// reference's unclamped arccos may return NaN).
__device__ __forceinline__ int cone_screen(const PairGeom &g) {
    if (!(g.b2 > 0.0) || !(g.s2 > g.R * g.R)) return 2;      // asin(R/|rel| > 1) = NaN in the reference: let cone_exact decide
    const double inab = fast_rsqrt(g.s2), s = g.R * inab, c = g.dab * inab * fast_rsqrt(g.b2);
    const double C02 = 0.99980000666657776, S02 = 0.019998666693333084;
    const double rhs2 = fma(-s, s, 1.0) * (C02 * C02), sS = s * S02;
    const double lo = c + sS, hi = c - sS;
    if (lo < -1e-9 || (lo >= 0.0 && lo * lo < rhs2 * (1.0 - 1e-9))) return 0;
    if (s > 0.03 && hi > 1e-9 && hi * hi > rhs2 * (1.0 + 1e-9) && c < 1.0 - 1e-9) return 1;
    return 2;
}
__device__ __forceinline__ bool cone_exact(const PairGeom &g, double &alpha_h) {
    const double nab = sqrt(g.s2);
    alpha_h = py_round2_hundredths(rare_asin(g.R / nab));       // get_alpha, vel_obs3D.py:8-17 (IEEE ops: rare path)
    const double AB = nab * sqrt(g.b2);
    const double cosang = (AB != 0.0) ? g.dab / AB : 0.0;
    const double beta_h = rint(rare_acos(cosang) * 100.0); // get_beta: unclamped arccos, numpy rounding
    return alpha_h > beta_h;                               // NaN beta -> outside, as in the reference
}
// Returns true when the expected collision time is below the threshold (vo_flag); then t_out, md, iet are valid.
__device__ __forceinline__ bool cone_time(const Params &P, double vx, double vy, double vz, double r, const double *so, int st,
                                          double ax, double ay, double az, const PairGeom &g, double &t_out, double &md, double &iet) {
    const double mvx = so[3 * st], mvy = so[4 * st], mvz = so[5 * st], mr = so[6 * st];
    const double rel_vx = 2.0 * ax - mvx - vx, rel_vy = 2.0 * ay - mvy - vy, rel_vz = 2.0 * az - mvz - vz;
    const double exp_time = cal_vo_exp_tim(g.rel_x, g.rel_y, g.rel_z, rel_vx, rel_vy, rel_vz, r, mr);
    if (!(exp_time < P.ctime)) return false;
    const double d2 = g.rel_y * g.rel_y + g.rel_x * g.rel_x + g.rel_z * g.rel_z;
    t_out = exp_time;
    md = sqrt(d2) - mr;
    iet = fast_div(1.0, exp_time + 0.2);
    return true;
}

// ---- top-nm neighbour list (rvo_inter.py:50-59) ------------------------------------------------------
// Kept sorted most-urgent-first; "more urgent" = later in the reference's stable descending sort by
// (-iet, md): larger iet, then smaller md, then larger drone index.
// MAXR <= 4 (N <= 5): registers, fully unrolled predicated insertion.  Larger lists live in shared memory
// (SoA [slot][thread], conflict-free) -- in registers they would spill (50+ registers for MAXR = 10).
__device__ __forceinline__ bool more_urgent(double a_iet, double a_md, int a_j, double b_iet, double b_md, int b_j) {
    return (a_iet > b_iet) || (a_iet == b_iet && (a_md < b_md || (a_md == b_md && a_j > b_j)));
}
template <int MAXR, bool IN_SMEM = (MAXR > 4)>
struct TopList;

template <int MAXR>
struct TopList<MAXR, false> {
    __host__ __device__ static constexpr size_t smem_bytes(int) { return 0; }
    double iet[MAXR], md[MAXR];
    int j[MAXR], ah[MAXR];
    int cnt;
    __device__ __forceinline__ void init(void *, int, int) {}
    __device__ __forceinline__ void clear() { cnt = 0; }
    __device__ __forceinline__ void insert(int cap, double c_iet, double c_md, int c_j, int c_ah) {
        bool placed = false;
#pragma unroll
        for (int s = 0; s < MAXR; s++) {
            bool valid = s < cnt;
            bool sw = valid && more_urgent(c_iet, c_md, c_j, iet[s], md[s], j[s]);
            if (sw) {
                double t0 = iet[s]; iet[s] = c_iet; c_iet = t0;
                double t1 = md[s]; md[s] = c_md; c_md = t1;
                int t2 = j[s]; j[s] = c_j; c_j = t2;
                int t3 = ah[s]; ah[s] = c_ah; c_ah = t3;
            }
            if (!valid && !placed && s < cap) { iet[s] = c_iet; md[s] = c_md; j[s] = c_j; ah[s] = c_ah; placed = true; }
        }
        if (placed) cnt++;
    }
    template <typename F>
    __device__ __forceinline__ void for_each(F &&f) const {   // f(slot, iet, md, j, ah) for slot = 0 (most urgent) .. cnt-1
        if (cnt == 0) return;                                  // the common case in sparse airspace: one branch, not MAXR predicated bodies
#pragma unroll
        for (int s = 0; s < MAXR; s++) if (s < cnt) f(s, iet[s], md[s], j[s], ah[s]);
    }
};

template <int MAXR>
struct TopList<MAXR, true> {
    __host__ __device__ static constexpr size_t smem_bytes(int threads) { return (size_t)threads * MAXR * 24; }
    double *iet, *md; int *j, *ah;   // this thread's column; slot s at [s * st]
    int st, cnt;
    __device__ __forceinline__ void init(void *base, int threads, int tid) {
        st = threads;
        iet = reinterpret_cast<double *>(base) + tid; md = iet + (size_t)MAXR * threads;
        j = reinterpret_cast<int *>(reinterpret_cast<double *>(base) + 2 * (size_t)MAXR * threads) + tid; ah = j + (size_t)MAXR * threads;
    }
    __device__ __forceinline__ void clear() { cnt = 0; }
    __device__ __forceinline__ void insert(int cap, double c_iet, double c_md, int c_j, int c_ah) {
        int s = cnt;                                   // walk up from the least urgent end, shifting entries down
        while (s > 0 && more_urgent(c_iet, c_md, c_j, iet[(s - 1) * st], md[(s - 1) * st], j[(s - 1) * st])) {
            if (s < cap) { iet[s * st] = iet[(s - 1) * st]; md[s * st] = md[(s - 1) * st]; j[s * st] = j[(s - 1) * st]; ah[s * st] = ah[(s - 1) * st]; }
            s--;
        }
        if (s < cap) { iet[s * st] = c_iet; md[s * st] = c_md; j[s * st] = c_j; ah[s * st] = c_ah; }
        if (cnt < cap) cnt++;
    }
    template <typename F>
    __device__ __forceinline__ void for_each(F &&f) const {
        for (int s = 0; s < cnt; s++) f(s, iet[s * st], md[s * st], j[s * st], ah[s * st]);
    }
};

// ir_gym.calculate_angle_between_vectors, 2nd definition (ir_gym.py:447-473), folded with the angle
// buckets of rvo_reward_cal (ir_gym.py:90-100): theta = acos(c) is only compared with pi/18, pi/6, pi/3,
// pi/2, and acos is monotone, so the bucket is read off c against the cosines of those angles
// (thresholds found by bisection against the libm the reference calls).  NaN -> -4 like the reference's else-branch.
__device__ __forceinline__ double angle_punish(double A0, double A1, double A2, double B0, double B1, double B2) {
    const double eps = 1e-8;
    const double mA2 = A0 * A0 + A1 * A1 + A2 * A2 + eps, mB2 = B0 * B0 + B1 * B1 + B2 * B2 + eps;
    const double dp = A0 * B0 + A1 * B1 + A2 * B2;
    const double c = np_clip(dp * fast_rsqrt(mA2 * mB2), -1.0 + eps, 1.0 - eps);   // dp / (|A| |B|) to ~2 ulp
    // smallest doubles c with glibc acos(c) < pi/18, pi/6, pi/3, pi/2 (bisection over doubles)
    if (c >= 0x1.f838b8c811c18p-1) return 3.0;
    if (c >= 0x1.bb67ae8584cabp-1) return 1.0;
    if (c >= 0x1.0000000000002p-1) return 0.5;
    if (c >= 0x1.8d313198a2e04p-53) return 0.0;
    return -4.0;
}

template <int MAXR, int MAXT, bool WL, bool SL>
__global__ void uaisa_env_kernel(const __grid_constant__ Params P, const __grid_constant__ Launch L);
template <int MAXR, int MAXT, bool WL, bool SL>      // the same kernel compiled under a 128-register cap (many-neighbour instantiations without slices)
__global__ void uaisa_env_kernel_d(const __grid_constant__ Params P, const __grid_constant__ Launch L);
__global__ void uaisa_cal_des_kernel(const __grid_constant__ Params P, float *des);
__global__ void uaisa_rvo_select_kernel(const __grid_constant__ Params P, double vmax, double acceler, float *vel_out, float *ctrl_out);
__global__ void uaisa_action_transform_kernel(const __grid_constant__ Params P, const float *a_inc, float acceler, float *a_inc_rounded, float *action);
__global__ void uaisa_pack_kernel(int A, int obs_dim, int nm, const float *obs, const uint8_t *count, const int *nbr, float *self_obs,
                                  void *rows_v, int *n_rows, int cap);
__global__ void uaisa_gae_kernel(const float *rew, const float *val, const uint8_t *cut, int T, int A, float gamma, float lam, float *adv, float *ret);

}  // namespace uaisa

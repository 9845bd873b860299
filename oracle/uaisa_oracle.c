/*
 * uaisa_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, fp64 restatement of the reference's live env step
 *   mdin.drone_step            uaisa_env/drone_envs/mdin.py:19-30
 *   ir_gym.*                   uaisa_env/drone_envs/ir_gym.py:50-133,136-311,334-383,448-490
 *   env_base.drone_step        uaisa_env/drone_envs/env_base.py:135-144
 *   drone.*                    uaisa_env/drone_envs/drone.py:86-129,145-225,254-297,319-352,366-406,431-490
 *   rvo_inter.*                uaisa_env/vel_obs/rvo_inter.py:20-107,116-228
 *   vel_obs3D.*                uaisa_env/vel_obs/vel_obs3D.py:8-32,44-66,145-202
 * It is the checker for the CUDA path and the CPU baseline that bench.py times; it is NOT shipped
 * and the product library never links it.  Parity pinned: see uaisa_oracle.h.
 *
 * Floating-point fidelity notes (all verified against the reference running under numpy 2.3.5 /
 * glibc in the build container):
 *   - np.dot / np.linalg.norm on 3- and 2-vectors evaluate as an FMA chain
 *     fma(x2,y2,fma(x1,y1,x0*y0)) (OpenBLAS ddot tail loop)            -> np_dot3 / np_norm3 / np_norm2
 *   - `x ** 2` on a Python float or an np.float64 scalar calls libm pow(x, 2.0), which is not
 *     always the correctly rounded square                               -> SQ()
 *   - builtin round(float, 2) (vel_obs3D.py:15) is exact decimal rounding; round(np.float64, 2)
 *     (vel_obs3D.py:64) and np.round are rint(x*10^k)/10^k               -> uo_pyround2 / uo_npround
 * Compile with -ffp-contract=off so that nothing else is fused.
 */
#include "uaisa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OBS_SELF 12
#define ROW 9
#define SQ(x) pow((x), 2.0)

struct UoBatch {
    UoConfig cfg;
    int E, N, W, M, S;
    double *wp; int32_t *npts; double *route_len; double *radius; double *prio;
    double *bld; int32_t *nbld; double *map; int32_t *env2scn;
    double *pos, *vel, *yaw, *pitch, *real_len, *max_dev, *extra_len;
    int32_t *wp_idx; uint8_t *arrive, *dest;
    int64_t cnt_in_range, cnt_cone;
};

static int g_threads = 0;
void uo_set_threads(int n) { g_threads = n; }
int uo_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static int nthreads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

void uo_default_config(UoConfig *c) {
    c->goal_threshold = 0.4; c->max_acc = 1.0; c->max_angle_change = 90.0; c->dt = 1.0;
    c->vel_max = 1.0; c->neighbor_range = 10.0; c->building_range = 5.0; c->building_zgate = 2.0;
    c->ctime_threshold = 2.0; c->delta_t = 1.0; c->exp_radius = 0.2; c->env_train = 1; c->nm = 10;
}

/* ---- numpy / python arithmetic primitives --------------------------------------------------- */
static inline double np_dot3(const double *a, const double *b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
static inline double np_norm3(const double *a) { return sqrt(np_dot3(a, a)); }
static inline double np_norm2(const double *a) { return sqrt(fma(a[1], a[1], a[0] * a[0])); }
static inline double np_clip(double x, double lo, double hi) { return (x != x) ? x : fmin(fmax(x, lo), hi); } /* np.clip propagates NaN */

double uo_npround(double x, int nd) {
    static const double p10[4] = {1.0, 10.0, 100.0, 1000.0};
    return rint(x * p10[nd]) / p10[nd];
}
double uo_pyround2(double x) {
    if (!isfinite(x)) return x;
    double hi = x * 100.0, lo = fma(x, 100.0, -hi), r = rint(hi), d = hi - r;
    if (d == 0.5 && lo > 0.0) r += 1.0; else if (d == -0.5 && lo < 0.0) r -= 1.0;
    return r / 100.0;
}
/* Python / numpy float modulo (drone.py:457): result takes the sign of the divisor */
static inline double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) { if ((b < 0.0) != (m < 0.0)) m += b; } else m = copysign(0.0, b);
    return m;
}
static inline double wraptopi(double t) { /* vel_obs3D.py:195-202 */
    if (t > M_PI) t = t - 2 * M_PI;
    if (t < -M_PI) t = t + 2 * M_PI;
    return t;
}

/* ---- per-agent accessors ---------------------------------------------------------------------- */
static inline const double *WP(const UoBatch *b, int s, int k, int w) { return b->wp + (((size_t)s * b->N + k) * b->W + w) * 3; }

/* drone.arrive / destination_arrive distance test (drone.py:172-192) */
static inline int within_goal(const UoBatch *b, const double *p, const double *q) {
    double d[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
    return np_norm3(d) <= b->cfg.goal_threshold;
}

/* drone.calculate_deviation (drone.py:366-406) */
static double deviation(const double *st, const double *en, const double *p) {
    double dx = en[0] - st[0], dy = en[1] - st[1], dz = en[2] - st[2];
    double mag = sqrt(SQ(dx) + SQ(dy) + SQ(dz));
    if (mag == 0) return 0;
    double dxh = dx / mag, dyh = dy / mag, dzh = dz / mag;
    double px = p[0] - st[0], py = p[1] - st[1], pz = p[2] - st[2];
    double t = px * dxh + py * dyh + pz * dzh;
    double qx = st[0] + t * dxh, qy = st[1] + t * dyh, qz = st[2] + t * dzh;
    return sqrt(SQ(p[0] - qx) + SQ(p[1] - qy) + SQ(p[2] - qz));
}

/* drone.cal_des_vel (drone.py:199-210) + relative (:340-352) + angles_to_direction (:319-328) */
static void cal_des_vel(const UoBatch *b, const double *p, const double *cur, double *des) {
    double dif[3] = {cur[0] - p[0], cur[1] - p[1], cur[2] - p[2]};
    double dis = np_norm3(dif);
    if (dis > b->cfg.goal_threshold) {
        double az = atan2(dif[1], dif[0]);
        double el = atan2(dif[2], np_norm2(dif));
        double dir[3] = {cos(az) * cos(el), sin(az) * cos(el), sin(el)};
        for (int k = 0; k < 3; k++) des[k] = uo_npround(b->cfg.vel_max * dir[k], 3);
    } else {
        des[0] = des[1] = des[2] = 0.0;
    }
}

/* drone.dronestate (drone.py:254-263): 12-vector, side effect on max_deviation */
static void dronestate(UoBatch *b, int s, int a, int k, double *S12) {
    const double *p = b->pos + 3 * (size_t)a, *v = b->vel + 3 * (size_t)a;
    int i = b->wp_idx[a];
    const double *cur = WP(b, s, k, i), *prev = WP(b, s, k, i - 1);
    cal_des_vel(b, p, cur, S12 + 8);
    double dev = deviation(prev, cur, p);
    if (dev > b->max_dev[a]) b->max_dev[a] = dev;
    S12[0] = p[0]; S12[1] = p[1]; S12[2] = p[2]; S12[3] = v[0]; S12[4] = v[1]; S12[5] = v[2];
    S12[6] = b->radius[(size_t)s * b->N + k]; S12[7] = b->prio[(size_t)s * b->N + k];
    S12[11] = dev;
}

/* vel_obs3D.cal_vo_exp_tim (vel_obs3D.py:145-182) */
static double cal_vo_exp_tim(double rx, double ry, double rz, double ovx, double ovy, double ovz, double ra, double rb) {
    double r = ra + rb;
    double vx = -ovx, vy = -ovy, vz = -ovz;
    double a = SQ(vx) + SQ(vy) + SQ(vz);
    double bb = 2 * rx * vx + 2 * ry * vy + 2 * rz * vz;
    double c = SQ(rx) + SQ(ry) + SQ(rz) - SQ(r);
    if (c <= 0) return 0;
    double temp = SQ(bb) - 4 * a * c;
    double t;
    if (temp <= 0) {
        t = INFINITY;
    } else {
        double t1 = (-bb + sqrt(temp)) / (2 * a);
        double t2 = (-bb - sqrt(temp)) / (2 * a);
        if (t1 < 0 && t2 < 0) return -1;
        double t3 = t1 >= 0 ? t1 : INFINITY;
        double t4 = t2 >= 0 ? t2 : INFINITY;
        t = t4 < t3 ? t4 : t3; /* python min(t3, t4) */
    }
    return t;
}

typedef struct { double row[ROW]; int vo; double t; int col; double md; int full; } PairOut;

/* rvo_inter.config_vo_circle2 (rvo_inter.py:116-196) */
static void vo_pair(const UoConfig *cfg, const double *sa, const double *sb, const double *action, PairOut *o) {
    double act[3] = {action[0], action[1], action[2]};
    if (np_norm3(act) < 1e-5) act[0] = act[1] = act[2] = 0.0;
    double x = sa[0], y = sa[1], z = sa[2], vx = sa[3], vy = sa[4], vz = sa[5], r = sa[6];
    double mx = sb[0], my = sb[1], mz = sb[2], mvx = sb[3], mvy = sb[4], mvz = sb[5], mr = sb[6];
    o->vo = 0; o->col = 0; o->full = 0;
    double rel_x = mx - x, rel_y = my - y, rel_z = mz - z;
    double dis_mr = sqrt(SQ(rel_y) + SQ(rel_x) + SQ(rel_z));
    double real_dis_mr = dis_mr;
    if (cfg->env_train) {
        if (dis_mr <= r + mr) { dis_mr = r + mr; o->col = 1; }
    } else {
        if (dis_mr <= r - cfg->exp_radius + mr) o->col = 1;
        if (dis_mr <= r + mr) dis_mr = r + mr;
    }
    o->row[0] = x; o->row[1] = y; o->row[2] = z; o->row[3] = rel_x; o->row[4] = rel_y; o->row[5] = rel_z;
    if (o->col) { o->row[6] = 0; o->row[7] = 0; o->row[8] = 0; o->t = 0; o->md = dis_mr; return; }
    double dotp = vx * rel_x + vy * rel_y + vz * rel_z;
    if (dotp <= 0) { o->row[6] = 0; o->row[7] = -1; o->row[8] = -1; o->t = 0; o->md = dis_mr; return; }
    o->full = 1;
    /* get_alpha (vel_obs3D.py:8-17) */
    double ab[3] = {mx - x, my - y, mz - z};
    double alpha = uo_pyround2(wraptopi(asin((r + mr) / np_norm3(ab))));
    /* get_PAA (vel_obs3D.py:19-32) */
    double pr = sa[7] / (sa[7] + sb[7]);
    double PAA[3] = {pr * (2 * x + (vx + mvx) * 1), pr * (2 * y + (vy + mvy) * 1), pr * (2 * z + (vz + mvz) * 1)};
    double rel_vx = 2 * act[0] - mvx - vx, rel_vy = 2 * act[1] - mvy - vy, rel_vz = 2 * act[2] - mvz - vz;
    /* vo_out_jud_vector (rvo_inter.py:212-228) + get_beta (vel_obs3D.py:44-66) */
    double A[3] = {rel_x, rel_y, rel_z};
    double B[3] = {x + 2 * act[0] * cfg->delta_t - PAA[0], y + 2 * act[1] * cfg->delta_t - PAA[1], z + 2 * act[2] * cfg->delta_t - PAA[2]};
    double dab = np_dot3(A, B), AB = np_norm3(A) * np_norm3(B);
    double cosang = (AB != 0) ? dab / AB : 0.0;
    double beta = uo_npround(wraptopi(acos(cosang)), 2);
    double exp_time = INFINITY;
    if (alpha > beta) { /* inside the cone */
        exp_time = cal_vo_exp_tim(rel_x, rel_y, rel_z, rel_vx, rel_vy, rel_vz, r, mr);
        if (exp_time < cfg->ctime_threshold) o->vo = 1; else exp_time = INFINITY;
    }
    o->row[0] = PAA[0]; o->row[1] = PAA[1]; o->row[2] = PAA[2];
    o->row[6] = alpha; o->row[7] = real_dis_mr - mr; o->row[8] = 1 / (exp_time + 0.2);
    o->t = exp_time; o->md = real_dis_mr - mr;
}

int uo_pair(const UoConfig *cfg, const double *sa8, const double *sb8, const double *act, double *out9, double *t, double *md) {
    PairOut o; vo_pair(cfg, sa8, sb8, act, &o);
    memcpy(out9, o.row, sizeof(o.row)); *t = o.t; *md = o.md;
    return o.vo | (o.col << 1);
}

/* rvo_inter.preprocess neighbour test (rvo_inter.py:89-97) */
static inline int in_neighbour_set(const UoConfig *cfg, const double *sa, const double *sb) {
    if (sa[0] == sb[0] && sa[1] == sb[1] && sa[2] == sb[2]) return 0;
    double dif[3] = {sa[0] - sb[0], sa[1] - sb[1], sa[2] - sb[2]};
    return np_norm3(dif) <= cfg->neighbor_range;
}

typedef struct { double row[ROW]; int j; } Row;

/* rvo_inter.config_vo_inf (rvo_inter.py:20-61). S12 = all states of the env; ego = k. */
static void config_vo_inf(UoBatch *b, int s, const double *S12all, int k, const double *act, Row *rows, int *nrows,
                          int *vo_flag, double *min_exp_time, int *collision, int64_t *c_rng, int64_t *c_cone) {
    const UoConfig *cfg = &b->cfg;
    const double *sa = S12all + 12 * k;
    int N = b->N, n = 0, col = 0, vo = 0;
    double mt = INFINITY;
    /* buildings: preprocess filter (rvo_inter.py:99-105) then check_col_with_budilding (:198-209) */
    for (int m = 0; m < b->nbld[s]; m++) {
        const double *bd = b->bld + ((size_t)s * b->M + m) * 4;
        if (bd[2] > sa[2] - cfg->building_zgate) {
            double diff[2] = {sa[0] - bd[0], sa[1] - bd[1]};
            if (np_norm2(diff) <= cfg->building_range) {
                if (sa[2] <= bd[2]) {
                    double dis = sqrt(SQ(sa[0] - bd[0]) + SQ(sa[1] - bd[1]));
                    if (dis <= sa[6] + bd[3]) col = 1;
                }
            }
        }
    }
    Row *all = (Row *)malloc(sizeof(Row) * (size_t)(N > 1 ? N : 1));
    for (int j = 0; j < N; j++) {
        if (j == k) continue;
        const double *sb = S12all + 12 * j;
        if (!in_neighbour_set(cfg, sa, sb)) continue;
        PairOut o; vo_pair(cfg, sa, sb, act, &o);
        (*c_rng)++; (*c_cone) += o.full;
        if (o.vo) { memcpy(all[n].row, o.row, sizeof(o.row)); all[n].j = j; n++; vo = 1; if (o.t < mt) mt = o.t; }
        if (o.col) col = 1;
    }
    /* list.sort(reverse=True, key=(-row[8], row[7])) is stable: equal keys keep original order
     * (rvo_inter.py:50); insertion sort, descending, stable */
    for (int a = 1; a < n; a++) {
        Row t = all[a]; int c = a - 1;
        while (c >= 0) {
            double k1c = -all[c].row[8], k1t = -t.row[8];
            int less = (k1c < k1t) || (k1c == k1t && all[c].row[7] < t.row[7]); /* key(c) < key(t) */
            if (!less) break;
            all[c + 1] = all[c]; c--;
        }
        all[c + 1] = t;
    }
    int nm = cfg->nm, first = n > nm ? n - nm : 0, kept = 0;
    if (nm > 0) for (int a = first; a < n; a++) rows[kept++] = all[a];
    free(all);
    *nrows = kept; *vo_flag = vo; *min_exp_time = mt; *collision = col;
}

/* ir_gym.calculate_angle_between_vectors, 2nd definition wins (ir_gym.py:447-473) */
static double angle_between(const double *A, const double *B) {
    double eps = 1e-8;
    double magA = sqrt(SQ(A[0]) + SQ(A[1]) + SQ(A[2]) + eps);
    double magB = sqrt(SQ(B[0]) + SQ(B[1]) + SQ(B[2]) + eps);
    double dp = A[0] * B[0] + A[1] * B[1] + A[2] * B[2];
    if (magA < 1e-6 || magB < 1e-6) return 0.0;
    double c = dp / (magA * magB);
    c = np_clip(c, -1.0 + eps, 1.0 - eps);
    return acos(c);
}

/* ir_gym.rvo_reward_cal (ir_gym.py:64-133) on top of rvo_inter.config_vo_reward (rvo_inter.py:63-83) */
static double rvo_reward(UoBatch *b, const double *S12all, int k, const double *act, int64_t *c_rng, int64_t *c_cone) {
    const UoConfig *cfg = &b->cfg;
    const double *sa = S12all + 12 * k;
    int vo = 0; double mt = INFINITY;
    for (int j = 0; j < b->N; j++) {
        if (j == k) continue;
        const double *sb = S12all + 12 * j;
        if (!in_neighbour_set(cfg, sa, sb)) continue;
        PairOut o; vo_pair(cfg, sa, sb, act, &o);
        (*c_rng)++; (*c_cone) += o.full;
        if (o.vo) { vo = 1; if (o.t < mt) mt = o.t; }
    }
    double des[3] = {uo_npround(sa[8], 3), uo_npround(sa[9], 3), uo_npround(sa[10], 3)};
    double vel_penalty = 0.2 * np_norm3(act) / np_norm3(des);
    double ang = angle_between(des, act), angle_punish;
    if (-M_PI / 18 < ang && ang < M_PI / 18) angle_punish = 3;
    else if (-M_PI / 6 < ang && ang < M_PI / 6) angle_punish = 1;
    else if (-M_PI / 3 < ang && ang < M_PI / 3) angle_punish = 0.5;
    else if (-M_PI / 2 < ang && ang < M_PI / 2) angle_punish = 0;
    else angle_punish = -4;
    double safety = 0;
    if (vo) {
        double urgency = 0;
        if (mt < 2) urgency = -8.0 * exp(-mt / 0.5);
        safety = -2.5 + urgency;
    }
    return uo_npround(angle_punish + vel_penalty + safety, 3);
}

/* drone.move_forward with stop=truthy (env_base.py:142 positional mismatch; drone.py:96-129,431-490) */
static void move_forward(UoBatch *b, int s, int a, int k, const double *act) {
    const UoConfig *cfg = &b->cfg;
    double *p = b->pos + 3 * (size_t)a, *v = b->vel + 3 * (size_t)a;
    double velocity = np_norm3(v);
    double acc = np_clip(act[0] * cfg->max_acc, -cfg->max_acc, cfg->max_acc);
    double dyaw = np_clip(act[1] * cfg->max_angle_change, -cfg->max_angle_change, cfg->max_angle_change);
    double dpitch = np_clip(act[2] * cfg->max_angle_change, -cfg->max_angle_change, cfg->max_angle_change);
    double nv = velocity + acc * 1; nv = (0.0 > nv) ? 0.0 : nv; /* python max(nv, 0) */
    double yaw = py_mod(b->yaw[a] + dyaw, 360.0);
    double pitch = np_clip(b->pitch[a] + dpitch, -90, 90);
    b->yaw[a] = yaw; b->pitch[a] = pitch;
    double yr = yaw * (M_PI / 180.0), prd = pitch * (M_PI / 180.0); /* np.deg2rad */
    double nvel[3] = {nv * cos(prd) * cos(yr), nv * cos(prd) * sin(yr), nv * sin(prd)};
    if (b->dest[a]) nvel[0] = nvel[1] = nvel[2] = 0.0;
    double np_[3] = {p[0] + nvel[0] * cfg->dt, p[1] + nvel[1] * cfg->dt, p[2] + nvel[2] * cfg->dt};
    double dif[3] = {np_[0] - p[0], np_[1] - p[1], np_[2] - p[2]};
    b->real_len[a] = b->real_len[a] + np_norm3(dif);
    p[0] = np_[0]; p[1] = np_[1]; p[2] = np_[2]; v[0] = nvel[0]; v[1] = nvel[1]; v[2] = nvel[2];
    int i = b->wp_idx[a], n = b->npts[(size_t)s * b->N + k];
    if (within_goal(b, p, WP(b, s, k, i))) {
        int at_dest = within_goal(b, p, WP(b, s, k, n - 1));
        if (at_dest) b->extra_len[a] = b->real_len[a] - b->route_len[(size_t)s * b->N + k]; /* drone.py:188 */
        if (!at_dest && i < n - 1) { b->wp_idx[a] = i + 1; b->arrive[a] = 0; }
    }
}

/* ir_gym.mov_reward (ir_gym.py:256-311) + calculate_penalty_with_exp (:476-490) */
static double mov_reward(int collision, int arr_now, int waypoint_num, int n_points, int dest_now, double dev, int len_flag, double exlen) {
    if (collision) return -50;
    double reward = 0;
    if (arr_now) reward += 3.0 * pow(0.95, (double)(n_points - waypoint_num));
    if (dest_now) reward += 20.0;
    double d = dev * 10;
    double dev_pen = -1.5 * (2 / (1 + exp(-(d - 5) / 0.3)));
    double ex_pen = 0;
    if (len_flag) {
        ex_pen = -0.3 * log(exlen + 1 + 1e-6);
        if (ex_pen < -6 || ex_pen != ex_pen) ex_pen = -6;
    }
    return uo_npround(reward + dev_pen + ex_pen, 3);
}

static void write_obs(const UoConfig *cfg, const double *S12, const Row *rows, int nrows, double *obs, int32_t *count, int32_t *nbr) {
    int L = OBS_SELF + ROW * cfg->nm;
    for (int q = 0; q < L; q++) obs[q] = 0.0;
    for (int q = 0; q < OBS_SELF; q++) obs[q] = uo_npround(S12[q], 2);
    for (int r = 0; r < nrows; r++) for (int q = 0; q < ROW; q++) obs[OBS_SELF + ROW * r + q] = uo_npround(rows[r].row[q], 2);
    *count = nrows;
    for (int r = 0; r < cfg->nm; r++) nbr[r] = r < nrows ? rows[r].j : -1;
}

static void step_env(UoBatch *b, int e, const double *action, double *obs, int32_t *count, int32_t *nbr,
                     double *reward, uint8_t *done, uint8_t *info, uint8_t *finish, int64_t *c_rng, int64_t *c_cone) {
    int N = b->N, s = b->env2scn[e], base = e * N, nm = b->cfg.nm, L = OBS_SELF + ROW * nm;
    double *S12 = (double *)malloc(sizeof(double) * 12 * (size_t)N);
    double *rA = (double *)malloc(sizeof(double) * (size_t)N);
    Row *rows = (Row *)malloc(sizeof(Row) * (size_t)(nm > 0 ? nm : 1));
    /* PASS A (ir_gym.py:50-62) */
    for (int k = 0; k < N; k++) dronestate(b, s, base + k, k, S12 + 12 * k);
    for (int k = 0; k < N; k++) rA[k] = rvo_reward(b, S12, k, action + 3 * (size_t)(base + k), c_rng, c_cone);
    /* MOVE (env_base.py:135-144) */
    for (int k = 0; k < N; k++) move_forward(b, s, base + k, k, action + 3 * (size_t)(base + k));
    /* PASS B (ir_gym.py:136-254) */
    for (int k = 0; k < N; k++) dronestate(b, s, base + k, k, S12 + 12 * k);
    for (int k = 0; k < N; k++) {
        int a = base + k;
        const double *p = b->pos + 3 * (size_t)a;
        int i = b->wp_idx[a], n = b->npts[(size_t)s * N + k];
        int arr_now = 0, dest_now = 0;
        if (within_goal(b, p, WP(b, s, k, i)) && !b->arrive[a]) { b->arrive[a] = 1; arr_now = 1; }
        if (b->arrive[a]) {
            int at_dest = within_goal(b, p, WP(b, s, k, n - 1));
            if (at_dest) b->extra_len[a] = b->real_len[a] - b->route_len[(size_t)s * N + k];
            if (at_dest && !b->dest[a]) { b->dest[a] = 1; dest_now = 1; }
        }
        double dev = S12[12 * k + 11];
        double exlen = b->real_len[a] - b->route_len[(size_t)s * N + k] + 4;
        int len_flag = exlen > 0;
        int nrows, vo, col; double mt;
        config_vo_inf(b, s, S12, k, action + 3 * (size_t)a, rows, &nrows, &vo, &mt, &col, c_rng, c_cone);
        const double *map = b->map + 3 * (size_t)s; /* drone.drone_out_map (drone.py:213-225) */
        if (p[0] < 0 || p[0] > map[0] || p[1] < 0 || p[1] > map[1] || p[2] < 0 || p[2] > map[2]) col = 1;
        write_obs(&b->cfg, S12 + 12 * k, rows, nrows, obs + (size_t)a * L, count + a, nbr + (size_t)a * nm);
        double rB = mov_reward(col, arr_now, i, n - 1, dest_now, dev, len_flag, exlen);
        reward[a] = rA[k] + rB;
        done[a] = (uint8_t)col; info[a] = b->arrive[a]; finish[a] = b->dest[a];
    }
    free(S12); free(rA); free(rows);
}

static void observe_env(UoBatch *b, int e, double *obs, int32_t *count, int32_t *nbr) {
    int N = b->N, s = b->env2scn[e], base = e * N, nm = b->cfg.nm, L = OBS_SELF + ROW * nm;
    double *S12 = (double *)malloc(sizeof(double) * 12 * (size_t)N);
    Row *rows = (Row *)malloc(sizeof(Row) * (size_t)(nm > 0 ? nm : 1));
    const double zero[3] = {0, 0, 0};
    int64_t d0 = 0, d1 = 0;
    /* env_observation: total_states() then observation(drone, others) which calls dronestate again */
    for (int k = 0; k < N; k++) dronestate(b, s, base + k, k, S12 + 12 * k);
    for (int k = 0; k < N; k++) {
        int nrows, vo, col; double mt;
        config_vo_inf(b, s, S12, k, zero, rows, &nrows, &vo, &mt, &col, &d0, &d1);
        write_obs(&b->cfg, S12 + 12 * k, rows, nrows, obs + (size_t)(base + k) * L, count + base + k, nbr + (size_t)(base + k) * nm);
    }
    free(S12); free(rows);
}

void uo_step(UoBatch *b, const double *action, double *obs, int32_t *count, int32_t *nbr, double *reward,
             uint8_t *done, uint8_t *info, uint8_t *finish) {
    int64_t c_rng = 0, c_cone = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads()) reduction(+ : c_rng, c_cone)
    for (int e = 0; e < b->E; e++) step_env(b, e, action, obs, count, nbr, reward, done, info, finish, &c_rng, &c_cone);
    b->cnt_in_range = c_rng; b->cnt_cone = c_cone;
}

void uo_observe(UoBatch *b, const uint8_t *env_mask, double *obs, int32_t *count, int32_t *nbr) {
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int e = 0; e < b->E; e++) if (!env_mask || env_mask[e]) observe_env(b, e, obs, count, nbr);
}

void uo_reset(UoBatch *b, const uint8_t *mask) {
    int A = b->E * b->N;
    for (int a = 0; a < A; a++) {
        if (mask && !mask[a]) continue;
        int e = a / b->N, k = a % b->N, s = b->env2scn[e];
        const double *st = WP(b, s, k, 0);
        b->pos[3 * a] = st[0]; b->pos[3 * a + 1] = st[1]; b->pos[3 * a + 2] = st[2];
        b->vel[3 * a] = b->vel[3 * a + 1] = b->vel[3 * a + 2] = 0.0;
        b->yaw[a] = b->pitch[a] = 0.0; b->wp_idx[a] = 1; b->arrive[a] = b->dest[a] = 0;
        b->real_len[a] = 0.0; b->max_dev[a] = 0.0; /* extra_len is NOT cleared (drone.py:270-291) */
    }
}

void uo_autoreset(UoBatch *b, const uint8_t *done, const uint8_t *finish, int32_t *ep_len, int32_t max_ep_len,
                  int32_t epoch_ended, uint8_t *reset_mask, uint8_t *env_observe, uint8_t *path_cut) {
    int N = b->N;
    for (int e = 0; e < b->E; e++) {
        int base = e * N, all_fin = 1, any_fin = 0, any_done = 0, max_len = 0;
        for (int k = 0; k < N; k++) {
            int a = base + k;
            ep_len[a] += 1; reset_mask[a] = 0; path_cut[a] = 0;
            all_fin &= finish[a] != 0; any_fin |= finish[a] != 0; any_done |= done[a] != 0;
            if (ep_len[a] > max_len) max_len = ep_len[a];
        }
        int terminal = any_fin || max_len > max_ep_len;
        env_observe[e] = 0;
        if (any_done) { /* multi_ppo.py:230-242 */
            for (int k = 0; k < N; k++) if (done[base + k]) { reset_mask[base + k] = 1; ep_len[base + k] = 0; }
            env_observe[e] = 1;
        }
        if (epoch_ended || all_fin) { /* :244-264 */
            for (int k = 0; k < N; k++) { reset_mask[base + k] = 1; ep_len[base + k] = 0; path_cut[base + k] = 1; }
            env_observe[e] = 1;
        } else if (terminal) { /* :266-281 */
            for (int k = 0; k < N; k++) {
                int a = base + k;
                if (finish[a] || ep_len[a] > max_ep_len) { reset_mask[a] = 1; ep_len[a] = 0; }
                path_cut[a] = 1;
            }
            env_observe[e] = 1;
        }
    }
    uo_reset(b, reset_mask);
}

UoBatch *uo_create(const UoConfig *cfg, int E, int N, int W, int M, int S, const double *waypoints, const int32_t *n_points,
                   const double *radius, const double *priority, const double *buildings, const int32_t *n_buildings,
                   const double *map_size, const int32_t *env2scn) {
    UoBatch *b = (UoBatch *)calloc(1, sizeof(UoBatch));
    b->cfg = *cfg; b->E = E; b->N = N; b->W = W; b->M = M; b->S = S;
    size_t A = (size_t)E * N, SN = (size_t)S * N;
#define DUP(dst, src, n, T) do { dst = (T *)malloc(sizeof(T) * ((n) > 0 ? (n) : 1)); if (src) memcpy(dst, src, sizeof(T) * (n)); } while (0)
    DUP(b->wp, waypoints, SN * W * 3, double); DUP(b->npts, n_points, SN, int32_t);
    DUP(b->radius, radius, SN, double); DUP(b->prio, priority, SN, double);
    DUP(b->bld, buildings, (size_t)S * M * 4, double); DUP(b->nbld, n_buildings, (size_t)S, int32_t);
    DUP(b->map, map_size, (size_t)S * 3, double); DUP(b->env2scn, env2scn, (size_t)E, int32_t);
#undef DUP
    b->route_len = (double *)malloc(sizeof(double) * SN);
    for (size_t q = 0; q < SN; q++) { /* drone.calculate_total_length (drone.py:409-429) */
        double tot = 0.0;
        const double *w = b->wp + q * W * 3;
        for (int i = 0; i + 1 < b->npts[q]; i++) {
            double dx = w[3 * (i + 1)] - w[3 * i], dy = w[3 * (i + 1) + 1] - w[3 * i + 1], dz = w[3 * (i + 1) + 2] - w[3 * i + 2];
            tot += sqrt(SQ(dx) + SQ(dy) + SQ(dz));
        }
        b->route_len[q] = tot;
    }
    b->pos = (double *)calloc(A * 3, sizeof(double)); b->vel = (double *)calloc(A * 3, sizeof(double));
    b->yaw = (double *)calloc(A, sizeof(double)); b->pitch = (double *)calloc(A, sizeof(double));
    b->real_len = (double *)calloc(A, sizeof(double)); b->max_dev = (double *)calloc(A, sizeof(double));
    b->extra_len = (double *)calloc(A, sizeof(double)); b->wp_idx = (int32_t *)calloc(A, sizeof(int32_t));
    b->arrive = (uint8_t *)calloc(A, 1); b->dest = (uint8_t *)calloc(A, 1);
    uo_reset(b, NULL);
    return b;
}

void uo_destroy(UoBatch *b) {
    if (!b) return;
    free(b->wp); free(b->npts); free(b->route_len); free(b->radius); free(b->prio); free(b->bld); free(b->nbld);
    free(b->map); free(b->env2scn); free(b->pos); free(b->vel); free(b->yaw); free(b->pitch); free(b->real_len);
    free(b->max_dev); free(b->extra_len); free(b->wp_idx); free(b->arrive); free(b->dest); free(b);
}

#define CPY(dst, src, n, T) do { if (dst) memcpy(dst, src, sizeof(T) * (n)); } while (0)
void uo_get_state(const UoBatch *b, double *pos, double *vel, double *yaw, double *pitch, int32_t *wp_idx, uint8_t *arrive,
                  uint8_t *dest, double *real_len, double *max_dev, double *extra_len) {
    size_t A = (size_t)b->E * b->N;
    CPY(pos, b->pos, A * 3, double); CPY(vel, b->vel, A * 3, double); CPY(yaw, b->yaw, A, double);
    CPY(pitch, b->pitch, A, double); CPY(wp_idx, b->wp_idx, A, int32_t); CPY(arrive, b->arrive, A, uint8_t);
    CPY(dest, b->dest, A, uint8_t); CPY(real_len, b->real_len, A, double); CPY(max_dev, b->max_dev, A, double);
    CPY(extra_len, b->extra_len, A, double);
}
void uo_set_state(UoBatch *b, const double *pos, const double *vel, const double *yaw, const double *pitch, const int32_t *wp_idx,
                  const uint8_t *arrive, const uint8_t *dest, const double *real_len, const double *max_dev, const double *extra_len) {
    size_t A = (size_t)b->E * b->N;
    if (pos) memcpy(b->pos, pos, sizeof(double) * A * 3);
    if (vel) memcpy(b->vel, vel, sizeof(double) * A * 3);
    if (yaw) memcpy(b->yaw, yaw, sizeof(double) * A);
    if (pitch) memcpy(b->pitch, pitch, sizeof(double) * A);
    if (wp_idx) memcpy(b->wp_idx, wp_idx, sizeof(int32_t) * A);
    if (arrive) memcpy(b->arrive, arrive, A);
    if (dest) memcpy(b->dest, dest, A);
    if (real_len) memcpy(b->real_len, real_len, sizeof(double) * A);
    if (max_dev) memcpy(b->max_dev, max_dev, sizeof(double) * A);
    if (extra_len) memcpy(b->extra_len, extra_len, sizeof(double) * A);
}
void uo_cal_des(const UoBatch *b, double *des) {
    int A = b->E * b->N;
    for (int a = 0; a < A; a++) {
        int e = a / b->N, k = a % b->N, s = b->env2scn[e];
        cal_des_vel(b, b->pos + 3 * (size_t)a, WP(b, s, k, b->wp_idx[a]), des + 3 * (size_t)a);
    }
}
void uo_get_route_len(const UoBatch *b, double *route_len) { memcpy(route_len, b->route_len, sizeof(double) * (size_t)b->S * b->N); }
void uo_get_counters(const UoBatch *b, int64_t *r, int64_t *c) { *r = b->cnt_in_range; *c = b->cnt_cone; }

/* =====================================================================================================================
 * K7 -- classic RVO velocity selection: reciprocal_vel_obs.cal_vel -> vel_candidate -> vo_out2 -> vel_select -> penalty
 * (uaisa_env/vel_obs/reciprocal_vel_obs.py:21-166).
 *
 * PARITY UNPINNED.  The reference's code for this path is dead (no caller, grep cal_vel) and cannot run: it raises on the
 * first call.  This is therefore a CORRECTED restatement -- it follows the reference statement by statement and fixes only
 * what stops it from running or makes it meaningless; every fix is listed here and in DESIGN.md:
 *   F1  config_vo is defined twice (:56, :63) and cal_vel calls it with the list (:26): the per-neighbour form (:63-83) is
 *       mapped over the neighbour list.
 *   F2  Pa = state[0:4] / Pb = odro[0:4] (:65, :70) feed 4 values into 3-value unpackings (get_PAA): positions are [0:3].
 *   F3  get_alpha calls math.asin(R / |rel|), a ValueError for overlapping drones: alpha = round(pi/2, 2) there.
 *   F4  vo_out2: `arr_AA_Anew.append = ...` (:122) -> append(...);  `odro_rvo[4:6]` (:118, a 2-vector) -> the relative
 *       position odro_rvo[3:6].
 *   F5  vel_select returns None when vo_outside is non-empty (:134-138): the selected velocity is returned in both cases.
 *   F6  penalty (:140-166) evaluates cal_exp_tim with the agent's CURRENT velocity -- a constant that cannot influence the
 *       arg-min; the candidate velocity is used instead (what the `vel` argument is for).  tc_min = inf gives 1/inf = 0.
 *   F7  no candidate survives the 0.3 speed floor (:98) -> min() of an empty list raises: the current velocity is kept.
 *   F8  "self" is not in the list cal_vel receives (the caller passes the OTHER drones): excluded here by index.
 *   F9  np.arange(lo, hi, 0.5) (:93-96) excludes hi: with acceler = 0.5 the grid would be {v-0.5, v} per axis (never v+0.5), so a drone at
 *       rest could only pick negative components and -- with the 0.3 speed floor -- never its desired direction.  The grid INCLUDES the
 *       end point: lo + 0.5 i for every i with lo + 0.5 i <= hi + 1e-9 ("within +-acceler of v").
 * Kept as they are: the hard-coded neighbour range 10 (:43), buildings filtered but never used (:47-54, config_vo only sees
 * drones), the 0.5 grid step and the 0.3 speed floor, builtin round for alpha / numpy-style round for beta, the unclamped arccos, first
 * minimum wins (Python min; candidates in vx-outer, vz-inner order), ir_gym's constructor values vmax = 2, acceler = 0.5, delta_t = 1 (ir_gym.py:18-29).
 * Arithmetic of the distances is plain IEEE mul/add/sqrt (x*x, not pow): this restatement defines it.
 * ===================================================================================================================== */
static double k7_exp_tim(const double *Pa, const double *Pb, const double *Va, const double *Vb, double ra, double rb) { /* vel_obs3D.cal_exp_tim :104-143 */
    double rx = Pa[0] - Pb[0], ry = Pa[1] - Pb[1], rz = Pa[2] - Pb[2];
    double vx = Va[0] - Vb[0], vy = Va[1] - Vb[1], vz = Va[2] - Vb[2];
    double r = ra + rb;
    double a = vx * vx + vy * vy + vz * vz;
    double b = 2 * rx * vx + 2 * ry * vy + 2 * rz * vz;
    double c = rx * rx + ry * ry + rz * rz - r * r;
    if (c <= 0) return 0;
    double temp = b * b - 4 * a * c;
    if (temp <= 0) return INFINITY;
    double sq = sqrt(temp);
    double t1 = (-b + sq) / (2 * a), t2 = (-b - sq) / (2 * a);
    double t3 = t1 >= 0 ? t1 : INFINITY, t4 = t2 >= 0 ? t2 : INFINITY;
    return t3 < t4 ? t3 : t4;
}
static double k7_dist(const double *p, const double *q) { /* reciprocal_vel_obs.distance :169-171 */
    double dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}
#define K7_MAXC 5 /* candidates per axis: 2*acceler/0.5 + 1 for acceler <= 1 */

/* ctrl_out (may be NULL): the selected velocity expressed in the env's own controls, because mdin.drone_step re-reads its "action" as
 * [acceleration, yaw change / 90 deg, pitch change / 90 deg] (drone.py:431-490): acc = |v*| - |v|, yaw / pitch steered onto v*'s heading, each clipped
 * to [-1, 1].  Not in the reference (its only usage example, uaisa_env/gym_env_test.py:12, feeds the velocity itself and spirals): this is the
 * adapter that makes the selection usable as a scripted baseline controller (uaisa_ppo.post_train, policy "rvo"). */
void uo_rvo_select(const UoBatch *b, double vmax, double acceler, double *vel_out /*[A][3]*/, double *ctrl_out /*[A][3] or NULL*/) {
    const int N = b->N;
    const double dt = b->cfg.delta_t;
#pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int e = 0; e < b->E; e++) {
        const int s = b->env2scn[e];
        for (int k = 0; k < N; k++) {
            const int a = e * N + k;
            const double *Pa = b->pos + 3 * (size_t)a, *Va = b->vel + 3 * (size_t)a;
            const double ra = b->radius[(size_t)s * N + k], pra = b->prio[(size_t)s * N + k];
            double des[3];
            cal_des_vel(b, Pa, WP(b, s, k, b->wp_idx[a]), des);                 /* agent_state[8:11] */
            /* vel_candidate :85-96: axis grids */
            double grid[3][K7_MAXC]; int ng[3];
            for (int ax = 0; ax < 3; ax++) {
                double lo = np_clip(Va[ax] - acceler, -vmax, vmax), hi = np_clip(Va[ax] + acceler, -vmax, vmax);
                int n = 0;                                                       /* F9: end point included */
                while (n < K7_MAXC && lo + n * 0.5 <= hi + 1e-9) { grid[ax][n] = lo + n * 0.5; n++; }
                ng[ax] = n;
            }
            int have_out = 0, have_in = 0;
            double best_out = 0, best_in = 0, vout[3] = {0, 0, 0}, vin[3] = {0, 0, 0};
            for (int ix = 0; ix < ng[0]; ix++) for (int iy = 0; iy < ng[1]; iy++) for (int iz = 0; iz < ng[2]; iz++) {
                const double v[3] = {grid[0][ix], grid[1][iy], grid[2][iz]};
                if (sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) < 0.3) continue;       /* :98 */
                const double Pn[3] = {Pa[0] + v[0] * dt, Pa[1] + v[1] * dt, Pa[2] + v[2] * dt};   /* vo_out2 :111-114 */
                int inside = 0; double tc_min = INFINITY;
                for (int j = 0; j < N; j++) {
                    if (j == k) continue;                                                    /* F8 */
                    const int o = e * N + j;
                    const double *Pb = b->pos + 3 * (size_t)o, *Vb = b->vel + 3 * (size_t)o;
                    const double dif[3] = {Pa[0] - Pb[0], Pa[1] - Pb[1], Pa[2] - Pb[2]};
                    if (!(np_norm3(dif) <= 10.0)) continue;                                  /* preprocess :38-44 */
                    const double rb = b->radius[(size_t)s * N + j], prb = b->prio[(size_t)s * N + j];
                    const double rel[3] = {Pb[0] - Pa[0], Pb[1] - Pa[1], Pb[2] - Pa[2]};       /* get_rvo_array */
                    const double nab = np_norm3(rel), R = ra + rb;
                    const double alpha = (R / nab <= 1.0) ? uo_pyround2(wraptopi(asin(R / nab))) : uo_pyround2(M_PI / 2);   /* get_alpha, F3 */
                    const double pr = pra / (pra + prb);
                    const double PAA[3] = {pr * (2 * Pa[0] + (Va[0] + Vb[0]) * 1), pr * (2 * Pa[1] + (Va[1] + Vb[1]) * 1), pr * (2 * Pa[2] + (Va[2] + Vb[2]) * 1)};
                    const double B[3] = {Pn[0] - PAA[0], Pn[1] - PAA[1], Pn[2] - PAA[2]};
                    const double AB = nab * np_norm3(B);                                     /* get_beta :44-66 */
                    const double cosang = (AB != 0) ? np_dot3(rel, B) / AB : 0.0;
                    const double beta = uo_npround(wraptopi(acos(cosang)), 2);
                    if (alpha > beta) inside = 1;                                            /* :125-126 (NaN beta -> outside) */
                    const double tc = k7_exp_tim(Pa, Pb, v, Vb, ra, rb);                     /* penalty :151-153, F6 */
                    if (tc < tc_min) tc_min = tc;
                }
                if (!inside) {                                                               /* vel_select :134-135 */
                    const double dd = k7_dist(v, des);
                    if (!have_out || dd < best_out) { have_out = 1; best_out = dd; vout[0] = v[0]; vout[1] = v[1]; vout[2] = v[2]; }
                } else {                                                                     /* :137, penalty :140-166 with factor 1 */
                    const double tc_inv = (tc_min == 0) ? INFINITY : 1.0 / tc_min;
                    const double pen = 1.0 * tc_inv + k7_dist(des, v);
                    if (!have_in || pen < best_in) { have_in = 1; best_in = pen; vin[0] = v[0]; vin[1] = v[1]; vin[2] = v[2]; }
                }
            }
            double *out = vel_out + 3 * (size_t)a;
            if (have_out) { out[0] = vout[0]; out[1] = vout[1]; out[2] = vout[2]; }
            else if (have_in) { out[0] = vin[0]; out[1] = vin[1]; out[2] = vin[2]; }
            else { out[0] = Va[0]; out[1] = Va[1]; out[2] = Va[2]; }                         /* F7 */
            if (ctrl_out) {
                double *c = ctrl_out + 3 * (size_t)a;
                const double sp_new = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]), sp = sqrt(Va[0] * Va[0] + Va[1] * Va[1] + Va[2] * Va[2]);
                double dyaw = 0, dpitch = 0;
                if (sp_new > 0) {
                    const double yaw_t = atan2(out[1], out[0]) * (180.0 / M_PI), pitch_t = atan2(out[2], sqrt(out[0] * out[0] + out[1] * out[1])) * (180.0 / M_PI);
                    dyaw = py_mod(yaw_t - b->yaw[a] + 180.0, 360.0) - 180.0;
                    dpitch = pitch_t - b->pitch[a];
                }
                c[0] = np_clip(sp_new - sp, -1.0, 1.0); c[1] = np_clip(dyaw / 90.0, -1.0, 1.0); c[2] = np_clip(dpitch / 90.0, -1.0, 1.0);
            }
        }
    }
}

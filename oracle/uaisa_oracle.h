/*
 * uaisa_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference env step).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The product path (libuaisa_b200.so) never does.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_reference.py, run in
 * the build container where /root/reference exists) against the unmodified reference imported
 * through oracle/ref_shim.py, and (everywhere) against the golden vectors under tests/golden/
 * generated from that same reference by oracle/gen_golden.py.
 */
#ifndef UAISA_ORACLE_H
#define UAISA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hyper-parameters that the reference hard-codes in constructors (SURVEY.md section 5 "Config"). */
typedef struct UoConfig {
    double goal_threshold;   /* drone.py:15   0.4  */
    double max_acc;          /* drone.py:72   1.0  */
    double max_angle_change; /* drone.py:73   90.0 */
    double dt;               /* drone.py:14   1    */
    double vel_max;          /* drone.py:15   1.0 (all three axes) */
    double neighbor_range;   /* rvo_inter.py:96   10 */
    double building_range;   /* rvo_inter.py:104  5  */
    double building_zgate;   /* rvo_inter.py:101  2  */
    double ctime_threshold;  /* rvo_inter.py:11   2  */
    double delta_t;          /* rvo_inter.py:11   1  */
    double exp_radius;       /* rvo_inter.py:11   0.2 */
    int32_t env_train;       /* ir_gym.py:18  True */
    int32_t nm;              /* mdin.py:7     10 */
} UoConfig;

typedef struct UoBatch UoBatch;

void uo_default_config(UoConfig *c);

/* E envs x N drones; S scenarios (W = max waypoints, M = max buildings), env2scn[E] in [0,S). */
UoBatch *uo_create(const UoConfig *cfg, int E, int N, int W, int M, int S,
                   const double *waypoints /*[S][N][W][3]*/, const int32_t *n_points /*[S][N]*/,
                   const double *radius /*[S][N]*/, const double *priority /*[S][N]*/,
                   const double *buildings /*[S][M][4] x,y,h,r*/, const int32_t *n_buildings /*[S]*/,
                   const double *map_size /*[S][3]*/, const int32_t *env2scn /*[E]*/);
void uo_destroy(UoBatch *b);
void uo_set_threads(int n); /* OpenMP threads used by the batched entry points (0 = all) */
int  uo_get_max_threads(void);

/* drone.reset (drone.py:270-291) for agents with mask!=0 (NULL = all). No observation. */
void uo_reset(UoBatch *b, const uint8_t *agent_mask);
/* ir_gym.env_observation (ir_gym.py:372-383) for envs with env_mask!=0 (NULL = all):
 * obs [A][12+9nm] zero padded, count[A], nbr[A][nm] (-1 = empty). Rows of unselected envs untouched. */
void uo_observe(UoBatch *b, const uint8_t *env_mask, double *obs, int32_t *count, int32_t *nbr);
/* mdin.drone_step (mdin.py:19-30) for every env. */
void uo_step(UoBatch *b, const double *action /*[A][3]*/, double *obs, int32_t *count, int32_t *nbr,
             double *reward, uint8_t *done, uint8_t *info, uint8_t *finish);
/* trainer-side reset policy, multi_ppo.py:226-281 (per env): returns per-agent reset mask and per-env
 * "observe" flag and applies the resets; ep_len[A] is updated in place. path_cut[A]=1 where finish_path
 * would be called. */
void uo_autoreset(UoBatch *b, const uint8_t *done, const uint8_t *finish, int32_t *ep_len,
                  int32_t max_ep_len, int32_t epoch_ended, uint8_t *reset_mask, uint8_t *env_observe,
                  uint8_t *path_cut);

/* SoA state access (doubles). Any pointer may be NULL. */
void uo_get_state(const UoBatch *b, double *pos, double *vel, double *yaw, double *pitch,
                  int32_t *wp_idx, uint8_t *arrive, uint8_t *dest, double *real_len, double *max_dev,
                  double *extra_len);
void uo_set_state(UoBatch *b, const double *pos, const double *vel, const double *yaw,
                  const double *pitch, const int32_t *wp_idx, const uint8_t *arrive,
                  const uint8_t *dest, const double *real_len, const double *max_dev,
                  const double *extra_len);
/* drone.cal_des_vel for all agents (ir_gym.cal_des_list, ir_gym.py:44-46) -> [A][3] */
void uo_cal_des(const UoBatch *b, double *des);
/* route_len[S][N] as computed by drone.calculate_total_length (drone.py:409-429) */
void uo_get_route_len(const UoBatch *b, double *route_len);
/* counters for the in-cone evaluations of the last uo_step (roofline P_cone) */
void uo_get_counters(const UoBatch *b, int64_t *in_range_pairs, int64_t *cone_evals);

/* single pair function, exposed for known-answer tests: rvo_inter.config_vo_circle2
 * (rvo_inter.py:116-196). out9 = row, returns bit0 vo_flag, bit1 collision; t, md via pointers. */
int uo_pair(const UoConfig *cfg, const double *sa8, const double *sb8, const double *act, double *out9,
            double *t, double *md);
/* K7, classic RVO velocity selection (reciprocal_vel_obs.py:21-166): PARITY UNPINNED -- the reference's code is dead and raises on its
 * first call; this is the corrected restatement documented at the function (uaisa_oracle.c).  vel_out[A][3]. */
void uo_rvo_select(const UoBatch *b, double vmax, double acceler, double *vel_out, double *ctrl_out /* [A][3] env controls, may be NULL */);
double uo_pyround2(double x);          /* builtin round(x, 2) on a Python float */
double uo_npround(double x, int nd);   /* np.round(x, nd) */

#ifdef __cplusplus
}
#endif
#endif

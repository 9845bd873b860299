/*
 * uaisa_b200.h -- C ABI of libuaisa_b200.so: the batched 3-D RVO multi-drone environment step on
 * NVIDIA B200 (sm_100a).  This is the drop-in boundary for the hot path of
 * ZSHCRWY25/3DRVO-MARL-CollisionAvoidance.
 *
 * The reference has no FFI: its boundary is the duck-typed Python object `mdin`
 * (uaisa_env/drone_envs/mdin.py:6-49) that train/policy/multi_ppo.py drives.  Each entry point below
 * names the reference method it replaces; the Python side that binds these symbols with ctypes is
 * `3drvo-marl-collisionavoidance_b200/uaisa_env/_capi.py`.
 *
 * Conventions
 *   - every function returns 0 on success or a negative UAISA_E* code; nothing throws or aborts
 *     across the boundary; uaisa_last_error() returns a thread-local message for the last failure.
 *   - the handle is opaque and library-owned, bound to the CUDA device current at uaisa_create().
 *   - unless a parameter is marked HOST, I/O buffers are caller-owned DEVICE pointers (e.g. torch CUDA
 *     tensors); the library never frees or retains them past the call.
 *   - calls are asynchronous on the caller's `stream` (a cudaStream_t passed as void*, NULL = legacy
 *     default stream); no hidden synchronisation or allocation in step/observe/reset/autoreset.
 *   - a handle is not re-entrant; independent handles may be used from different threads
 *     (the reference runs a train env and a test env concurrently: multi_ppo.py:291-292).
 *   - agents are numbered a = env * N + drone;  A = E * N.
 *   - observation layout per agent (ir_gym.py:203-229): 12 self values [p(3) v(3) r prio des_vel(3) dev]
 *     then nm rows of 9 [apex(3) rel_pos(3) alpha min_dis 1/(t+0.2)], most urgent LAST, zero padded;
 *     `count` = number of valid rows; `nbr` = in-env drone index of each row (-1 = empty).
 */
#ifndef UAISA_B200_H
#define UAISA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAISA_OK 0
#define UAISA_EINVAL (-1)  /* bad argument / configuration */
#define UAISA_ECUDA (-2)   /* a CUDA runtime call failed (message has the cudaError string) */
#define UAISA_ENOMEM (-3)
#define UAISA_ESTATE (-4)  /* call order: scenarios not set */

#define UAISA_OBS_SELF 12
#define UAISA_ROW 9
#define UAISA_MAX_NM 16
#define UAISA_MAX_DRONES 1024

/* Hyper-parameters the reference hard-codes in constructors (citations: SURVEY.md section 5). */
typedef struct UaisaConfig {
    int32_t n_envs;          /* E */
    int32_t n_drones;        /* N  (data_1.json drone_num)            <= UAISA_MAX_DRONES */
    int32_t max_waypoints;   /* W                                      >= 2 */
    int32_t max_buildings;   /* M                                      >= 1 */
    int32_t n_scenarios;     /* S  distinct scenarios shared by the envs */
    int32_t nm;              /* neighbour slots, mdin.py:7 neighbors_num=10   <= UAISA_MAX_NM */
    int32_t env_train;       /* ir_gym.py:18 env_train=True */
    int32_t threads_per_block; /* 0 = auto */
    int32_t count_pairs;     /* 1 = maintain the pair counters of uaisa_get_counters (costs atomics) */
    int32_t neighbor_slices; /* 0 = auto; s > 1: s threads per drone share its neighbour loop (N > 32 only; s * threads_per_block <= 128) */
    double goal_threshold;   /* drone.py:15      0.4 */
    double max_acc;          /* drone.py:72      1.0 */
    double max_angle_change; /* drone.py:73      90  */
    double dt;               /* drone.py:14      1   */
    double vel_max;          /* drone.py:15      1   */
    double neighbor_range;   /* rvo_inter.py:96  10  */
    double building_range;   /* rvo_inter.py:104 5   */
    double building_zgate;   /* rvo_inter.py:101 2   */
    double ctime_threshold;  /* rvo_inter.py:11  2   */
    double delta_t;          /* rvo_inter.py:11  1   */
    double exp_radius;       /* rvo_inter.py:11  0.2 */
} UaisaConfig;

/* Mutable per-agent state, structure-of-arrays, fp32 (drone.py:19-76).  Used by get/set_state:
 * each non-NULL pointer is a caller buffer (device or host, any cudaMemcpyDefault-able address). */
typedef struct UaisaStateView {
    float *pos_yaw;     /* [A][4]  x, y, z, yaw (deg)     drone.state, drone.yaw   */
    float *vel_pitch;   /* [A][4]  vx, vy, vz, pitch (deg) drone.vel, drone.pitch   */
    float *real_len;    /* [A]     drone.real_route_len */
    float *max_dev;     /* [A]     drone.max_deviation  */
    float *extra_len;   /* [A]     drone.extra_len      */
    int32_t *wp_flags;  /* [A]     bits 0-15 drone.i, bit 16 arrive_flag, bit 17 destination_arrive_flag */
} UaisaStateView;

typedef struct UaisaHandle UaisaHandle;

const char *uaisa_last_error(void);
int uaisa_version(void);
int uaisa_default_config(UaisaConfig *cfg);

/* mdin.__init__ -> ir_gym.__init__ -> env_base.__init__ (mdin.py:7-16, env_base.py:15-24) */
int uaisa_create(const UaisaConfig *cfg, UaisaHandle **out);
int uaisa_destroy(UaisaHandle *h);

/* env_base.load_data (env_base.py:26-47): scenario constants, HOST pointers, copied to the device.
 * waypoints[S][N][W][3], n_points[S][N], radius[S][N], priority[S][N], buildings[S][M][4]=(x,y,h,r),
 * n_buildings[S], map_size[S][3], env2scenario[E].  Doubles: a json scenario is kept exactly.
 * Also resets every agent (drone.__init__ state).  Synchronous. */
int uaisa_set_scenarios(UaisaHandle *h, const double *waypoints, const int32_t *n_points, const double *radius,
                        const double *priority, const double *buildings, const int32_t *n_buildings,
                        const double *map_size, const int32_t *env2scenario);

/* mdin.drone_step (mdin.py:19-30): action[A][3] -> obs[A][12+9nm], count[A], nbr[A][nm], reward[A],
 * done[A] (collision), info[A] (arrive_flag), finish[A] (destination_arrive_flag).
 * Any output pointer may be NULL (not written). */
int uaisa_step(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward,
               uint8_t *done, uint8_t *info, uint8_t *finish, void *stream);

/* mdin.drone_reset / drone_reset_one (mdin.py:38-46) + ir_gym.env_observation (ir_gym.py:372-383):
 * resets agents with agent_mask[a]!=0 (NULL = all agents) as drone.reset (drone.py:270-291), then, when
 * obs != NULL, writes the observations of every agent of each env that had a reset. */
int uaisa_reset(UaisaHandle *h, const uint8_t *agent_mask, float *obs, uint8_t *count, int32_t *nbr, void *stream);

/* ir_gym.env_observation (ir_gym.py:372-383) for envs with env_mask[e]!=0 (NULL = all). */
int uaisa_observe(UaisaHandle *h, const uint8_t *env_mask, float *obs, uint8_t *count, int32_t *nbr, void *stream);

/* Trainer-side reset policy of multi_ppo.training_loop (multi_ppo.py:221-281), per env, on device:
 * ep_len += 1; done_i -> reset i; all finish or epoch_ended -> reset env; any finish or
 * ep_len > max_ep_len -> reset those; then env_observation for every touched env (obs/count/nbr are
 * overwritten for those envs only).  reset_mask[A] / path_cut[A] (finish_path would be called) may be NULL. */
int uaisa_autoreset(UaisaHandle *h, const uint8_t *done, const uint8_t *finish, int32_t max_ep_len, int32_t epoch_ended,
                    float *obs, uint8_t *count, int32_t *nbr, uint8_t *reset_mask, uint8_t *path_cut, void *stream);

/* T open-loop steps in ONE launch (state stays on chip between steps): action[T][A][3]; outputs are
 * [T][...] with the same per-step layouts as uaisa_step (NULL = skip).  auto_reset!=0 applies the
 * uaisa_autoreset policy after every step (max_ep_len as given, epoch_ended = 0). */
int uaisa_rollout(UaisaHandle *h, int32_t T, const float *action, float *obs, uint8_t *count, int32_t *nbr,
                  float *reward, uint8_t *done, uint8_t *info, uint8_t *finish, int32_t auto_reset,
                  int32_t max_ep_len, void *stream);

/* uaisa_step followed by uaisa_autoreset in ONE launch (same outputs, bit for bit): the env part of one iteration of multi_ppo.training_loop
 * (multi_ppo.py:210-281).  reset_mask / path_cut [A] may be NULL. */
int uaisa_step_autoreset(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward, uint8_t *done,
                         uint8_t *info, uint8_t *finish, int32_t max_ep_len, int32_t epoch_ended, uint8_t *reset_mask, uint8_t *path_cut, void *stream);

/* state access (teacher forcing, drone_list[i].vel): copies between the handle and caller buffers */
int uaisa_get_state(UaisaHandle *h, const UaisaStateView *dst, void *stream);
int uaisa_set_state(UaisaHandle *h, const UaisaStateView *src, void *stream);
/* ir_gym.cal_des_list (ir_gym.py:44-46): des[A][3] */
int uaisa_cal_des(UaisaHandle *h, float *des, void *stream);
/* K7 -- classic RVO velocity selection, reciprocal_vel_obs.cal_vel (uaisa_env/vel_obs/reciprocal_vel_obs.py:21-31) with vel_candidate :85-107, vo_out2
 * :109-128, vel_select :131-138, penalty :140-166, on the handle's current state: vel_out[A][3] = the selected velocity of every drone (grid of 0.5
 * within +-acceler of its velocity, clipped to +-vmax; ir_gym.py:18-29 constructs the class with vmax = 2, acceler = 0.5).  PARITY UNPINNED: the
 * reference's code for this is dead and raises; the behaviour is the corrected restatement oracle/uaisa_oracle.c:uo_rvo_select (fixes F1-F9 listed
 * there).  ctrl_out[A][3] (may be NULL) = the same velocity expressed in the controls mdin.drone_step actually applies
 * ([acc, yaw change / 90, pitch change / 90], drone.py:431-490), i.e. what to feed uaisa_step to fly it. */
int uaisa_rvo_select(UaisaHandle *h, double vmax, double acceler, float *vel_out, float *ctrl_out, void *stream);
/* Trainer action transform (multi_ppo.py:197-205): a_inc[A][3] = the policy's float32 sample;
 * a_inc_rounded = np.round(a_inc, 2) (what the buffer stores, may be NULL); action = np.round(acceler * a_inc_rounded +
 * drone.vel, 2) = what uaisa_step consumes. */
int uaisa_action_transform(UaisaHandle *h, const float *a_inc, float acceler, float *a_inc_rounded, float *action, void *stream);
/* GAE-lambda with last_val = 0 over [T][A] buffers (multi_PPObuf.finish_path, multi_ppo.py:68-77): a path ends after
 * step t where cut[t][a] != 0 (uaisa_autoreset's path_cut) and at t = T-1.  adv/ret [T][A]. */
int uaisa_gae(const float *rew, const float *val, const uint8_t *cut, int32_t T, int32_t A, float gamma, float lam,
              float *adv, float *ret, void *stream);
/* ir_gym.indicators_deviation / indicators_extra_len (ir_gym.py:414-420): max_dev[A], extra_len[A] */
int uaisa_indicators(UaisaHandle *h, float *max_dev, float *extra_len, void *stream);
/* HOST outputs; synchronises `stream`.  in_range_pairs / cone_evals: pair evaluations and full cone
 * evaluations since the last call (only with cfg.count_pairs); nonfinite_obs: observations that
 * contained NaN/Inf (the reference raises ValueError there, ir_gym.py:232-239). */
int uaisa_get_counters(UaisaHandle *h, int64_t *in_range_pairs, int64_t *cone_evals, int64_t *nonfinite_obs, void *stream);
/* kernel launches issued through this handle so far (bench.py's gpu_launches) */
int64_t uaisa_launch_count(const UaisaHandle *h);
/* Profiling aid: per-region SM-cycle sums (summed over warps) since the last call; all zero unless the
 * library was built with -DUAISA_REGION_CLOCKS.  out16 is HOST; synchronises `stream`. */
int uaisa_get_region_clocks(UaisaHandle *h, int64_t *out16, void *stream);

/* Host-buffer convenience used for the end-to-end number: H2D(action) -> step [-> trainer auto-reset
 * + re-observe, as uaisa_rollout's auto_reset] -> D2H(outputs); all pointers HOST (pinned for full
 * speed); synchronous on return. */
int uaisa_step_host(UaisaHandle *h, const float *action, float *obs, uint8_t *count, int32_t *nbr, float *reward,
                    uint8_t *done, uint8_t *info, uint8_t *finish, int32_t auto_reset, int32_t max_ep_len);

/* ---- fused rollout forward of the actor-critic (train/policy/policy_rnn_ac.py:57-69, SURVEY 8a row a-P1) --------------------
 * uaisa_policy_features: obs[A][obs_dim] -> feat[A][288] = LayerNorm(268)(cat(self 12, h_fwd + h_bwd)) zero padded to 288, with the
 *   bi-GRU run for ONE step from h0 = 0 (exact for agents with <= 1 neighbour row; the caller patches the others).
 *   wih_t [2][9][768]: weight_ih_l0 / weight_ih_l0_reverse transposed to [k][3H]; bias [2][4][256]: b_ir+b_hr, b_iz+b_hz, b_in, b_hn.
 * uaisa_policy_heads: feat -> mu[A][3] = tanh(actor MLP), v[A] = critic MLP on the tensor cores (tcgen05.mma kind::tf32, TMEM
 *   accumulators).  weights: 12 DEVICE pointers W1a[256][288] b1a[256] W2a[256][256] b2a[256] W3a[3][256] b3a[3] W1c b1c W2c b2c
 *   W3c[1][256] b3c[1] (W1 rows zero padded from 268 to 288).  Both are stateless and asynchronous on `stream`. */
int uaisa_policy_features(const float *obs, int32_t A, int32_t obs_dim, const float *wih_t, const float *bias, const float *ln_w,
                          const float *ln_b, float *feat, void *stream);
int uaisa_policy_heads(const float *feat, int32_t A, const float *const *weights, float *mu, float *v, void *stream);

/* ---- K6, complete: the whole rollout forward in ONE kernel (csrc/uaisa_policy_fwd.cu) ------------------------------------------------------------
 * uaisa_policy_pack:    repacks the module's parameters (22 DEVICE pointers, torch layouts: weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, the same
 *   four with suffix _reverse, ln.weight, ln.bias, actor net_out.{0,2,4}.{weight,bias}, critic v_net.{0,2,4}.{weight,bias}) into `wpack`
 *   ([uaisa_policy_pack_floats()] floats, 16-byte aligned): every weight tile as the swizzled shared-memory image the tensor cores read, so that the
 *   forward kernel fetches a tile with ONE bulk copy.  Call it after every optimiser step (in place: `wpack` keeps its address).
 * uaisa_policy_forward:  obs[A][obs_dim] + count[A] -> mu[A][3] = tanh(actor MLP), v[A] = critic MLP, with the masked bi-GRU recurrence over each
 *   agent's max(count, 1) neighbour rows (policy_rnn_ac.py:102-168), LayerNorm(268) and both MLPs on the tensor cores (tcgen05.mma kind::tf32,
 *   fp32 accumulation in TMEM).  feat_out (may be NULL): the LayerNorm output [A][268] in the module's [self | h] order.  scratch: DEVICE buffer of
 *   ceil(A/128) * 128 * 1024 bytes (forward-direction GRU state of tiles that run the recurrence).  stats (may be NULL): int32[2] {tiles that ran the
 *   recurrence (+=), max steps (max)}.  Stateless, asynchronous on `stream`, capturable in a CUDA graph. */
int uaisa_policy_pack_floats(void);
int uaisa_policy_pack(const float *const *params, float *wpack, void *stream);
int uaisa_policy_forward(const float *obs, const uint8_t *count, int32_t A, int32_t obs_dim, const float *wpack, float *mu, float *v, float *feat_out,
                         void *scratch, int32_t *stats, void *stream);
/* uaisa_policy_sample: the sampling half of rnn_ac.step (train/policy/policy_rnn_ac.py:57-69; GaussianActor._distribution :206-228): std = clamp(std_factor
 *   exp(log_std[3]) + 1e-6, 1e-4, 10), a_out[A][3] = mu + std eps, logp[A] = sum_q(-(a - mu)^2 / (2 std^2) - log std - log sqrt(2 pi)), in one launch and in
 *   the operation order of the torch expression.  eps[A][3]: standard normal draws from the caller's generator.  All DEVICE pointers. */
int uaisa_policy_sample(const float *mu, const float *eps, const float *log_std, float std_factor, int32_t A, float *a_out, float *logp, void *stream);

/* ---- optimiser epilogue of the PPO update (train/policy/multi_ppo.py:355-376; SURVEY.md section 5 "fused grad-norm + clip + Adam") -----------
 * Parameters, gradients and Adam moments are FLAT fp32 DEVICE buckets.  One update iteration ends with
 *   uaisa_opt_prepare : partial sums of |g|^2 (fp64, fixed partition: bit-identical on every rank) when want_norm != 0, and the early-stop rule
 *                       on the device -- `ctl[0]` latches once *kl > target_kl ("Early stopping ... break", multi_ppo.py:362-364), otherwise
 *                       `step[0]` (the optimiser's step count) and `ctl[1]` (iterations applied) are incremented.  kl / ctl may be NULL
 *                       (value optimiser: no early stop).
 *   uaisa_clip_adam   : clip_grad_norm_(max_norm) over the WHOLE gradient bucket g[0..n_total) in place (multi_ppo.py:367; max_norm <= 0: off),
 *                       then torch.optim.Adam's step (betas, eps as given, no weight decay) on p[r_off .. r_off+r_cnt); nothing is touched
 *                       when ctl[0] is set.
 * partials: double[uaisa_opt_partials_len()].  Asynchronous on `stream`, capturable in a CUDA graph. */
int uaisa_opt_partials_len(void);
int uaisa_opt_prepare(const float *g, int64_t n, int32_t want_norm, double *partials, const float *kl, float target_kl, int32_t *ctl,
                      int32_t *step, void *stream);
int uaisa_clip_adam(float *p, float *g, float *m, float *v, int64_t n_total, int64_t r_off, int64_t r_cnt, const double *partials,
                    float max_norm, float lr, float beta1, float beta2, float eps, const int32_t *step, const int32_t *ctl, void *stream);

/* Reader forward / backward for update minibatches whose samples all have <= 1 neighbour row (rnn_Reader.obs_rnn_list with L = 1, policy_rnn_ac.py:129-168:
 * the bi-GRU from h0 = 0 -- W_hh never enters --, the sum of both directions, the concat with the 12 self values, LayerNorm(268)) and their gradients, fp32.
 * params / grads: 8 DEVICE pointers in torch layouts: weight_ih_l0 [768][9], bias_ih_l0 [768], bias_hh_l0 [768], the same three of the reverse direction,
 * ln.weight [268], ln.bias [268].  obs [*][obs_dim]; idx (may be NULL): int64[B], the row of `obs` each sample reads (the minibatch is never materialised);
 * feat / dfeat [B][268]; stats [B][2] (LayerNorm mean and rstd, saved by the forward for the backward).  uaisa_reader_bwd ADDS this minibatch's gradients to
 * `grads`.  hsum (may be NULL in both): [B][256], the pre-LayerNorm sum of the two directions' states -- saved by the forward, it spares the backward a second
 * evaluation of the gates.  Asynchronous on `stream`, capturable in a CUDA graph. */
int uaisa_reader_fwd(const float *obs, const int64_t *idx, int32_t B, int32_t obs_dim, const float *const *params, float *feat, float *stats, float *hsum, void *stream);
int uaisa_reader_bwd(const float *obs, const int64_t *idx, int32_t B, int32_t obs_dim, const float *const *params, const float *dfeat, const float *stats,
                     const float *hsum, float *const *grads, void *stream);

/* ---- actor / critic MLP forward and backward of the PPO update on the tensor cores (csrc/uaisa_mlp.cu) -----------------------------------------
 * Replaces torch autograd's GEMMs for GaussianActor.net_out / Critic.v_net (policy_rnn_ac.py:206-257; backward through multi_ppo.py:379-404).
 * uaisa_mlp_gemm:  C[M][N] (+)= A * B^T, tcgen05.mma kind::tf32 (fp32 storage, TF32 operand precision, fp32 accumulation), N <= 288.
 *   a_mn = 0: A is stored [M][lda] (K contiguous); a_mn = 1: A is stored [K][lda] (M contiguous) -- likewise b_mn for B ([N][ldb] / [K][ldb]).
 *   splits > 1 cuts K over that many CTAs per row tile; it requires epi = 4 and a C the caller has zeroed (or wants added to).
 *   epi: 0 store | 1 relu(acc + bias[n]) | 2 = 1 plus Y[m][o] = sum_n C[m][n] W3[o][n] + b3[o] (o < nout <= 3; C may be NULL) |
 *        3 acc where mask[m][n] > 0 else 0 (mask [M][ldm]) | 4 atomic add into C.
 * uaisa_mlp_out_bwd: dZ2[b][j] = H2[b][j] > 0 ? sum_o dy[b][o] W3[o][j] : 0; ADDS db2[j], dW3[o][j], db3[o] (hidden width 256, nout <= 3).
 * uaisa_colsum: out[j] += sum_r X[r][j].   All asynchronous on `stream`, capturable in a CUDA graph; 0 ok, -1 bad arguments, -2 launch failure. */
int uaisa_mlp_gemm(const float *A, int32_t lda, int32_t a_mn, const float *B, int32_t ldb, int32_t b_mn, float *C, int32_t ldc, int32_t M, int32_t N, int32_t K,
                    int32_t splits, int32_t epi, const float *bias, const float *mask, int32_t ldm, const float *W3, const float *b3, float *Y, int32_t nout,
                    void *stream);
int uaisa_mlp_out_bwd(const float *dy, const float *H2, const float *W3, int32_t B, int32_t nout, float *dZ2, float *dW3, float *db3, float *db2, void *stream);
int uaisa_colsum(const float *X, int32_t rows, int32_t cols, float *out, void *stream);

/* PPO losses of one update iteration and their gradients with respect to the heads' outputs (multi_ppo.compute_loss_pi / compute_loss_v, multi_ppo.py:379-404),
 * one thread per minibatch sample; idx (may be NULL): int64[B], the buffer row of each sample; ok: uint8 per buffer row (0 = advantage / return not finite:
 * masked out).  uaisa_ppo_wsum writes scal[0] = max(1, usable samples); scal is float[4], zeroed ONCE by the caller ([1..2] are the kernel's counters, re-armed by itself).  uaisa_ppo_pi_loss: y [B][3] = the actor's pre-tanh output -> dy [B][3] =
 * d(loss / W) / dy with inv_world = 1 / W; ADDS d/dlog_std to dlogstd[3], approx-KL / W to slots[0] and loss / W to slots[1].  uaisa_ppo_v_loss: y [B] =
 * the critic's output -> dy [B]; ADDS the loss to loss_out[0].  Asynchronous on `stream`, capturable in a CUDA graph. */
int uaisa_ppo_wsum(const uint8_t *ok, const int64_t *idx, int32_t B, float *scal, void *stream);
int uaisa_ppo_pi_loss(const float *y, const float *log_std, float std_factor, const float *act, const float *logp_old, const float *adv, const uint8_t *ok,
                      const int64_t *idx, int32_t B, float clip_ratio, float inv_world, const float *scal, float *dy, float *dlogstd, float *slots, void *stream);
int uaisa_ppo_v_loss(const float *y, const float *ret, const uint8_t *ok, const int64_t *idx, int32_t B, float inv_world, const float *scal, float *dy,
                     float *loss_out, void *stream);

/* ---- latency-optimised host round trip ------------------------------------------------------------------------
 * One env step with HOST action in / HOST results out, returned in the reference's own ragged spirit: the 12 self
 * values of every observation plus only the neighbour rows that exist (an observation is 12 + 9*max(k,1) values in
 * the reference, ir_gym.py:203-229; the fixed-slot padding is never shipped over PCIe).  Internally ONE CUDA graph is
 * launched per call and holds a single node: the env kernel writes the result block straight into mapped pinned memory and
 * its last CTA raises a ready word the host polls (no copy-engine D2H, no stream synchronisation).  Action blocks up to
 * 1 MiB are read by the kernel from mapped pinned memory; larger ones are copied by the copy engine straight from the
 * caller's buffer (pinned memory makes that a direct DMA).  Pointers in `out` refer to library-owned pinned host memory and
 * stay valid until the next call on this
 * handle.  Synchronous. */
typedef struct UaisaRowRec {
    float v[9];        /* apex(3) rel_pos(3) alpha min_dis 1/(t+0.2), rounded to 2 dp */
    int32_t agent;     /* a = env * N + drone */
    int32_t slot;      /* position r of the row in that agent's observation (0 = least urgent, count-1 = most urgent) */
    int32_t nbr;       /* in-env drone index of the neighbour */
} UaisaRowRec;
typedef struct UaisaHostStep {
    const float *reward;       /* [A] */
    const float *self_obs;     /* [A][12] */
    const uint8_t *done, *info, *finish, *count;   /* [A] each */
    const UaisaRowRec *rows;   /* [n_rows], unordered */
    int32_t n_rows;
    int32_t reserved;
} UaisaHostStep;
int uaisa_step_host_sparse(UaisaHandle *h, const float *action /* HOST [A][3] */, int32_t auto_reset, int32_t max_ep_len,
                           UaisaHostStep *out);

#ifdef __cplusplus
}
#endif
#endif /* UAISA_B200_H */

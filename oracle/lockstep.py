"""TEST INFRASTRUCTURE ONLY: lock-step driver that runs the imported reference env
(oracle/ref_shim.py) and records everything a parity check needs.  Used by
tests/test_oracle_vs_reference.py (build container only) and oracle/gen_golden.py."""
from __future__ import annotations

import numpy as np

import ref_shim


def training_like_actions(rng, vel):
    a_inc = np.round(rng.uniform(-1, 1, size=vel.shape), 2)  # multi_ppo.py:197-205
    return np.round(0.5 * a_inc + vel, 2)


def goal_seeking_actions(rng, des, noise=0.3):
    return np.round(des + rng.normal(0, noise, size=des.shape), 2)  # gym_env_test.py:12 + noise


def steering_actions(rng, des, vel, yaw, pitch, speed=0.8, noise=0.05):
    des = np.asarray(des, np.float64)
    sp = np.linalg.norm(vel, axis=-1)
    yaw_des = np.degrees(np.arctan2(des[..., 1], des[..., 0]))
    pitch_des = np.degrees(np.arctan2(des[..., 2], np.linalg.norm(des[..., :2], axis=-1)))
    dyaw = (yaw_des - yaw + 180.0) % 360.0 - 180.0
    moving = np.linalg.norm(des, axis=-1) > 0
    a0 = np.where(moving, speed - sp, -sp)
    a = np.stack([a0, dyaw / 90.0, (pitch_des - pitch) / 90.0], axis=-1)
    a = a + rng.normal(0, noise, size=a.shape)
    return np.round(np.clip(a, -1, 1), 2)


def mixed_actions(rng, step, des, st, block=50):
    mode = (step // block) % 4
    if mode == 0:
        return training_like_actions(rng, st["vel"])
    if mode == 1:
        return goal_seeking_actions(rng, des)
    return steering_actions(rng, des, st["vel"], st["yaw"], st["pitch"], speed=0.6 if mode == 2 else 1.0,
                            noise=0.05 if mode == 2 else 0.2)


def run_reference(scenario_dict, nm, steps, seed, reset_on=("done", "finish"), full_reset_every=0, block=50, f32=False):
    """Run the reference; return a dict of stacked per-step arrays (a golden trace).

    Per step t the trace holds the state BEFORE the step, the action, the step outputs, the state
    AFTER the step, the reset mask applied afterwards and the observation returned by
    ``env_observation()`` after those resets (only when any reset happened).

    f32 = True: before every step (and before the post-reset observation) the reference is re-loaded with float64 COPIES of the fp32 image of its own
    state, and it is stepped with the fp32 image of the action -- what a device that stores fp32 state would hand it.  The
    recorded pre-states / actions are then exactly representable on the device, so the CUDA path must reproduce the recorded outputs with NO
    input-rounding slack (tests/test_gpu_parity.py::test_against_reference_golden_exact)."""

    def q32(st):
        return {k: (np.asarray(v, np.float64).astype(np.float32).astype(np.float64) if np.asarray(v).dtype.kind == "f" else v) for k, v in st.items()}
    env = ref_shim.make_reference_env(scenario_dict, nm=nm)
    rng = np.random.default_rng(seed)
    N = len(env.ir_gym.drone_list)
    obs0, cnt0 = ref_shim.pad_obs(env.drone_reset(False), nm)
    tr = {k: [] for k in ("action", "obs", "count", "reward", "done", "info", "finish", "reset_mask", "obs_after_reset",
                          "count_after_reset", "full_reset")}
    pre = {k: [] for k in ref_shim.STATE_FIELDS}
    post = {k: [] for k in ref_shim.STATE_FIELDS}
    for t in range(steps):
        st = ref_shim.get_state(env)
        if f32:
            st = q32(st)
            ref_shim.set_state(env, st)
        for k in pre:
            pre[k].append(st[k])
        des = np.array(env.ir_gym.cal_des_list(), dtype=np.float64)
        act = mixed_actions(rng, t, des, st, block)
        if f32:
            act = act.astype(np.float32).astype(np.float64)
        o, r, d, i, f = env.drone_step([a for a in act])
        po, pc = ref_shim.pad_obs(o, nm)
        st2 = ref_shim.get_state(env)
        for k in post:
            post[k].append(st2[k])
        d = np.array(d, dtype=np.uint8); f = np.array(f, dtype=np.uint8)
        tr["action"].append(act); tr["obs"].append(po); tr["count"].append(pc)
        tr["reward"].append(np.array(r, dtype=np.float64)); tr["done"].append(d)
        tr["info"].append(np.array(i, dtype=np.uint8)); tr["finish"].append(f)
        mask = np.zeros(N, np.uint8)
        if "done" in reset_on:
            mask |= d
        if "finish" in reset_on:
            mask |= f
        full = bool(full_reset_every and (t + 1) % full_reset_every == 0)
        if full:
            mask[:] = 1
            po2, pc2 = ref_shim.pad_obs(env.drone_reset(False), nm)
        elif mask.any():
            if f32:
                ref_shim.set_state(env, q32(ref_shim.get_state(env)))      # the device resets / re-observes from its fp32 post-step state
            for k in np.nonzero(mask)[0]:
                env.drone_reset_one(False, int(k))
            po2, pc2 = ref_shim.pad_obs(env.ir_gym.env_observation(), nm)
        else:
            po2, pc2 = np.zeros_like(po), np.zeros_like(pc)
        tr["reset_mask"].append(mask); tr["obs_after_reset"].append(po2); tr["count_after_reset"].append(pc2)
        tr["full_reset"].append(np.uint8(full))
    out = {k: np.stack(v) for k, v in tr.items()}
    out.update({"pre_" + k: np.stack(v) for k, v in pre.items()})
    out.update({"post_" + k: np.stack(v) for k, v in post.items()})
    out["obs0"], out["count0"] = obs0, cnt0
    out["nm"] = np.int32(nm)
    return out

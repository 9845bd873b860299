"""Shared helpers for the test-suite (path setup, action streams, comparisons)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def training_like_actions(rng, vel):
    """multi_ppo.py:197-205: a_inc~U(-1,1)^3 rounded 2dp; action = round2(0.5*a_inc + v)."""
    a_inc = np.round(rng.uniform(-1, 1, size=vel.shape), 2)
    return np.round(0.5 * a_inc + vel, 2)


def goal_seeking_actions(rng, des, noise=0.3):
    """Desired-velocity controller of uaisa_env/gym_env_test.py:12 plus noise, rounded 2dp."""
    return np.round(des + rng.normal(0, noise, size=des.shape), 2)


def steering_actions(rng, des, vel, yaw, pitch, speed=0.8, noise=0.05):
    """A controller that actually reaches waypoints under the env's [acc, yaw/90, pitch/90]
    interpretation of the action (drone.py:435-447): steer yaw/pitch onto the desired direction
    and hold `speed`.  Rounded to 2 dp like the trainer's actions."""
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


# ---- deterministic steering policy for the evaluation-loop parity test (oracle/gen_eval_golden.py, tests/test_post_train_gpu.py) --------
# A function of the observation only: estimate heading from the velocity, turn towards the desired velocity, hold 0.8 speed.  The
# numpy form runs inside the reference's policy_test (one observation at a time, float64 2-dp values); the torch form first recovers
# those 2-dp doubles from the fp32 observation image and then evaluates the same float64 formulas, so both return the same float32.
def goal_policy_np(o):
    o = np.asarray(o, dtype=np.float64)
    vel, des = o[3:6], o[8:11]
    speed = np.sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2])
    moving = speed > 1e-6
    yaw = np.degrees(np.arctan2(vel[1], vel[0])) if moving else 0.0
    pitch = np.degrees(np.arcsin(np.clip(vel[2] / speed, -1.0, 1.0))) if moving else 0.0
    dn = np.sqrt(des[0] * des[0] + des[1] * des[1] + des[2] * des[2])
    if dn > 1e-9:
        dyaw = np.degrees(np.arctan2(des[1], des[0])) - yaw
        dyaw = (dyaw + 180.0) % 360.0 - 180.0
        dpitch = np.degrees(np.arcsin(np.clip(des[2] / dn, -1.0, 1.0))) - pitch
    else:
        dyaw = dpitch = 0.0
    ctrl = np.array([np.clip(0.8 - speed, -1.0, 1.0), np.clip(dyaw / 90.0, -1.0, 1.0), np.clip(dpitch / 90.0, -1.0, 1.0)])
    return np.clip(ctrl - vel, -2.0, 2.0).astype(np.float32)          # action = a_inc + vel = the controls (acceler_vel = 1)


def goal_policy_torch(obs, count=None):
    import torch
    o = torch.round(obs.double() * 100.0) / 100.0
    vel, des = o[:, 3:6], o[:, 8:11]
    speed = torch.sqrt(vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1] + vel[:, 2] * vel[:, 2])
    moving = speed > 1e-6
    safe = torch.where(moving, speed, torch.ones_like(speed))
    yaw = torch.where(moving, torch.rad2deg(torch.atan2(vel[:, 1], vel[:, 0])), torch.zeros_like(speed))
    pitch = torch.where(moving, torch.rad2deg(torch.asin(torch.clamp(vel[:, 2] / safe, -1.0, 1.0))), torch.zeros_like(speed))
    dn = torch.sqrt(des[:, 0] * des[:, 0] + des[:, 1] * des[:, 1] + des[:, 2] * des[:, 2])
    has = dn > 1e-9
    dsafe = torch.where(has, dn, torch.ones_like(dn))
    dyaw = torch.rad2deg(torch.atan2(des[:, 1], des[:, 0])) - yaw
    dyaw = torch.remainder(dyaw + 180.0, 360.0) - 180.0
    dpitch = torch.rad2deg(torch.asin(torch.clamp(des[:, 2] / dsafe, -1.0, 1.0))) - pitch
    dyaw = torch.where(has, dyaw, torch.zeros_like(dyaw)); dpitch = torch.where(has, dpitch, torch.zeros_like(dpitch))
    ctrl = torch.stack([torch.clamp(0.8 - speed, -1.0, 1.0), torch.clamp(dyaw / 90.0, -1.0, 1.0), torch.clamp(dpitch / 90.0, -1.0, 1.0)], dim=1)
    return torch.clamp(ctrl - vel, -2.0, 2.0).float()


def eval_noise_table(episodes, steps, n_drones, seed=0, scale=0.3):
    """Per (episode, step in episode, drone) action noise shared by both sides of the evaluation-loop parity test (float32, 2-dp grid).
    Indexing by episode and step-in-episode (not by a global step) keeps later episodes aligned when free-running fp32 drift ends
    one episode a step early or late."""
    return np.round(np.random.default_rng(seed).uniform(-scale, scale, (episodes, steps, n_drones, 3)), 2).astype(np.float32)


class NoisyGoalPolicyNp:
    """goal_policy_np + table noise; the reference calls it once per drone per step (post_train.py:62-66).  An episode start is
    recognised from the reference env itself: drone 0 has not moved yet (drone.reset zeroes real_route_len, drone.py:270-291)."""

    def __init__(self, table, env):
        self.table, self.env, self.ep, self.t, self.i = table, env, -1, 0, 0

    def __call__(self, o):
        n = self.table.shape[2]
        if self.i == 0:
            if float(self.env.ir_gym.drone_list[0].real_route_len) == 0.0:
                self.ep += 1; self.t = 0
            else:
                self.t += 1
        out = (goal_policy_np(o) + self.table[self.ep % self.table.shape[0], self.t % self.table.shape[1], self.i]).astype(np.float32)
        self.i = (self.i + 1) % n
        return out


class NoisyGoalPolicyTorch:
    """Same policy for the batched loop: one call per step for all E*N agents, told each env's episode index and step."""
    wants_episode_info = True

    def __init__(self, table, device="cuda"):
        import torch
        self.table = torch.from_numpy(table).to(device)

    def __call__(self, obs, count, ep_index, ep_step):
        rows = self.table[ep_index % self.table.shape[0], ep_step % self.table.shape[1]]        # [E, N, 3]
        return (goal_policy_torch(obs) + rows.reshape(-1, 3)).float()

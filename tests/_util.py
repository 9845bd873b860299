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

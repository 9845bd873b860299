"""CPU: the episode bookkeeping of the batched evaluation loop (uaisa_ppo.post_train) on a scripted stand-in for VecMdin --
episode ends (done / max length / all finished), per-env quotas, drone-0 returns incl. inf, speed averages, the result line."""
import math

import numpy as np
import torch

import _util  # noqa: F401
from uaisa_ppo.post_train import post_train


class ScriptedVec:
    """E envs x N drones; env e collides at step `done_at[e]` of every episode, or all its drones finish at `finish_at[e]`."""

    def __init__(self, E, N, done_at, finish_at, speed):
        self.E, self.N, self.A, self.device = E, N, E * N, torch.device("cpu")
        self.done_at, self.finish_at = torch.tensor(done_at), torch.tensor(finish_at)
        self.speed = torch.tensor(speed, dtype=torch.float32)
        self.obs = torch.zeros(self.A, 21); self.count = torch.zeros(self.A, dtype=torch.uint8)
        self.reward = torch.zeros(self.A); self.done = torch.zeros(self.A, dtype=torch.uint8)
        self.info = torch.zeros(self.A, dtype=torch.uint8); self.finish = torch.zeros(self.A, dtype=torch.uint8)
        self.t = torch.zeros(E, dtype=torch.long)          # steps since the env's last reset
        self.resets = 0

    def reset(self, mask=None, observe=True):
        if mask is None:
            self.t.zero_()
        else:
            self.t[mask.view(self.E, self.N)[:, 0].bool()] = 0
            self.resets += int(mask.view(self.E, self.N)[:, 0].sum())
        return self.obs

    def get_state(self):
        vp = torch.zeros(self.A, 4)
        vp[:, 0] = self.speed.repeat_interleave(self.N)     # |v| = speed of the env
        return {"vel_pitch": vp}

    def step(self, action):
        assert action.shape == (self.A, 3) and action.dtype == torch.float32
        self.t += 1
        d = (self.t == self.done_at); f = (self.t == self.finish_at)
        self.done = d.repeat_interleave(self.N).to(torch.uint8)
        self.finish = f.repeat_interleave(self.N).to(torch.uint8); self.info = self.finish.clone()
        r = torch.full((self.E, self.N), -1.0); r[:, 1:] = 100.0          # only drone 0 counts (post_train.py:82)
        r[f, 0] = float("inf")                                           # the reference's inf reward at the destination
        self.reward = r.view(-1)


def test_episode_bookkeeping_and_result_line(tmp_path):
    # env 0: collision at step 3; env 1: all finished at step 5; env 2: nothing happens -> cut at max_ep_len = 7
    env = ScriptedVec(3, 2, done_at=[3, 99, 99], finish_at=[99, 5, 99], speed=[0.5, 1.0, 0.25])
    pt = post_train(env, num_episodes=6, max_ep_len=7)
    res = pt.policy_test(lambda obs, cnt: torch.zeros(obs.shape[0], 3), policy_name="scripted", result_path=str(tmp_path))
    eps = res["episodes"]
    assert len(eps) == 6 and env.resets == 6
    by_env = {e: [ep for ep in eps if ep["env"] == e] for e in range(3)}
    assert [ep["ep_len"] for ep in by_env[0]] == [3, 3] and [ep["ep_len"] for ep in by_env[1]] == [5, 5] and [ep["ep_len"] for ep in by_env[2]] == [7, 7]
    assert all(not ep["arrived"] and not ep["finished"] for ep in by_env[0] + by_env[2]) and all(ep["arrived"] and ep["finished"] for ep in by_env[1])
    assert [ep["ep_ret"] for ep in by_env[0]] == [-3.0, -3.0] and all(math.isinf(ep["ep_ret"]) for ep in by_env[1]) and by_env[2][1]["ep_ret"] == -7.0
    assert all(abs(ep["speed"] - s) < 1e-6 for e, s in ((0, 0.5), (1, 1.0), (2, 0.25)) for ep in by_env[e])
    assert [(ep["step"], ep["env"]) for ep in eps] == sorted((ep["step"], ep["env"]) for ep in eps)      # (step, env) order
    assert res["success_rate"] == 2 / 6 and res["mean_len"] == 5.0 and res["std_len"] == 0.0
    assert res["average_speed"] == float(np.round(np.mean([0.5, 0.5, 1.0, 1.0, 0.25, 0.25]), 2))
    line = open(str(tmp_path) + "/result.txt").read().strip()
    assert line == res["line"] == "policy_name: scripted  successful rate: 33.33% average EpLen: 5.0 std length 0.0 average speed: 0.58 std speed 0.31"


def test_action_rule_rounds_the_increment_not_the_action():
    env = ScriptedVec(1, 2, done_at=[2], finish_at=[99], speed=[0.3])
    seen = []
    orig = env.step
    env.step = lambda a: (seen.append(a.clone()), orig(a))[1]
    post_train(env, num_episodes=1, max_ep_len=5, acceler_vel=0.5).policy_test(lambda obs, cnt: torch.full((obs.shape[0], 3), 0.126), policy_name="x")
    a = seen[0]
    # 0.5 * round(0.126, 2) + v, v = (0.3, 0, 0): the increment is rounded to 2 dp, the sum is not
    assert torch.allclose(a[:, 0], torch.tensor(0.5 * 0.13 + 0.3)) and torch.allclose(a[:, 1], torch.tensor(0.065))

"""Batched evaluation loop: ``post_train.policy_test`` of the reference (train/policy/post_train.py:11-128) over E
environments at once on the GPU.

Reference semantics kept per environment:
  * action  = ``acceler_vel * np.round(policy(o), 2) + drone.vel``                       (post_train.py:64-73; NOT re-rounded)
  * an episode ends when any drone is done, after ``max_ep_len`` steps, or when all drones finished (:86); then the whole
    environment is reset (:99)
  * episode return = sum of drone 0's reward (:82); episode speed = mean over steps of the mean drone speed (:78-80, :87)
  * an episode is "successful" for the length statistics when every drone's info flag is set (:89-90) and counts towards the
    success rate when every drone finished (:103-104)
  * the result line has the reference's format (:122-128)

Every environment runs ``ceil(num_episodes / E)`` episodes (no bias towards short episodes); with E = 1 the episode order is the
reference's.  All per-step work stays on the device; the host only learns how many environments still have episodes to run.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch


class post_train:
    def __init__(self, env, num_episodes: int = 100, max_ep_len: int = 150, acceler_vel: float = 1.0, std_factor: float = 0.001,
                 inf_print: bool = False, **kwargs):
        self.env = env                                  # uaisa_env.vec_env.VecMdin
        self.num_episodes, self.max_ep_len = int(num_episodes), int(max_ep_len)
        self.acceler_vel, self.std_factor, self.inf_print = float(acceler_vel), float(std_factor), bool(inf_print)
        self.drone_number = env.N

    # ------------------------------------------------------------------------------------------------------------
    def _policy_fn(self, policy) -> Callable:
        """``policy``: an ``rnn_ac`` (sampled with ``std_factor`` like load_policy/get_action, post_train.py:131-152) or a
        callable ``f(obs[A, obs_dim], count[A]) -> a_inc[A, 3]`` on CUDA tensors; a callable with ``wants_episode_info = True`` is also
        given each env's episode index and step within the episode (scripted policies)."""
        if hasattr(policy, "step") and hasattr(policy, "pi"):
            policy.eval()
            return lambda obs, cnt: policy.step(obs, cnt, self.std_factor)[0]
        return policy

    @torch.no_grad()
    def policy_test(self, policy, policy_name: str = "policy", result_path: Optional[str] = None, result_name: str = "/result.txt") -> dict:
        env = self.env
        E, N, dev = env.E, env.N, env.device
        per_env = max(1, math.ceil(self.num_episodes / E))
        fn = self._policy_fn(policy)
        env.reset()
        ep_ret = torch.zeros(E, dtype=torch.float64, device=dev)
        sp_sum = torch.zeros(E, dtype=torch.float64, device=dev)
        ep_len = torch.zeros(E, dtype=torch.int64, device=dev)
        n_done = torch.zeros(E, dtype=torch.int64, device=dev)
        rec = []                                            # (step, env ids, return, length, speed, all-info, all-finish) of the episodes that ended
        t = 0
        while True:
            running = n_done < per_env
            if t % 8 == 0 and not bool(running.any()):     # the only host synchronisation of the loop
                break
            a_inc = (fn(env.obs, env.count, n_done, ep_len) if getattr(fn, "wants_episode_info", False) else fn(env.obs, env.count)).to(torch.float32)
            a_inc = torch.round(a_inc * 100.0) / 100.0                                   # np.round(float32, 2)
            vel = env.get_state()["vel_pitch"][:, :3]
            action = ((self.acceler_vel * a_inc).double() + vel.double()).float().contiguous()   # float32(a) promoted, + float64 velocity
            env.step(action)
            vel = env.get_state()["vel_pitch"][:, :3].double()
            sp_sum += vel.norm(dim=1).view(E, N).mean(dim=1)
            ep_ret += env.reward.view(E, N)[:, 0].double()
            ep_len += 1
            done_e = env.done.view(E, N).bool().any(dim=1)
            fin_e = env.finish.view(E, N).bool().all(dim=1)
            info_e = env.info.view(E, N).bool().all(dim=1)
            end = running & (done_e | (ep_len == self.max_ep_len) | fin_e)
            rec.append((t, end, ep_ret.clone(), ep_len.clone(), sp_sum / ep_len.double(), info_e, fin_e))
            env.reset(end.repeat_interleave(N).to(torch.uint8))                          # whole-env reset + re-observation of those envs
            ep_ret = torch.where(end, torch.zeros_like(ep_ret), ep_ret)                   # (not a multiplication: returns can be inf, ir_gym.py:88)
            sp_sum = torch.where(end, torch.zeros_like(sp_sum), sp_sum)
            ep_len = torch.where(end, torch.zeros_like(ep_len), ep_len)
            n_done += end.long()
            t += 1
        # ---- gather the episodes in (step, env) order -- the reference's order when E == 1
        eps = []
        for (tt, end, r_, l_, s_, i_, f_) in rec:
            idx = torch.nonzero(end).flatten()
            if idx.numel():
                for e, r, l, s, i, f in zip(idx.tolist(), r_[idx].tolist(), l_[idx].tolist(), s_[idx].tolist(), i_[idx].tolist(), f_[idx].tolist()):
                    eps.append({"step": tt, "env": e, "ep_ret": r, "ep_len": l, "speed": s, "arrived": bool(i), "finished": bool(f)})
        if self.inf_print:
            for n, ep in enumerate(eps):
                print("%s, Episode %d \t EpRet %.3f \t EpLen %d \t EpSpeed  %.3f" % ("Successful" if ep["arrived"] else "Fail", n, ep["ep_ret"], ep["ep_len"], ep["speed"]))
        lens = [ep["ep_len"] for ep in eps if ep["arrived"]]
        speeds = [ep["speed"] for ep in eps]
        sn = sum(ep["finished"] for ep in eps)
        mean_len = 0 if not lens else np.round(np.mean(lens), 2)
        std_len = 0 if not lens else np.round(np.std(lens), 2)
        average_speed = np.round(np.mean(speeds), 2)
        std_speed = np.round(np.std(speeds), 2)
        line = " ".join(str(x) for x in ("policy_name: " + policy_name, " successful rate: {:.2%}".format(sn / len(eps)), "average EpLen:", mean_len,
                                         "std length", std_len, "average speed:", average_speed, "std speed", std_speed))
        if result_path is not None:
            with open(result_path + result_name, "a") as f:
                print(line, file=f)
        print(line)
        return {"episodes": eps, "success_rate": sn / len(eps), "mean_len": float(mean_len), "std_len": float(std_len),
                "average_speed": float(average_speed), "std_speed": float(std_speed), "line": line, "env_steps": t}

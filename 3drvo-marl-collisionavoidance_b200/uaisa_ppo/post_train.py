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

``policy = "rvo"`` runs the non-learned baseline instead: the classic RVO velocity selection (reciprocal_vel_obs.cal_vel,
uaisa_env/vel_obs/reciprocal_vel_obs.py:21-166 -- dead code in the reference, here kernel K7 ``uaisa_rvo_select``) picks a velocity per drone
and the env is fed that velocity expressed in its own controls (rounded to 2 dp like every action of the trainer).

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
        self.rvo_vmax, self.rvo_acceler = float(kwargs.get("rvo_vmax", 2.0)), float(kwargs.get("rvo_acceler", 0.5))     # ir_gym.py:18-29 constructor values

    # ------------------------------------------------------------------------------------------------------------
    def _policy_fn(self, policy, static_shapes: bool) -> Callable:
        """``policy``: an ``rnn_ac`` (sampled with ``std_factor`` like load_policy/get_action, post_train.py:131-152), a
        ``FusedPolicy`` around one (fused forward: K6a features + tcgen05 heads), or a
        callable ``f(obs[A, obs_dim], count[A]) -> a_inc[A, 3]`` on CUDA tensors; a callable with ``wants_episode_info = True`` is also
        given each env's episode index and step within the episode (scripted policies)."""
        if isinstance(policy, str):
            if policy != "rvo":
                raise ValueError("the only scripted policy is 'rvo' (classic RVO velocity selection)")
            self._reader = None
            return None
        if hasattr(policy, "step") and (hasattr(policy, "pi") or hasattr(policy, "ac")):     # rnn_ac, or uaisa_ppo.fused.FusedPolicy wrapping one
            module = policy if hasattr(policy, "pi") else policy.ac
            module.eval()
            # inside a CUDA graph the GRU path of agents with >= 2 neighbour rows needs a static capacity (checked after the run)
            cap = max(64, self.env.A // 4) if static_shapes else 0
            self._reader = module.pi.rnn_reader if static_shapes else None
            if static_shapes and getattr(self._reader, "overflow", None) is None:
                self._reader.overflow = torch.zeros((), dtype=torch.int64, device=self.env.device)
            return lambda obs, cnt: policy.step(obs, cnt, self.std_factor, multi_cap=cap)[0]
        self._reader = None
        return policy

    @torch.no_grad()
    def policy_test(self, policy, policy_name: str = "policy", result_path: Optional[str] = None, result_name: str = "/result.txt",
                    use_cuda_graph: bool = False) -> dict:
        """``use_cuda_graph``: capture one evaluation step and replay it (static-capacity GRU path for agents with >= 2 neighbour rows).
        Measured no faster than the eager loop (profiles/eval_throughput.py: the step is bound by the policy forward, not by launches),
        so it is off by default."""
        env = self.env
        E, N, dev = env.E, env.N, env.device
        per_env = max(1, math.ceil(self.num_episodes / E))
        cap = per_env * self.max_ep_len + 32
        use_cuda_graph = bool(use_cuda_graph) and cap * E <= 200_000_000
        fn = self._policy_fn(policy, use_cuda_graph)
        info_fn = getattr(fn, "wants_episode_info", False)
        env.reset()
        # persistent per-env accumulators and the episode log, all on the device and all updated in place (one captured step replays them)
        ep_ret = torch.zeros(E, dtype=torch.float64, device=dev)
        sp_sum = torch.zeros(E, dtype=torch.float64, device=dev)
        ep_len = torch.zeros(E, dtype=torch.int64, device=dev)
        n_done = torch.zeros(E, dtype=torch.int64, device=dev)
        t_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        left = torch.ones((), dtype=torch.int64, device=dev)                 # envs that still have episodes to run
        log_end = torch.zeros((cap, E), dtype=torch.bool, device=dev)
        log_ret = torch.zeros((cap, E), dtype=torch.float64, device=dev); log_spd = torch.zeros((cap, E), dtype=torch.float64, device=dev)
        log_len = torch.zeros((cap, E), dtype=torch.int32, device=dev)
        log_info = torch.zeros((cap, E), dtype=torch.bool, device=dev); log_fin = torch.zeros((cap, E), dtype=torch.bool, device=dev)

        def one_step():
            running = n_done < per_env
            if fn is None:                                                               # scripted baseline: K7 velocity selection -> env controls
                _, ctrl = env.rvo_select(self.rvo_vmax, self.rvo_acceler)
                action = (torch.round(ctrl.double() * 100.0) / 100.0).float().contiguous()
            else:
                a_inc = (fn(env.obs, env.count, n_done, ep_len) if info_fn else fn(env.obs, env.count)).to(torch.float32)
                a_inc = torch.round(a_inc * 100.0) / 100.0                               # np.round(float32, 2)
                vel = env.get_state()["vel_pitch"][:, :3]
                action = ((self.acceler_vel * a_inc).double() + vel.double()).float().contiguous()   # float32(a) promoted, + float64 velocity
            env.step(action)
            vel = env.get_state()["vel_pitch"][:, :3].double()
            sp_sum.add_(vel.norm(dim=1).view(E, N).mean(dim=1))
            ep_ret.add_(env.reward.view(E, N)[:, 0].double())
            ep_len.add_(1)
            done_e = env.done.view(E, N).bool().any(dim=1)
            fin_e = env.finish.view(E, N).bool().all(dim=1)
            info_e = env.info.view(E, N).bool().all(dim=1)
            end = running & (done_e | (ep_len == self.max_ep_len) | fin_e)
            log_end.index_copy_(0, t_dev, end[None]); log_ret.index_copy_(0, t_dev, ep_ret[None]); log_len.index_copy_(0, t_dev, ep_len.int()[None])
            log_spd.index_copy_(0, t_dev, (sp_sum / ep_len.double())[None]); log_info.index_copy_(0, t_dev, info_e[None]); log_fin.index_copy_(0, t_dev, fin_e[None])
            env.reset(end.repeat_interleave(N).to(torch.uint8))                          # whole-env reset + re-observation of those envs
            ep_ret.masked_fill_(end, 0.0); sp_sum.masked_fill_(end, 0.0); ep_len.masked_fill_(end, 0)   # (not a multiplication: returns can be inf, ir_gym.py:88)
            n_done.add_(end.long())
            t_dev.add_(1)
            left.copy_((n_done < per_env).sum())

        graph = None
        if use_cuda_graph:
            if self._reader is not None:
                self._reader.overflow.zero_()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):               # warm-up on a side stream: these are real steps of the evaluation
                one_step(); one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                one_step()
        while int(left) > 0:                            # the only host synchronisations of the loop: every 16 steps
            room = cap - 1 - int(t_dev)
            if room <= 0:
                break
            for _ in range(min(16, room)):
                graph.replay() if graph is not None else one_step()
        if self._reader is not None and int(self._reader.overflow) > 0:
            raise RuntimeError("more agents with >= 2 neighbour rows than the static GRU capacity of the captured step; rerun with use_cuda_graph=False")
        t = int(t_dev)
        # ---- gather the episodes in (step, env) order -- the reference's order when E == 1
        te = torch.nonzero(log_end[:t]).cpu().numpy()
        pick = lambda L: L[:t][log_end[:t]].cpu().numpy()  # noqa: E731   (row-major boolean selection = (step, env) order)
        r_, l_, s_, i_, f_ = pick(log_ret), pick(log_len), pick(log_spd), pick(log_info), pick(log_fin)
        eps = [{"step": int(te[q, 0]), "env": int(te[q, 1]), "ep_ret": float(r_[q]), "ep_len": int(l_[q]), "speed": float(s_[q]),
                "arrived": bool(i_[q]), "finished": bool(f_[q])} for q in range(te.shape[0])]
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

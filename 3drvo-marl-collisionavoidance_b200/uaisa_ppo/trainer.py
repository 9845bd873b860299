"""Batched PPO rollout + update around VecMdin: the data-parallel form of ``multi_ppo.training_loop``
(train/policy/multi_ppo.py:166-339).

What is kept from the reference (cited per line below): the per-step sequence (policy step -> 2-dp action
transform -> env step -> store -> trainer reset policy -> re-observe), GAE(gamma .99, lambda .97) with last_val = 0
and paths cut exactly where ``finish_path`` is called, NO advantage normalisation, the PPO-clip loss, the approx-KL
early stop, grad-norm clipping at 2.0 over ALL actor-critic parameters for the policy step, separate Adam optimisers for
pi and v that both touch the shared reader, and the hyper-parameters of train_process.py:53-63.

What changes (documented deviation, SURVEY.md section 7 "PPO semantics vs batching"): the reference loops over
robots and does 50+50 FULL-BATCH iterations on each robot's own 300 samples; here every agent of every env feeds one
batch that is cut into minibatches, and with world_size > 1 the gradients are all-reduced (NCCL, one flat bucket per
iteration: 680 991 fp32 = 2.7 MB, plus the KL scalar) before the identical clip + Adam step on every rank.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from uaisa_env import _capi
from uaisa_env.vec_env import VecMdin
from .policy import rnn_ac


@dataclass
class PPOConfig:                      # train_process.py:53-63 defaults
    steps_per_epoch: int = 300
    max_ep_len: int = 500
    gamma: float = 0.99
    lam: float = 0.97
    clip_ratio: float = 0.2
    pi_lr: float = 4e-6
    vf_lr: float = 5e-5
    train_pi_iters: int = 50
    train_v_iters: int = 50
    target_kl: float = 0.05
    max_grad_norm: float = 2.0        # multi_ppo.py:367
    minibatch: int = 65536            # samples per gradient step (the reference's "batch" is one robot's 300 samples)
    std_factor: float = 1.0
    use_cuda_graph: bool = True       # capture one rollout step (policy + transform + env step + reset policy + stores)
    graph_update: bool = True         # capture the update iterations in CUDA graphs (single-GPU; multi-GPU runs them eagerly)
    use_fused_policy: bool = True     # rollout forward through csrc/uaisa_policy.cu (K6a features + K6b tcgen05 heads, TF32 inputs)
    multi_cap_frac: float = 1.0 / 64  # static capacity of the multi-row GRU path in the captured step (fraction of A)


def gae(rew, val, cut, gamma, lam):
    """uaisa_gae: [T,A] float32 CUDA tensors -> (adv, ret)."""
    T, A = rew.shape
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _capi.check(_capi.lib().uaisa_gae(p(rew), p(val), p(cut), T, A, float(gamma), float(lam), p(adv), p(ret),
                                      C.c_void_p(torch.cuda.current_stream(rew.device).cuda_stream)))
    return adv, ret


class PPOTrainer:
    def __init__(self, env: VecMdin, ac: Optional[rnn_ac] = None, cfg: Optional[PPOConfig] = None, seed: int = 7):
        self.env, self.cfg = env, cfg or PPOConfig()
        self.dev = env.device
        torch.manual_seed(seed)        # multi_ppo.py:103-105
        self.ac = (ac or rnn_ac()).to(self.dev)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.world > 1:             # same parameters everywhere (the reference's intended bcast, multi_ppo.py:179-181)
            for p in self.ac.parameters():
                dist.broadcast(p.data, 0)
            torch.manual_seed(seed + 1000 * dist.get_rank())   # same weights, different exploration noise per shard
        cap_ok = self.cfg.graph_update and self.world == 1
        self.pi_opt = torch.optim.Adam(self.ac.pi.parameters(), lr=self.cfg.pi_lr, capturable=cap_ok)     # multi_ppo.py:115-116
        self.vf_opt = torch.optim.Adam(self.ac.v.parameters(), lr=self.cfg.vf_lr, capturable=cap_ok)
        self._upd = None              # captured update graphs
        T, A, D = self.cfg.steps_per_epoch, env.A, env.obs_dim
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.buf = {"obs": torch.empty((T, A, D), **f32), "count": torch.empty((T, A), dtype=torch.uint8, device=self.dev),
                    "act": torch.empty((T, A, 3), **f32), "rew": torch.empty((T, A), **f32), "val": torch.empty((T, A), **f32),
                    "logp": torch.empty((T, A), **f32), "cut": torch.empty((T, A), dtype=torch.uint8, device=self.dev)}
        self.action = torch.empty((A, 3), **f32)
        self.env.reset()
        self.stats = {}
        self._graph = None
        self._graph_fast = None       # same step without the multi-row GRU path (taken when no agent has >= 2 neighbour rows)
        self._flag_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.fused = None
        if self.cfg.use_fused_policy:
            from .fused import FusedPolicy
            self.fused = FusedPolicy(self.ac, self.dev)
        self._t = torch.zeros(1, dtype=torch.long, device=self.dev)           # device-side step counter (graph-safe buffer index)
        self._ndone = torch.zeros((), dtype=torch.int64, device=self.dev)
        self._last = torch.zeros((), dtype=torch.int64, device=self.dev)

    # ------------------------------------------------------------------------------------------------ rollout
    @torch.no_grad()
    def _one_step(self, epoch_ended: bool, cap: int):
        """One rollout step writing into buf[..][t] through the device-side counter (multi_ppo.py:193-281)."""
        env, cfg, b, t = self.env, self.cfg, self.buf, self._t
        b["obs"].index_copy_(0, t, env.obs.unsqueeze(0)); b["count"].index_copy_(0, t, env.count.unsqueeze(0))      # :217-219 store(obs, ...)
        stepper = self.fused if self.fused is not None else self.ac
        a_inc, v, logp = stepper.step(env.obs, env.count, cfg.std_factor, multi_cap=cap)                          # :196
        _, a_r = env.action_transform(a_inc, self.action)                                                          # :197-208
        env.step(self.action)                                                                                      # :210
        b["act"].index_copy_(0, t, a_r.unsqueeze(0)); b["rew"].index_copy_(0, t, env.reward.unsqueeze(0))
        b["val"].index_copy_(0, t, v.unsqueeze(0)); b["logp"].index_copy_(0, t, logp.unsqueeze(0))
        self._ndone += env.done.sum()
        env.autoreset(max_ep_len=cfg.max_ep_len, epoch_ended=epoch_ended)                                          # :226-281 (+ env_observation)
        b["cut"].index_copy_(0, t, env.path_cut.unsqueeze(0))
        t += 1
        self._flag_dev.copy_((env.count > 1).any().to(torch.int32).reshape(1))      # does the NEXT step need the multi-row path?
        self._flag_host.copy_(self._flag_dev, non_blocking=True)

    @torch.no_grad()
    def rollout(self):
        """steps_per_epoch env steps of every env (multi_ppo.py:183-281).  The T-1 ordinary steps replay one CUDA graph
        (no host synchronisation inside: the policy's multi-row path runs at a static capacity); the last step of the
        epoch (full reset, multi_ppo.py:244-264) runs eagerly."""
        cfg = self.cfg
        T = cfg.steps_per_epoch
        self._t.zero_(); self._ndone.zero_()
        rd = self.ac.pi.rnn_reader
        if getattr(rd, "overflow", None) is None:
            rd.overflow = torch.zeros((), dtype=torch.int64, device=self.dev)
        rd.overflow.zero_()
        cap = max(64, int(self.env.A * cfg.multi_cap_frac)) if cfg.use_cuda_graph else 0
        if cfg.use_cuda_graph and self._graph is None and T > 1:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):            # warm-up on a side stream (allocator, cuBLAS workspaces), then capture
                self._one_step(False, cap); self._one_step(False, -1)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            self.env.reset()
            g, gf = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(False, cap)
            with torch.cuda.graph(gf, pool=g.pool()):
                self._one_step(False, -1)
            self._graph, self._graph_fast = g, gf
            self._t.zero_(); self._ndone.zero_(); rd.overflow.zero_()
            self.env.reset()
        self._flag_host[0] = int((self.env.count > 1).any())
        for t in range(T - 1):
            if self._graph is not None:
                torch.cuda.current_stream(self.dev).synchronize()        # the flag of the previous step has landed (one small bubble per step)
                (self._graph if int(self._flag_host[0]) else self._graph_fast).replay()
            else:
                self._one_step(False, cap)
        torch.cuda.current_stream(self.dev).synchronize()
        self._one_step(True, cap if int(self._flag_host[0]) or not cfg.use_cuda_graph else -1)
        self.stats["collisions"] = self._ndone
        if cap and int(rd.overflow) > 0:
            raise RuntimeError("%d agent-steps had >= 2 neighbour rows beyond the static GRU capacity (%d); raise PPOConfig.multi_cap_frac" % (int(rd.overflow), cap))
        return self.buf

    # ------------------------------------------------------------------------------------------------ update
    def _allreduce_grads(self, params, extra: Optional[torch.Tensor] = None):
        if self.world == 1:
            return extra
        flat = [p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel(), device=self.dev) for p in params]
        bucket = torch.cat(flat + ([extra.reshape(-1)] if extra is not None else []))
        dist.all_reduce(bucket)
        bucket /= self.world
        off = 0
        for p in params:
            n = p.numel()
            if p.grad is not None:
                p.grad.copy_(bucket[off:off + n].view_as(p.grad))
            off += n
        return bucket[off:] if extra is not None else None

    def _build_update_graphs(self, n, mb):
        """Capture one policy iteration (forward + loss + KL + backward + clip | optimiser step) and one value iteration
        (forward + loss + backward + step) over static minibatch buffers (multi_ppo.py:355-376)."""
        cfg, dev = self.cfg, self.dev
        u = {"idx": torch.zeros(mb, dtype=torch.long, device=dev), "kl": torch.zeros((), device=dev), "loss_pi": torch.zeros((), device=dev),
             "loss_v": torch.zeros((), device=dev), "n": n, "mb": mb, "cap": -1}   # -1: no sample has >= 2 neighbour rows (checked per epoch)
        rd = self.ac.pi.rnn_reader
        pi_params = list(self.ac.pi.parameters())

        def pi_fb():
            u["idx"].random_(0, n)
            idx = u["idx"]
            w = self._ok[idx].float()
            dist_, _, logp = self.ac.evaluate(self._obs[idx], self._cnt[idx], self._act[idx], multi_cap=u["cap"], heads="pi")
            ratio = torch.exp(logp - self._logp_old[idx])
            a_ = torch.nan_to_num(self._adv[idx], 0.0, 0.0, 0.0)
            clip_adv = torch.clamp(ratio, 1 - cfg.clip_ratio, 1 + cfg.clip_ratio) * a_
            loss = -(torch.min(ratio * a_, clip_adv) * w).sum() / w.sum().clamp(min=1)
            u["kl"].copy_((((self._logp_old[idx] - logp) * w).sum() / w.sum().clamp(min=1)).detach())
            u["loss_pi"].copy_(loss.detach())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(pi_params, max_norm=cfg.max_grad_norm)

        def v_all():
            u["idx"].random_(0, n)
            idx = u["idx"]
            w = self._ok[idx].float()
            _, v, _ = self.ac.evaluate(self._obs[idx], self._cnt[idx], None, multi_cap=u["cap"], heads="v")
            loss = (((v - torch.nan_to_num(self._ret[idx], 0.0, 0.0, 0.0)) ** 2) * w).sum() / w.sum().clamp(min=1)
            u["loss_v"].copy_(loss.detach())
            loss.backward()
            self.vf_opt.step()

        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):                       # warm-up (optimizer state, cuBLAS workspaces) on a side stream
            for _ in range(2):
                self.pi_opt.zero_grad(set_to_none=True); pi_fb(); self.pi_opt.step()
                self.vf_opt.zero_grad(set_to_none=True); v_all()
        torch.cuda.current_stream(dev).wait_stream(s)
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self.pi_opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g1):
            pi_fb()
        with torch.cuda.graph(g2, pool=g1.pool()):
            self.pi_opt.step()
        self.vf_opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g3, pool=g1.pool()):
            v_all()
        u.update(g_pi_fb=g1, g_pi_step=g2, g_v=g3)
        rd.overflow.zero_()
        return u

    def _update_graphed(self, n, mb):
        cfg = self.cfg
        if self._upd is None or self._upd["n"] != n or self._upd["mb"] != mb:
            # the warm-up iterations inside the builder are real (tiny-lr) optimiser steps on this epoch's data
            self._upd = self._build_update_graphs(n, mb)
        u = self._upd
        pi_iters = 0
        for i in range(cfg.train_pi_iters):
            u["g_pi_fb"].replay()
            if float(u["kl"].detach()) > cfg.target_kl:  # multi_ppo.py:362-364 (the step is NOT applied)
                break
            u["g_pi_step"].replay()
            pi_iters += 1
        for i in range(cfg.train_v_iters):
            u["g_v"].replay()
        rd = self.ac.pi.rnn_reader
        if int(rd.overflow) > 0:
            raise RuntimeError("update minibatches held more multi-row samples than the static GRU capacity (%d)" % u["cap"])
        return pi_iters, float(u["loss_v"]), float(u["kl"])

    def update(self):
        """PPO-clip update (multi_ppo.py:341-404) on the whole buffer."""
        cfg, b = self.cfg, self.buf
        adv, ret = gae(b["rew"], b["val"], b["cut"], cfg.gamma, cfg.lam)                            # finish_path, :68-77
        T, A = b["rew"].shape
        n = T * A
        obs, cnt = b["obs"].view(n, -1), b["count"].view(n)
        act, logp_old, adv, ret = b["act"].view(n, 3), b["logp"].view(n), adv.view(n), ret.view(n)
        ok = torch.isfinite(adv) & torch.isfinite(ret)        # the reference's inf/nan rewards (ir_gym.py:88) would poison Adam
        all_params = list(self.ac.parameters())
        mb = min(cfg.minibatch, n)
        if cfg.graph_update and self.world == 1 and not bool((cnt > 1).any()):   # multi-row samples present -> eager path below (exact, dynamic shapes)
            rd = self.ac.pi.rnn_reader
            if getattr(rd, "overflow", None) is None:
                rd.overflow = torch.zeros((), dtype=torch.int64, device=self.dev)
            rd.overflow.zero_()
            if getattr(self, "_obs", None) is None:      # static views/buffers the captured graphs read (same storage every epoch)
                self._obs, self._cnt, self._act, self._logp_old = obs, cnt, act, logp_old
                self._adv, self._ret, self._ok = torch.empty_like(adv), torch.empty_like(ret), torch.empty_like(ok)
            self._adv.copy_(adv); self._ret.copy_(ret); self._ok.copy_(ok)
            pi_iters, lv, klv = self._update_graphed(n, mb)
            if self.fused is not None:
                self.fused.refresh()
            self.stats.update(pi_iters=pi_iters, loss_v=lv, kl=klv, mean_ret=float(torch.nan_to_num(ret, 0.0, 0.0, 0.0).mean()))
            return self.stats
        pi_iters = 0
        for i in range(cfg.train_pi_iters):                                                        # :355-368
            idx = torch.randint(0, n, (mb,), device=self.dev) if mb < n else torch.arange(n, device=self.dev)
            w = ok[idx].float()
            self.pi_opt.zero_grad(set_to_none=True)
            dist_, _, logp = self.ac.evaluate(obs[idx], cnt[idx], act[idx], heads="pi")
            ratio = torch.exp(logp - logp_old[idx])
            a_ = torch.nan_to_num(adv[idx], 0.0, 0.0, 0.0)
            clip_adv = torch.clamp(ratio, 1 - cfg.clip_ratio, 1 + cfg.clip_ratio) * a_
            loss_pi = -(torch.min(ratio * a_, clip_adv) * w).sum() / w.sum().clamp(min=1)          # :393-397
            kl = ((logp_old[idx] - logp) * w).sum() / w.sum().clamp(min=1)                          # :399
            if self.world > 1:
                kl_all = kl.detach().clone(); dist.all_reduce(kl_all); kl_all /= self.world
            else:
                kl_all = kl.detach()
            if float(kl_all) > cfg.target_kl:                                                       # :362-364
                break
            loss_pi.backward()
            self._allreduce_grads(all_params)
            torch.nn.utils.clip_grad_norm_(all_params, max_norm=cfg.max_grad_norm)                  # :367 (all ac params)
            self.pi_opt.step()
            pi_iters += 1
        for i in range(cfg.train_v_iters):                                                         # :371-376
            idx = torch.randint(0, n, (mb,), device=self.dev) if mb < n else torch.arange(n, device=self.dev)
            w = ok[idx].float()
            self.vf_opt.zero_grad(set_to_none=True)
            _, v, _ = self.ac.evaluate(obs[idx], cnt[idx], None, heads="v")
            loss_v = (((v - torch.nan_to_num(ret[idx], 0.0, 0.0, 0.0)) ** 2) * w).sum() / w.sum().clamp(min=1)   # :379-383
            loss_v.backward()
            self._allreduce_grads(list(self.ac.v.parameters()))
            self.vf_opt.step()
        if self.fused is not None:
            self.fused.refresh()          # repack the updated weights in place (addresses stay fixed: the captured rollout step stays valid)
        self.stats.update(pi_iters=pi_iters, loss_v=float(loss_v.detach()), kl=float(kl_all), mean_ret=float(torch.nan_to_num(ret, 0.0, 0.0, 0.0).mean()))
        return self.stats

    def train_epoch(self):
        t0 = time.perf_counter()
        self.rollout()
        torch.cuda.synchronize(self.dev)
        t1 = time.perf_counter()
        self.update()
        torch.cuda.synchronize(self.dev)
        t2 = time.perf_counter()
        self.stats.update(rollout_s=t1 - t0, update_s=t2 - t1)
        return self.stats

    def save_model(self, path):
        """Same container as multi_ppo.save_model (multi_ppo.py:406-420): {'model_state', 'pi_optimizer', 'vf_optimizer'}."""
        torch.save(dict(model_state=self.ac.state_dict(), pi_optimizer=self.pi_opt.state_dict(), vf_optimizer=self.vf_opt.state_dict()), path)

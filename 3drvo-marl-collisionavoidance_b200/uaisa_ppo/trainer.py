"""Batched PPO rollout + update around VecMdin: the data-parallel form of ``multi_ppo.training_loop``
(train/policy/multi_ppo.py:166-339).

What is kept from the reference (cited per line below): the per-step sequence (policy step -> 2-dp action
transform -> env step -> store -> trainer reset policy -> re-observe), GAE(gamma .99, lambda .97) with last_val = 0
and paths cut exactly where ``finish_path`` is called, NO advantage normalisation, the PPO-clip loss, the approx-KL
early stop, grad-norm clipping at 2.0 over ALL actor-critic parameters for the policy step, separate Adam optimisers for
pi and v that both touch the shared reader (each with its own moments for it), and the hyper-parameters of train_process.py:53-63.

What changes (documented deviation, SURVEY.md section 7 "PPO semantics vs batching"): the reference loops over
robots and does 50+50 FULL-BATCH iterations on each robot's own 300 samples, i.e. every sample is visited 50+50 times.
Here every agent of every env feeds ONE buffer of n = T x A samples and an iteration is one minibatch of ``minibatch``
samples (over all ranks) taken from a fresh pseudo-random permutation of the buffer (without replacement).  With the defaults (50+50 iterations of
65 536) a sample is visited 50 * 65536 / n times per head -- 0.33 at n = 9.8 M -- so samples/s here is NOT a like-for-like
ratio to the reference's; ``PPOConfig.passes`` = k makes the iteration count cover the buffer k times instead, and bench.py
prints ``passes_per_sample`` next to every PPO number.

Data-parallel update: parameters, gradients and Adam moments are flat fp32 buckets (``FlatParams``); with world_size > 1
each iteration all-reduces ONE slice of the gradient bucket (NCCL over NVLink; the KL scalar rides in the same bucket), and
the early-stop rule, clip_grad_norm_ and Adam run on the device (csrc/uaisa_train.cu) -- an iteration has no host
synchronisation and is replayed as a CUDA graph (the all-reduce is a node of the graph).
"""
from __future__ import annotations

import ctypes as C
import math
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
    minibatch: int = 65536            # samples per gradient step over ALL ranks (each rank takes minibatch / world_size; the reference's "batch" is one robot's 300 samples)
    passes: float = 0.0               # > 0: ceil(passes * n / minibatch) iterations per head (cover the buffer `passes` times) instead of train_*_iters
    std_factor: float = 1.0
    use_cuda_graph: bool = True       # capture one rollout step (policy + transform + env step + reset policy + stores)
    graph_update: bool = True         # replay the update iterations as CUDA graphs (also with world_size > 1: the all-reduce is captured)
    use_fused_policy: bool = True     # rollout forward through the fused tcgen05 kernels (TF32 operands, fp32 accumulation)
    fused_kernel: str = "full"        # "full": csrc/uaisa_policy_fwd.cu (one kernel, in-kernel recurrence); "split": round-1 K6a + K6b pair
    multi_cap_frac: float = 1.0 / 64  # static capacity of the multi-row GRU path in the captured step (fraction of A)
    unroll_rollout: bool = True       # capture ALL steps of an epoch as one graph whose kernels write straight into the epoch buffers (no per-step copies of obs / count / reward / cut)
    native_update: bool = True        # update iterations without torch autograd: reader / MLP / loss forward and backward as a fixed sequence of our own kernels writing
                                      # straight into the flat gradient bucket (needs tc_mlp and minibatches without multi-row samples; otherwise the autograd path)
    tc_mlp: bool = True               # update: actor / critic MLP forward + backward on tcgen05 (csrc/uaisa_mlp.cu, TF32 operands) instead of torch autograd / cuBLAS fp32


def gae(rew, val, cut, gamma, lam):
    """uaisa_gae: [T,A] float32 CUDA tensors -> (adv, ret)."""
    T, A = rew.shape
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _capi.check(_capi.lib().uaisa_gae(p(rew), p(val), p(cut), T, A, float(gamma), float(lam), p(adv), p(ret),
                                      C.c_void_p(torch.cuda.current_stream(rew.device).cuda_stream)))
    return adv, ret


class FlatAdam:
    """Adam moments + device-side step counter of one optimiser over the flat bucket (torch.optim.Adam defaults, multi_ppo.py:115-116)."""

    def __init__(self, n, lr, dev, betas=(0.9, 0.999), eps=1e-8):
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)


class FlatParams:
    """Every actor-critic parameter as a view into ONE fp32 bucket ``p`` (and its gradient into ``g``), laid out
    [ v.v_net | shared rnn_reader | pi.net_out, pi.log_std | kl, loss, 2 spare ] so that the value optimiser's parameters
    ([0, v_hi)) and the policy optimiser's ([pi_lo, n)) are each one contiguous range -- one all-reduce per iteration."""

    def __init__(self, ac: rnn_ac, dev):
        v_only = list(ac.v.v_net.parameters())
        reader = list(ac.pi.rnn_reader.parameters())
        pi_only = list(ac.pi.net_out.parameters()) + [ac.pi.log_std]
        self.params = v_only + reader + pi_only
        if {id(q) for q in self.params} != {id(q) for q in ac.parameters()} or len(self.params) != len(list(ac.parameters())):
            raise ValueError("FlatParams does not cover the module's parameters (unexpected rnn_ac layout)")
        n_v, n_r, n_pi = (sum(q.numel() for q in grp) for grp in (v_only, reader, pi_only))
        self.n, self.pi_lo, self.v_hi = n_v + n_r + n_pi, n_v, n_v + n_r
        self.p = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.g = torch.zeros(self.n + 4, dtype=torch.float32, device=dev)
        off = 0
        for q in self.params:
            k = q.numel()
            self.p[off:off + k].copy_(q.data.reshape(-1))
            q.data = self.p[off:off + k].view_as(q)
            q.grad = self.g[off:off + k].view_as(q)
            off += k
        self.L = _capi.lib()
        self.partials = torch.zeros(int(self.L.uaisa_opt_partials_len()), dtype=torch.float64, device=dev)
        self.dev = dev

    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def prepare(self, opt: FlatAdam, want_norm: bool, kl=None, target_kl: float = 0.0, ctl=None):
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        rc = self.L.uaisa_opt_prepare(p(self.g), self.n, int(want_norm), p(self.partials), p(kl), float(target_kl), p(ctl), p(opt.step), self._s())
        if rc:
            raise _capi.UaisaError("uaisa_opt_prepare failed (%d)" % rc)

    def clip_adam(self, opt: FlatAdam, r_off: int, r_cnt: int, max_norm: float, ctl=None):
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        rc = self.L.uaisa_clip_adam(p(self.p), p(self.g), p(opt.m), p(opt.v), self.n, int(r_off), int(r_cnt), p(self.partials), float(max_norm), opt.lr,
                                    opt.betas[0], opt.betas[1], opt.eps, p(opt.step), p(ctl), self._s())
        if rc:
            raise _capi.UaisaError("uaisa_clip_adam failed (%d)" % rc)

    def torch_state_dict(self, opt: FlatAdam, lo: int, hi: int):
        """The optimiser's state in torch.optim.Adam's state_dict layout (multi_ppo.save_model stores one, multi_ppo.py:406-420)."""
        state, ids, off, k = {}, [], 0, 0
        for q in self.params:
            n = q.numel()
            if off >= lo and off + n <= hi:
                state[k] = {"step": opt.step.float().cpu().reshape(()), "exp_avg": opt.m[off:off + n].view_as(q).clone(), "exp_avg_sq": opt.v[off:off + n].view_as(q).clone()}
                ids.append(k); k += 1
            off += n
        group = {"lr": opt.lr, "betas": opt.betas, "eps": opt.eps, "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "params": ids}
        return {"state": state, "param_groups": [group]}


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
        self.ac.pi.tc_mlp = self.ac.v.tc_mlp = bool(self.cfg.tc_mlp)
        self.flat = FlatParams(self.ac, self.dev)
        self.pi_opt = FlatAdam(self.flat.n, self.cfg.pi_lr, self.dev)        # Adam(ac.pi.parameters()): reader + net_out + log_std = [pi_lo, n)
        self.vf_opt = FlatAdam(self.flat.n, self.cfg.vf_lr, self.dev)        # Adam(ac.v.parameters()):  v_net + reader = [0, v_hi)   (multi_ppo.py:115-116)
        self._ctl = torch.zeros(2, dtype=torch.int32, device=self.dev)       # device side of the KL early stop: {stopped latch, iterations applied}
        self._upd = None              # captured update graphs
        self.policy_path = "not built yet"
        self.update_path = "not run yet"
        T, A, D = self.cfg.steps_per_epoch, env.A, env.obs_dim
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.buf = {"obs": torch.empty((T, A, D), **f32), "count": torch.empty((T, A), dtype=torch.uint8, device=self.dev),
                    "act": torch.empty((T, A, 3), **f32), "rew": torch.empty((T, A), **f32), "val": torch.empty((T, A), **f32),
                    "logp": torch.empty((T, A), **f32), "cut": torch.empty((T, A), dtype=torch.uint8, device=self.dev)}
        self.action = torch.empty((A, 3), **f32)
        self.env.reset()
        self.stats = {}
        self._graph = None
        self._graph_unrolled = None   # the whole epoch as one graph (unroll_rollout)
        self._graph_fast = None       # same step without the multi-row GRU path (taken when no agent has >= 2 neighbour rows)
        self._flag_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.fused = None
        if self.cfg.use_fused_policy:
            from .fused import FusedPolicy
            self.fused = FusedPolicy(self.ac, self.dev, kernel=self.cfg.fused_kernel)
        self._single_graph = self.fused is not None and self.fused.kernel == "full"     # one captured step, no per-step flag read
        self.policy_path = ("one fused kernel: masked bi-GRU + LayerNorm + actor/critic MLPs on tcgen05 (TF32), bulk-copied pre-swizzled weight tiles" if self._single_graph
                            else "split kernels: K6a step-0 features + K6b tcgen05 heads, multi-row agents through torch" if self.fused is not None else "torch module (cuBLAS)")
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
        env.step_autoreset(self.action, max_ep_len=cfg.max_ep_len, epoch_ended=epoch_ended)                        # :210 + :226-281 (+ env_observation): one launch
        b["act"].index_copy_(0, t, a_r.unsqueeze(0)); b["rew"].index_copy_(0, t, env.reward.unsqueeze(0))
        b["val"].index_copy_(0, t, v.unsqueeze(0)); b["logp"].index_copy_(0, t, logp.unsqueeze(0))
        self._ndone += env.done.sum()
        b["cut"].index_copy_(0, t, env.path_cut.unsqueeze(0))
        t += 1
        if not self._single_graph:
            self._flag_dev.copy_((env.count > 1).any().to(torch.int32).reshape(1))      # does the NEXT step need the multi-row path?
            self._flag_host.copy_(self._flag_dev, non_blocking=True)

    @torch.no_grad()
    def _one_step_direct(self, t: int, last: bool):
        """Step t of an epoch with a STATIC buffer index (unrolled capture): the policy reads the observation where the previous step's env kernel
        wrote it (buf[obs][t]), the env kernel writes reward / path cut into buf[..][t] and the next observation into buf[obs][t+1] -- no copies
        of the big tensors.  The last step (full reset, multi_ppo.py:244-264) leaves the next epoch's first observation in the env's own buffers."""
        env, cfg, b = self.env, self.cfg, self.buf
        a_inc, _, _ = self.fused.step(b["obs"][t], b["count"][t], cfg.std_factor, out_v=b["val"][t], out_logp=b["logp"][t])   # :196; values and log-probabilities land in the buffer
        env.action_transform(a_inc, self.action, b["act"][t])                                                      # :197-208
        out = {"reward": b["rew"][t], "path_cut": b["cut"][t]}
        if not last:
            out["obs"] = b["obs"][t + 1]; out["count"] = b["count"][t + 1]
        env.step_autoreset(self.action, max_ep_len=cfg.max_ep_len, epoch_ended=last, out=out)                      # :210 + :226-281
        self._ndone += env.done.sum()

    @torch.no_grad()
    def _rollout_unrolled(self):
        """The whole epoch as ONE CUDA graph of T unrolled steps (captured once; replay = one launch per epoch)."""
        T = self.cfg.steps_per_epoch
        self._ndone.zero_()
        if self._graph_unrolled is None:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):            # warm-up on a side stream, then start again from a reset
                self.buf["obs"][0].copy_(self.env.obs); self.buf["count"][0].copy_(self.env.count)
                self._one_step_direct(0, False); self._one_step_direct(T - 1, True)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            self.env.reset()
            self._ndone.zero_()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.buf["obs"][0].copy_(self.env.obs); self.buf["count"][0].copy_(self.env.count)
                for t in range(T):
                    self._one_step_direct(t, t == T - 1)
            self._graph_unrolled = g
            self._ndone.zero_()
            self.env.reset()
        self._graph_unrolled.replay()
        self.stats["collisions"] = self._ndone
        return self.buf

    @torch.no_grad()
    def rollout(self):
        """steps_per_epoch env steps of every env (multi_ppo.py:183-281).  The T-1 ordinary steps replay ONE CUDA graph with no host
        synchronisation in between (the fused forward handles every neighbour-row count inside the kernel); the last step of the epoch
        (full reset, multi_ppo.py:244-264) runs eagerly.  With the torch policy / the split kernels (`use_fused_policy=False` or
        FusedPolicy(kernel="split")) two graphs are kept -- with / without the multi-row GRU path at a static capacity -- and a 4-byte
        flag read per step picks one."""
        cfg = self.cfg
        T = cfg.steps_per_epoch
        if cfg.use_cuda_graph and cfg.unroll_rollout and self._single_graph and 1 < T <= 2048:
            return self._rollout_unrolled()
        self._t.zero_(); self._ndone.zero_()
        rd = self.ac.pi.rnn_reader
        if getattr(rd, "overflow", None) is None:
            rd.overflow = torch.zeros((), dtype=torch.int64, device=self.dev)
        rd.overflow.zero_()
        single = self._single_graph
        cap = max(64, int(self.env.A * cfg.multi_cap_frac)) if (cfg.use_cuda_graph and not single) else 0
        if cfg.use_cuda_graph and self._graph is None and T > 1:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):            # warm-up on a side stream (allocator, cuBLAS workspaces), then capture
                self._one_step(False, cap); self._one_step(False, -1)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            self.env.reset()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(False, cap)
            self._graph = g
            if not single:
                gf = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf, pool=g.pool()):
                    self._one_step(False, -1)
                self._graph_fast = gf
            self._t.zero_(); self._ndone.zero_(); rd.overflow.zero_()
            self.env.reset()
        if not single:
            self._flag_host[0] = int((self.env.count > 1).any())
        for t in range(T - 1):
            if self._graph is not None:
                if single:
                    self._graph.replay()
                else:
                    torch.cuda.current_stream(self.dev).synchronize()        # the flag of the previous step has landed (one small bubble per step)
                    (self._graph if int(self._flag_host[0]) else self._graph_fast).replay()
            else:
                self._one_step(False, cap)
        if not single:
            torch.cuda.current_stream(self.dev).synchronize()
        self._one_step(True, cap if (single or int(self._flag_host[0]) or not cfg.use_cuda_graph) else -1)
        self.stats["collisions"] = self._ndone
        if cap and int(rd.overflow) > 0:
            raise RuntimeError("%d agent-steps had >= 2 neighbour rows beyond the static GRU capacity (%d); raise PPOConfig.multi_cap_frac" % (int(rd.overflow), cap))
        return self.buf

    # ------------------------------------------------------------------------------------------------ update
    def _next_idx(self):
        """Next minibatch of this epoch's permutation (sampling WITHOUT replacement; wraps around after a full pass).  The permutation is the
        affine bijection i -> (a i + b) mod n with a fresh random multiplier a coprime to n and offset b per epoch (device scalars, so the captured
        iteration stays valid): every sample is visited at most once per pass, and no n-element sort (torch.randperm: 20 ms at n = 7.9e7, 15 % of
        the update) is paid for the 8 % of the buffer an epoch's 50 + 50 iterations actually read."""
        u = self._u
        idx = (u["ab"][0] * ((u["off"] + u["ar"]) % u["n"]) + u["ab"][1]) % u["n"]
        u["off"] += u["mb"]
        return idx

    # ---- native iterations: no autograd tape, ~20 launches of our own kernels instead of ~75 (most of them element-wise torch kernels) ----------
    def _native(self):
        u, rd = self._u, self.ac.pi.rnn_reader
        return bool(self.cfg.native_update and self.cfg.tc_mlp and u["cap"] < 0 and rd.fused_step0 and self._obs.is_cuda)

    def _head_fb(self, head: str):
        """One head's forward, loss and backward on minibatch idx as a fixed kernel sequence (csrc/uaisa_train.cu, uaisa_mlp.cu): reader forward ->
        H1 -> H2, y -> loss + d loss / dy -> output layer backward -> dW2, dZ1, db1, dW1, dfeat -> reader backward.  Every gradient lands in the
        parameter's view of the flat bucket (zeroed by the caller); same arithmetic as the autograd path (tests/test_ppo_gpu.py)."""
        from .policy import _MlpTc
        cfg, F, u, W, L, dev = self.cfg, self.flat, self._u, self.world, self.flat.L, self.dev
        B, nout = u["mb"], (3 if head == "pi" else 1)
        ws = u.get("ws")
        if ws is None:
            e = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)  # noqa: E731
            ws = u["ws"] = {"feat": e(B, 268), "stats": e(B, 2), "hsum": e(B, 256), "H1": e(B, 256), "H2": e(B, 256), "y": e(B, 3), "dy": e(B, 3),
                            "dZ2": e(B, 256), "dZ1": e(B, 256), "dfeat": e(B, 268), "scal": torch.zeros(4, dtype=torch.float32, device=dev)}
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        s = F._s()
        rd = self.ac.pi.rnn_reader
        net = self.ac.pi.net_out if head == "pi" else self.ac.v.v_net
        (W1, b1), (W2, b2), (W3, b3) = ((net[i].weight, net[i].bias) for i in (0, 2, 4))
        idx = self._next_idx()
        rparams = rd._step0_params()
        pptr = (C.c_void_p * 8)(*[q.data_ptr() for q in rparams]); gptr = (C.c_void_p * 8)(*[q.grad.data_ptr() for q in rparams])

        def chk(rc, what):
            if rc:
                raise _capi.UaisaError("%s failed (%d)" % (what, rc))

        chk(L.uaisa_ppo_wsum(p(self._ok), p(idx), B, p(ws["scal"]), s), "uaisa_ppo_wsum")
        chk(L.uaisa_reader_fwd(p(self._obs), p(idx), B, int(self._obs.shape[1]), pptr, p(ws["feat"]), p(ws["stats"]), p(ws["hsum"]), s), "uaisa_reader_fwd")
        g = lambda *a, **k: _MlpTc._gemm(L, C, s, *a, **k)  # noqa: E731
        g(ws["feat"], 268, 0, W1, 268, 0, ws["H1"], 256, B, 256, 268, 1, 1, bias=b1)
        g(ws["H1"], 256, 0, W2, 256, 0, ws["H2"], 256, B, 256, 256, 1, 2, bias=b2, W3=W3, b3=b3, Y=ws["y"], nout=nout)
        if head == "pi":
            chk(L.uaisa_ppo_pi_loss(p(ws["y"]), p(self.ac.pi.log_std), float(cfg.std_factor), p(self._act), p(self._logp_old), p(self._adv), p(self._ok), p(idx), B,
                                    float(cfg.clip_ratio), 1.0 / W, p(ws["scal"]), p(ws["dy"]), p(self.ac.pi.log_std.grad), p(F.g[F.n:]), s), "uaisa_ppo_pi_loss")
        else:
            u["loss_v"].zero_()
            chk(L.uaisa_ppo_v_loss(p(ws["y"]), p(self._ret), p(self._ok), p(idx), B, 1.0 / W, p(ws["scal"]), p(ws["dy"]), p(u["loss_v"]), s), "uaisa_ppo_v_loss")
        chk(L.uaisa_mlp_out_bwd(p(ws["dy"]), p(ws["H2"]), p(W3), B, nout, p(ws["dZ2"]), p(W3.grad), p(b3.grad), p(b2.grad), s), "uaisa_mlp_out_bwd")
        sp, sp2 = max(1, min(_MlpTc.SPLITS, (B + 31) // 32)), max(1, min(_MlpTc.SPLITS_256, (B + 31) // 32))
        g(ws["dZ2"], 256, 1, ws["H1"], 256, 1, W2.grad, 256, 256, 256, B, sp2, 4)
        g(ws["dZ2"], 256, 0, W2, 256, 1, ws["dZ1"], 256, B, 256, 256, 1, 3, mask=ws["H1"], ldm=256)
        chk(L.uaisa_colsum(p(ws["dZ1"]), B, 256, p(b1.grad), s), "uaisa_colsum")
        g(ws["dZ1"], 256, 1, ws["feat"], 268, 1, W1.grad, 268, 256, 268, B, sp, 4)
        g(ws["dZ1"], 256, 0, W1, 268, 1, ws["dfeat"], 268, B, 268, 256, 1, 0)
        chk(L.uaisa_reader_bwd(p(self._obs), p(idx), B, int(self._obs.shape[1]), pptr, p(ws["dfeat"]), p(ws["stats"]), p(ws["hsum"]), gptr, s), "uaisa_reader_bwd")

    def _pi_fb(self):
        """Forward + PPO-clip loss + approx-KL + backward of one policy iteration (multi_ppo.py:385-404) into the flat gradient bucket, then
        ONE all-reduce of [reader | pi-only | kl, loss] across ranks.  No host round trip."""
        cfg, F, u, W = self.cfg, self.flat, self._u, self.world
        F.g[F.pi_lo:].zero_()                                      # pi_optimizer.zero_grad() (+ the kl / loss slots)
        if self._native():
            self._head_fb("pi")
            if W > 1:
                dist.all_reduce(F.g[F.pi_lo:])
            return
        idx = self._next_idx()
        w = self._ok[idx].float()
        wsum = w.sum().clamp(min=1)
        _, _, logp = self.ac.evaluate(self._obs, self._cnt, self._act[idx], multi_cap=u["cap"], heads="pi", index=idx)
        lo = self._logp_old[idx]
        ratio = torch.exp(logp - lo)
        a_ = torch.nan_to_num(self._adv[idx], 0.0, 0.0, 0.0)
        clip_adv = torch.clamp(ratio, 1 - cfg.clip_ratio, 1 + cfg.clip_ratio) * a_
        loss = -(torch.min(ratio * a_, clip_adv) * w).sum() / wsum                                # :393-397
        kl = ((lo - logp) * w).sum() / wsum                                                      # :399
        (loss / W).backward()
        F.g[F.n].copy_(kl.detach() / W); F.g[F.n + 1].copy_(loss.detach() / W)
        if W > 1:
            dist.all_reduce(F.g[F.pi_lo:])                        # sum of per-rank means / W = the global mean (equal minibatches per rank)

    def _pi_iter(self):
        """One policy iteration (multi_ppo.py:355-368): _pi_fb, then the device-side early-stop rule + clip_grad_norm_(ALL parameters, 2.0) +
        Adam on the policy optimiser's range (csrc/uaisa_train.cu)."""
        cfg, F = self.cfg, self.flat
        self._pi_fb()
        F.prepare(self.pi_opt, want_norm=True, kl=F.g[F.n:], target_kl=cfg.target_kl, ctl=self._ctl)
        F.clip_adam(self.pi_opt, F.pi_lo, F.n - F.pi_lo, max_norm=cfg.max_grad_norm, ctl=self._ctl)   # :367 clips all ac parameters' gradients

    def _v_fb(self):
        """Forward + MSE loss + backward of one value iteration (multi_ppo.py:379-383), all-reduce of [v_net | reader]."""
        F, u, W = self.flat, self._u, self.world
        F.g[:F.v_hi].zero_()                                       # vf_optimizer.zero_grad()
        if self._native():
            self._head_fb("v")
            if W > 1:
                dist.all_reduce(F.g[:F.v_hi])
            return
        idx = self._next_idx()
        w = self._ok[idx].float()
        _, v, _ = self.ac.evaluate(self._obs, self._cnt, None, multi_cap=u["cap"], heads="v", index=idx)
        loss = (((v - torch.nan_to_num(self._ret[idx], 0.0, 0.0, 0.0)) ** 2) * w).sum() / w.sum().clamp(min=1)
        (loss / W).backward()
        u["loss_v"].copy_(loss.detach())
        if W > 1:
            dist.all_reduce(F.g[:F.v_hi])

    def _v_iter(self):
        """One value iteration (multi_ppo.py:371-376): _v_fb, Adam on the value optimiser's range (no clipping, no early stop)."""
        F = self.flat
        self._v_fb()
        F.prepare(self.vf_opt, want_norm=False)
        F.clip_adam(self.vf_opt, 0, F.v_hi, max_norm=0.0)

    def _snapshot(self):
        F = self.flat
        return [t.clone() for t in (F.p, F.g, self.pi_opt.m, self.pi_opt.v, self.pi_opt.step, self.vf_opt.m, self.vf_opt.v, self.vf_opt.step, self._ctl, self._u["off"])]

    def _restore(self, snap):
        F = self.flat
        for dst, src in zip((F.p, F.g, self.pi_opt.m, self.pi_opt.v, self.pi_opt.step, self.vf_opt.m, self.vf_opt.v, self.vf_opt.step, self._ctl, self._u["off"]), snap):
            dst.copy_(src)

    def _build_update_graphs(self):
        """Capture one policy iteration and one value iteration as CUDA graphs (with world_size > 1 the NCCL all-reduce is a node of the
        graph).  The warm-up iterations that precede the capture run on a snapshot: parameters, gradients, Adam moments, step counters
        and the sampling offset are restored afterwards, so the first epoch applies exactly the counted iterations."""
        dev = self.dev
        snap = self._snapshot()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._pi_iter(); self._v_iter()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        kw = dict(capture_error_mode="thread_local") if self.world > 1 else {}
        with torch.cuda.graph(g1, **kw):
            self._pi_iter()
        with torch.cuda.graph(g2, pool=g1.pool(), **kw):
            self._v_iter()
        self._restore(snap)
        rd = self.ac.pi.rnn_reader
        if getattr(rd, "overflow", None) is not None:
            rd.overflow.zero_()
        return g1, g2

    def _begin_update(self, adv=None, ret=None):
        """GAE over the epoch's buffer (or injected advantages / returns: tests) into the static tensors every iteration reads; a fresh
        permutation; early-stop latch cleared.  -> (n, mb)"""
        cfg, b = self.cfg, self.buf
        if adv is None:
            adv, ret = gae(b["rew"], b["val"], b["cut"], cfg.gamma, cfg.lam)                        # finish_path, multi_ppo.py:68-77
        T, A = b["rew"].shape
        n = T * A
        mb = min(max(1, cfg.minibatch // self.world), n)      # this rank's share of the global minibatch
        if getattr(self, "_obs", None) is None:      # static views / buffers (same storage every epoch: graph-safe)
            self._obs, self._cnt = b["obs"].view(n, -1), b["count"].view(n)
            self._act, self._logp_old = b["act"].view(n, 3), b["logp"].view(n)
            self._adv, self._ret = torch.empty(n, dtype=torch.float32, device=self.dev), torch.empty(n, dtype=torch.float32, device=self.dev)
            self._ok = torch.empty(n, dtype=torch.bool, device=self.dev)
            self._u = {"n": n, "mb": mb, "cap": -1, "ab": torch.zeros(2, dtype=torch.long, device=self.dev), "off": torch.zeros((), dtype=torch.long, device=self.dev),
                       "ar": torch.arange(mb, device=self.dev), "loss_v": torch.zeros((), device=self.dev)}
        u = self._u
        self._adv.copy_(adv.reshape(n)); self._ret.copy_(ret.reshape(n))
        torch.logical_and(torch.isfinite(self._adv), torch.isfinite(self._ret), out=self._ok)   # the reference's inf/nan rewards (ir_gym.py:88) would poison Adam
        while True:                                            # multiplier coprime to n (host-side draw from torch's CPU generator: a handful of tries)
            a = int(torch.randint(1, max(n, 2), (1,)).item())
            if math.gcd(a, n) == 1:
                break
        u["ab"].copy_(torch.tensor([a, int(torch.randint(0, n, (1,)).item())], dtype=torch.long), non_blocking=True)
        u["off"].zero_(); self._ctl.zero_()
        return n, mb

    def update(self, adv=None, ret=None):
        """PPO-clip update (multi_ppo.py:341-404) on the whole buffer.

        Sampling (documented deviation, see the module docstring): every iteration takes the next `minibatch` samples of a fresh random
        permutation of the epoch's T x A samples (without replacement).  `PPOConfig.passes` > 0 makes the iteration count cover the
        buffer that many times (ceil(passes * n / minibatch) iterations per head) instead of train_pi_iters / train_v_iters."""
        cfg = self.cfg
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]      # device-side split of the update: prepare | policy iterations | value iterations
        ev[0].record()
        n, mb = self._begin_update(adv, ret)
        u = self._u
        n_pi = cfg.train_pi_iters if cfg.passes <= 0 else int(-(-cfg.passes * n // mb))
        n_v = cfg.train_v_iters if cfg.passes <= 0 else int(-(-cfg.passes * n // mb))
        multi = bool((self._cnt > 1).any())          # (one host sync per epoch)
        rd = self.ac.pi.rnn_reader
        if getattr(rd, "overflow", None) is None:
            rd.overflow = torch.zeros((), dtype=torch.int64, device=self.dev)
        rd.overflow.zero_()
        graphed = False
        if cfg.graph_update and not multi:           # multi-row samples present -> eager iterations (exact, dynamic shapes in the reader)
            u["cap"] = -1
            if self._upd is None:
                try:
                    self._upd = self._build_update_graphs()
                except Exception as ex:              # e.g. an NCCL build that cannot be captured: same iterations, launched eagerly
                    self._upd = False
                    self.stats["graph_update_error"] = repr(ex)
                    torch.cuda.synchronize(self.dev)
            graphed = bool(self._upd)
        else:
            u["cap"] = 0 if multi else -1
        ev[1].record()
        if graphed:
            g1, g2 = self._upd
            for _ in range(n_pi):
                g1.replay()
            ev[2].record()
            for _ in range(n_v):
                g2.replay()
        else:
            for _ in range(n_pi):
                self._pi_iter()
            ev[2].record()
            for _ in range(n_v):
                self._v_iter()
        ev[3].record()
        if self.fused is not None:
            self.fused.refresh()          # repack the updated weights in place (addresses stay fixed: the captured rollout step stays valid)
        ctl = self._ctl.tolist()          # the update's only host read: iterations applied / early-stop latch
        self.update_path = (("CUDA graph per iteration" if graphed else "eager iterations") + "; " + ("reader forward/backward kernels (uaisa_reader_fwd/bwd) + " if u["cap"] < 0 and self.ac.pi.rnn_reader.fused_step0 else "")
                            + ("MLP forward / dgrad / wgrad on tcgen05 (uaisa_mlp_gemm, TF32 operands)" if cfg.tc_mlp else "torch autograd GEMMs for the MLPs (cuBLAS fp32)")
                            + (" + loss kernels, no autograd tape" if self._native() else " + torch loss algebra / autograd")
                            + " + flat-bucket all-reduce + fused clip/Adam kernel")
        self.stats.update(update_prepare_ms=ev[0].elapsed_time(ev[1]), update_pi_ms=ev[1].elapsed_time(ev[2]), update_v_ms=ev[2].elapsed_time(ev[3]))
        self.stats.update(pi_iters=int(ctl[1]), early_stop=bool(ctl[0]), loss_v=float(u["loss_v"]), kl=float(self.flat.g[self.flat.n]), loss_pi=float(self.flat.g[self.flat.n + 1]),
                          passes_pi=ctl[1] * mb / n, passes_v=n_v * mb / n, mean_ret=float(torch.nan_to_num(self._ret, 0.0, 0.0, 0.0).mean()))
        return self.stats

    def train_epoch(self):
        nvtx = torch.cuda.nvtx
        t0 = time.perf_counter()
        nvtx.range_push("uaisa.rollout")
        self.rollout()
        torch.cuda.synchronize(self.dev)
        nvtx.range_pop()
        t1 = time.perf_counter()
        nvtx.range_push("uaisa.update")
        self.update()
        torch.cuda.synchronize(self.dev)
        nvtx.range_pop()
        t2 = time.perf_counter()
        self.stats.update(rollout_s=t1 - t0, update_s=t2 - t1)
        return self.stats

    def close(self):
        """Drop the captured CUDA graphs.  With world_size > 1 the update graphs hold NCCL kernels: release them BEFORE
        torch.distributed.destroy_process_group(), which otherwise waits on the communicator they still reference."""
        self._upd = None; self._graph = None; self._graph_fast = None; self._graph_unrolled = None
        torch.cuda.synchronize(self.dev)

    def save_model(self, path):
        """Same container as multi_ppo.save_model (multi_ppo.py:406-420): {'model_state', 'pi_optimizer', 'vf_optimizer'}."""
        F = self.flat
        torch.save(dict(model_state={k: v.detach().clone() for k, v in self.ac.state_dict().items()},
                        pi_optimizer=F.torch_state_dict(self.pi_opt, F.pi_lo, F.n), vf_optimizer=F.torch_state_dict(self.vf_opt, 0, F.v_hi)), path)

"""Shared GPU-vs-oracle lock-step harness (used by the -m gpu tests, smoke() and bench sanity)."""
from __future__ import annotations

import numpy as np
import torch

import _util  # noqa: F401
import oracle
from uaisa_env.vec_env import VecMdin, pack_state, unpack_state

STATE = ("pos", "vel", "yaw", "pitch", "wp_idx", "arrive_flag", "dest_flag", "real_len", "max_dev", "extra_len")


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def eq_or_nan(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return (a == b) | (np.isnan(a) & np.isnan(b))


class Stats:
    def __init__(self):
        self.n_steps = 0; self.n_agent_steps = 0
        self.flag_mismatch = 0; self.count_mismatch = 0; self.nbr_mismatch = 0
        self.obs_vals = 0; self.obs_inexact = 0; self.obs_maxrel = 0.0
        self.rew_vals = 0; self.rew_inexact = 0; self.rew_maxrel = 0.0
        self.state_vals = 0; self.state_inexact = 0; self.state_maxrel = 0.0
        self.rows = 0; self.dones = 0; self.finishes = 0; self.resets = 0

    def _rel(self, g, o):
        g = np.asarray(g, np.float64); o = np.asarray(o, np.float64)
        same = eq_or_nan(g, o)
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.where(same, 0.0, np.abs(g - o) / np.maximum(np.abs(o), 1e-3))
        rel = np.where(np.isnan(rel), np.inf, rel)
        return int((~same).sum()), float(rel.max()) if rel.size else 0.0

    def add_obs(self, g, o):
        n, r = self._rel(g, f32(o)); self.obs_vals += g.size; self.obs_inexact += n; self.obs_maxrel = max(self.obs_maxrel, r)

    def add_rew(self, g, o):
        n, r = self._rel(g, f32(o)); self.rew_vals += g.size; self.rew_inexact += n; self.rew_maxrel = max(self.rew_maxrel, r)

    def add_state(self, g, o):
        n, r = self._rel(g, f32(o)); self.state_vals += g.size; self.state_inexact += n; self.state_maxrel = max(self.state_maxrel, r)

    def summary(self):
        return {k: v for k, v in self.__dict__.items()}

    def assert_parity(self, tol=1e-5):
        """The contract of BASELINE.json: flags / neighbour indices bit-exact, floats within 1e-5 relative."""
        assert self.flag_mismatch == 0 and self.count_mismatch == 0 and self.nbr_mismatch == 0, self.summary()
        assert self.obs_maxrel <= tol and self.rew_maxrel <= tol and self.state_maxrel <= tol, self.summary()


def make_actions(rng, t, ob, st, block=25):
    """Mixed action stream from the oracle-side state (fp64 copies of the device's fp32 state)."""
    mode = (t // block) % 4
    des = ob.cal_des()
    if mode == 0:
        a = _util.training_like_actions(rng, st["vel"])
    elif mode == 1:
        a = _util.goal_seeking_actions(rng, des)
    else:
        a = _util.steering_actions(rng, des, st["vel"], st["yaw"], st["pitch"], speed=0.6 if mode == 2 else 1.0,
                                   noise=0.05 if mode == 2 else 0.2)
    return a.astype(np.float32)


def lockstep(scn, E, nm, steps, seed, env2scn=None, max_ep_len=60, epoch_len=0, block=25, threads_per_block=0,
             device="cuda", stats=None, neighbor_slices=0, **hyper):
    """Teacher-forced lock-step: every step the oracle is loaded with the device's fp32 state, both
    step on identical actions, every output is compared, then both apply the trainer reset policy."""
    stats = stats or Stats()
    env = VecMdin(scn, E, env2scenario=env2scn, neighbors_num=nm, device=device, threads_per_block=threads_per_block,
                  neighbor_slices=neighbor_slices, **hyper)
    ob = oracle.OracleBatch(scn, E, env2scn=env2scn, nm=nm, **hyper)
    rng = np.random.default_rng(seed)
    env.reset()
    ob.reset(); oobs, ocnt, onbr = ob.observe()
    stats.add_obs(env.obs.cpu().numpy(), oobs)
    assert np.array_equal(env.count.cpu().numpy(), ocnt.astype(np.uint8))
    ep_len = np.zeros(env.A, np.int32)
    for t in range(steps):
        st = unpack_state(env.get_state())
        ob.set_state(st)
        act = make_actions(rng, t, ob, st, block)
        act_t = torch.from_numpy(act).to(env.device)
        env.step(act_t)
        oobs, orew, odone, oinfo, ofin = ob.step(act.astype(np.float64))
        g = {k: getattr(env, k).cpu().numpy() for k in ("obs", "count", "nbr", "reward", "done", "info", "finish")}
        stats.n_steps += 1; stats.n_agent_steps += env.A
        stats.flag_mismatch += int((g["done"] != odone).sum() + (g["info"] != oinfo).sum() + (g["finish"] != ofin).sum())
        stats.count_mismatch += int((g["count"] != ob.count.astype(np.uint8)).sum())
        stats.nbr_mismatch += int((g["nbr"] != ob.nbr).sum())
        stats.add_obs(g["obs"], oobs); stats.add_rew(g["reward"], orew)
        stats.rows += int(ob.count.sum()); stats.dones += int(odone.sum()); stats.finishes += int(ofin.sum())
        gst = unpack_state(env.get_state()); ost = ob.get_state()
        for k in STATE:
            if k in ("wp_idx", "arrive_flag", "dest_flag"):
                stats.flag_mismatch += int((gst[k] != ost[k]).sum())
            else:
                stats.add_state(gst[k].astype(np.float32), ost[k])
        # trainer reset policy on both sides (device: uaisa_autoreset; oracle: uo_autoreset)
        epoch_ended = bool(epoch_len and (t + 1) % epoch_len == 0)
        ob.set_state(gst)   # both sides reset/observe from the same fp32 state
        env.autoreset(max_ep_len=max_ep_len, epoch_ended=epoch_ended)
        rmask, eobs, pcut = ob.autoreset(odone, ofin, ep_len, max_ep_len, epoch_ended)
        stats.flag_mismatch += int((env.reset_mask.cpu().numpy() != rmask).sum() + (env.path_cut.cpu().numpy() != pcut).sum())
        stats.resets += int(rmask.sum())
        if eobs.any():
            o2, c2, n2 = ob.observe(eobs)
            sel = np.repeat(eobs.astype(bool), env.N)
            stats.add_obs(env.obs.cpu().numpy()[sel], o2[sel])
            stats.count_mismatch += int((env.count.cpu().numpy()[sel] != c2[sel].astype(np.uint8)).sum())
            stats.nbr_mismatch += int((env.nbr.cpu().numpy()[sel] != n2[sel]).sum())
    env.close()
    return stats

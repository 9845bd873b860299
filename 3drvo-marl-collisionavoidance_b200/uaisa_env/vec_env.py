"""VecMdin: E environments x N drones stepped at once on one B200 through libuaisa_b200.so.

Batched counterpart of the reference's ``mdin`` facade (uaisa_env/drone_envs/mdin.py:6-49) and of the
``ir_gym`` methods the trainer reaches through it.  All tensors are CUDA tensors;
agent a = env * N + drone.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi
from .scenario import ScenarioSet


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class VecMdin:
    def __init__(self, scenarios: ScenarioSet, n_envs: int, env2scenario=None, neighbors_num: int = 10,
                 device="cuda", env_train: bool = True, acceler: float = 0.5, threads_per_block: int = 0,
                 count_pairs: bool = False, **hyper):
        if not torch.cuda.is_available():
            raise _capi.UaisaError("VecMdin needs a CUDA device (B200); there is no CPU fallback")
        self.L = _capi.lib()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        scenarios.validate()
        self.scn = scenarios
        self.E, self.N, self.nm = int(n_envs), scenarios.N, int(neighbors_num)
        self.A = self.E * self.N
        self.obs_dim = 12 + 9 * self.nm
        self.acceler = acceler            # ir_gym.py:34, read by multi_ppo.py:203
        cfg = _capi.default_config()
        cfg.n_envs, cfg.n_drones, cfg.max_waypoints = self.E, self.N, scenarios.W
        cfg.max_buildings, cfg.n_scenarios, cfg.nm = scenarios.M, scenarios.S, self.nm
        cfg.env_train, cfg.threads_per_block, cfg.count_pairs = int(env_train), int(threads_per_block), int(count_pairs)
        for k, v in hyper.items():
            if not hasattr(cfg, k):
                raise TypeError("unknown hyper-parameter %r" % k)
            setattr(cfg, k, v)
        self.cfg = cfg
        if env2scenario is None:
            env2scenario = np.arange(self.E, dtype=np.int32) % scenarios.S
        self.env2scenario = np.ascontiguousarray(env2scenario, np.int32)
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(self.L.uaisa_create(C.byref(cfg), C.byref(self.h)))
            arrs = [np.ascontiguousarray(scenarios.waypoints, np.float64), np.ascontiguousarray(scenarios.n_points, np.int32),
                    np.ascontiguousarray(scenarios.radius, np.float64), np.ascontiguousarray(scenarios.priority, np.float64),
                    np.ascontiguousarray(scenarios.buildings, np.float64), np.ascontiguousarray(scenarios.n_buildings, np.int32),
                    np.ascontiguousarray(scenarios.map_size, np.float64), self.env2scenario]
            _capi.check(self.L.uaisa_set_scenarios(self.h, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
        dev, A, nm = self.device, self.A, self.nm
        self.obs = torch.zeros((A, self.obs_dim), dtype=torch.float32, device=dev)
        self.count = torch.zeros(A, dtype=torch.uint8, device=dev)
        self.nbr = torch.full((A, nm), -1, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(A, dtype=torch.float32, device=dev)
        self.done = torch.zeros(A, dtype=torch.uint8, device=dev)
        self.info = torch.zeros(A, dtype=torch.uint8, device=dev)
        self.finish = torch.zeros(A, dtype=torch.uint8, device=dev)
        self.reset_mask = torch.zeros(A, dtype=torch.uint8, device=dev)
        self.path_cut = torch.zeros(A, dtype=torch.uint8, device=dev)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.uaisa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- the hot path --------------------------------------------------------------------------
    def step(self, action: torch.Tensor):
        """mdin.drone_step (mdin.py:19-30) for all envs. action: float32 CUDA [A,3] (or [E,N,3])."""
        if action.dtype != torch.float32 or not action.is_cuda or not action.is_contiguous() or action.numel() != self.A * 3:
            raise ValueError("action must be a contiguous float32 CUDA tensor with E*N*3 elements")
        _capi.check(self.L.uaisa_step(self.h, _ptr(action), _ptr(self.obs), _ptr(self.count), _ptr(self.nbr), _ptr(self.reward),
                                      _ptr(self.done), _ptr(self.info), _ptr(self.finish), self._stream()))
        return self.obs, self.reward, self.done, self.info, self.finish

    def rollout(self, actions: torch.Tensor, out: Optional[dict] = None, auto_reset: bool = False, max_ep_len: int = 300):
        """T open-loop steps in one launch. actions [T,A,3]; `out` maps output name -> [T,...] tensor (None = skip)."""
        T = int(actions.shape[0])
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous() or actions.numel() != T * self.A * 3:
            raise ValueError("actions must be a contiguous float32 CUDA tensor [T, E*N, 3]")
        out = out or {}
        g = lambda k: _ptr(out.get(k))  # noqa: E731
        _capi.check(self.L.uaisa_rollout(self.h, T, _ptr(actions), g("obs"), g("count"), g("nbr"), g("reward"), g("done"),
                                         g("info"), g("finish"), int(auto_reset), int(max_ep_len), self._stream()))
        return out

    def step_autoreset(self, action: torch.Tensor, max_ep_len: int = 300, epoch_ended: bool = False, out: Optional[dict] = None):
        """step() followed by autoreset() in one launch (same outputs, bit for bit): the env part of one trainer iteration (multi_ppo.py:210-281).
        `out` may redirect any of obs / count / reward / path_cut to caller tensors of the same shapes (e.g. slices of a rollout buffer: no copy).
        -> (obs of the next step, reward, done, info, finish, reset_mask, path_cut)."""
        if action.dtype != torch.float32 or not action.is_cuda or not action.is_contiguous() or action.numel() != self.A * 3:
            raise ValueError("action must be a contiguous float32 CUDA tensor [E*N, 3]")
        o = {"obs": self.obs, "count": self.count, "reward": self.reward, "path_cut": self.path_cut}
        for k, t in (out or {}).items():
            if k not in o or t.dtype != o[k].dtype or t.numel() != o[k].numel() or not t.is_cuda or not t.is_contiguous():
                raise ValueError("out[%r] must be a contiguous CUDA tensor with the dtype and size of the env's own %s buffer" % (k, k))
            o[k] = t
        _capi.check(self.L.uaisa_step_autoreset(self.h, _ptr(action), _ptr(o["obs"]), _ptr(o["count"]), _ptr(self.nbr), _ptr(o["reward"]), _ptr(self.done),
                                                _ptr(self.info), _ptr(self.finish), int(max_ep_len), int(bool(epoch_ended)), _ptr(self.reset_mask),
                                                _ptr(o["path_cut"]), self._stream()))
        return o["obs"], o["reward"], self.done, self.info, self.finish, self.reset_mask, o["path_cut"]

    def reset(self, agent_mask: Optional[torch.Tensor] = None, observe: bool = True):
        """mdin.drone_reset (agent_mask None) / drone_reset_one + env_observation (mdin.py:38-46)."""
        if agent_mask is not None and (agent_mask.dtype != torch.uint8 or agent_mask.numel() != self.A or not agent_mask.is_cuda):
            raise ValueError("agent_mask must be a uint8 CUDA tensor with E*N elements")
        o = (self.obs, self.count, self.nbr) if observe else (None, None, None)
        _capi.check(self.L.uaisa_reset(self.h, _ptr(agent_mask), _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), self._stream()))
        return self.obs

    def observe(self, env_mask: Optional[torch.Tensor] = None):
        """ir_gym.env_observation (ir_gym.py:372-383)."""
        if env_mask is not None and (env_mask.dtype != torch.uint8 or env_mask.numel() != self.E or not env_mask.is_cuda):
            raise ValueError("env_mask must be a uint8 CUDA tensor with E elements")
        _capi.check(self.L.uaisa_observe(self.h, _ptr(env_mask), _ptr(self.obs), _ptr(self.count), _ptr(self.nbr), self._stream()))
        return self.obs

    def autoreset(self, max_ep_len: int = 300, epoch_ended: bool = False, done=None, finish=None):
        """Trainer reset policy (multi_ppo.py:221-281) applied to the flags of the last step."""
        done = self.done if done is None else done
        finish = self.finish if finish is None else finish
        _capi.check(self.L.uaisa_autoreset(self.h, _ptr(done), _ptr(finish), int(max_ep_len), int(bool(epoch_ended)), _ptr(self.obs),
                                           _ptr(self.count), _ptr(self.nbr), _ptr(self.reset_mask), _ptr(self.path_cut), self._stream()))
        return self.obs, self.reset_mask, self.path_cut

    # -- state / indicators --------------------------------------------------------------------
    def _state_buffers(self):
        dev, A = self.device, self.A
        return {"pos_yaw": torch.empty((A, 4), dtype=torch.float32, device=dev), "vel_pitch": torch.empty((A, 4), dtype=torch.float32, device=dev),
                "real_len": torch.empty(A, dtype=torch.float32, device=dev), "max_dev": torch.empty(A, dtype=torch.float32, device=dev),
                "extra_len": torch.empty(A, dtype=torch.float32, device=dev), "wp_flags": torch.empty(A, dtype=torch.int32, device=dev)}

    def get_state(self) -> dict:
        b = self._state_buffers()
        v = _capi.UaisaStateView(*[b[n].data_ptr() for n, _ in _capi.UaisaStateView._fields_])
        _capi.check(self.L.uaisa_get_state(self.h, C.byref(v), self._stream()))
        return b

    def set_state(self, st: dict):
        ptrs = []
        keep = []
        for n, _ in _capi.UaisaStateView._fields_:
            t = st.get(n)
            if t is not None:
                want = torch.int32 if n == "wp_flags" else torch.float32
                t = t.to(device=self.device, dtype=want).contiguous()
                keep.append(t)
            ptrs.append(None if t is None else t.data_ptr())
        v = _capi.UaisaStateView(*ptrs)
        _capi.check(self.L.uaisa_set_state(self.h, C.byref(v), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()   # `keep` must outlive the async copies

    def cal_des(self) -> torch.Tensor:
        """ir_gym.cal_des_list (ir_gym.py:44-46) -> [A,3]."""
        des = torch.empty((self.A, 3), dtype=torch.float32, device=self.device)
        _capi.check(self.L.uaisa_cal_des(self.h, _ptr(des), self._stream()))
        return des

    def rvo_select(self, vmax: float = 2.0, acceler: float = 0.5, controls: bool = True):
        """K7, classic RVO velocity selection (reciprocal_vel_obs.cal_vel, reciprocal_vel_obs.py:21-166; corrected restatement, see include/uaisa_b200.h)
        -> (selected velocity [A,3], the same as env controls [A,3] to feed :meth:`step`, or None)."""
        vel = torch.empty((self.A, 3), dtype=torch.float32, device=self.device)
        ctrl = torch.empty((self.A, 3), dtype=torch.float32, device=self.device) if controls else None
        _capi.check(self.L.uaisa_rvo_select(self.h, float(vmax), float(acceler), _ptr(vel), _ptr(ctrl), self._stream()))
        return vel, ctrl

    def action_transform(self, a_inc: torch.Tensor, action_out: Optional[torch.Tensor] = None, a_inc_rounded: Optional[torch.Tensor] = None):
        """multi_ppo.py:197-205 on device: returns (action[A,3] for step(), rounded increment[A,3] for the buffer)."""
        if action_out is None:
            action_out = torch.empty((self.A, 3), dtype=torch.float32, device=self.device)
        if a_inc_rounded is None:
            a_inc_rounded = torch.empty((self.A, 3), dtype=torch.float32, device=self.device)
        _capi.check(self.L.uaisa_action_transform(self.h, _ptr(a_inc.contiguous()), float(self.acceler), _ptr(a_inc_rounded), _ptr(action_out), self._stream()))
        return action_out, a_inc_rounded

    def indicators(self):
        """(max_deviation[A], extra_len[A]) -- ir_gym.indicators_deviation / indicators_extra_len."""
        md = torch.empty(self.A, dtype=torch.float32, device=self.device)
        el = torch.empty(self.A, dtype=torch.float32, device=self.device)
        _capi.check(self.L.uaisa_indicators(self.h, _ptr(md), _ptr(el), self._stream()))
        return md, el

    def counters(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _capi.check(self.L.uaisa_get_counters(self.h, C.byref(a), C.byref(b), C.byref(c), self._stream()))
        return {"in_range_pairs": a.value, "cone_evals": b.value, "nonfinite_obs": c.value}

    REGIONS = ("prologue", "fill", "sync", "dronestate", "pairs", "reward_a", "move", "obsrow", "tail_b", "autoreset", "epilogue", "b_flags", "b_col", "b_reward", "b_out", "act")

    def region_clocks(self) -> dict:
        """Per-region SM-cycle sums (profiling builds only: -DUAISA_REGION_CLOCKS)."""
        buf = (C.c_int64 * 16)()
        _capi.check(self.L.uaisa_get_region_clocks(self.h, buf, self._stream()))
        return {n: int(buf[i]) for i, n in enumerate(self.REGIONS)}

    @property
    def launch_count(self) -> int:
        return int(self.L.uaisa_launch_count(self.h))

    def step_host(self, action: np.ndarray, out: dict, auto_reset: bool = False, max_ep_len: int = 300):
        """End-to-end call with HOST buffers (numpy, ideally pinned): H2D action, step, D2H outputs."""
        p = lambda k: None if out.get(k) is None else out[k].ctypes.data_as(C.c_void_p)  # noqa: E731
        with torch.cuda.device(self.device):
            _capi.check(self.L.uaisa_step_host(self.h, action.ctypes.data_as(C.c_void_p), p("obs"), p("count"), p("nbr"), p("reward"),
                                               p("done"), p("info"), p("finish"), int(auto_reset), int(max_ep_len)))
        return out


    ROW_DTYPE = np.dtype([("v", np.float32, (9,)), ("agent", np.int32), ("slot", np.int32), ("nbr", np.int32)])

    def step_host_sparse(self, action: np.ndarray, auto_reset: bool = False, max_ep_len: int = 300) -> dict:
        """Latency-optimised host round trip (uaisa_step_host_sparse): float32 [A,3] numpy action in; a dict of numpy
        VIEWS into library-owned pinned memory out (the SAME dict and views every call, valid until the next call): reward[A],
        self_obs[A,12], done/info/finish/count[A], rows[n_rows] (structured: v[9], agent, slot, nbr)."""
        if action.dtype != np.float32 or action.size != self.A * 3 or not action.flags.c_contiguous:
            raise ValueError("action must be a C-contiguous float32 array with E*N*3 elements")
        hv = self.__dict__.get("_hv")
        if hv is None:               # first call: one result struct, views into the library's pinned block and ONE result dict, reused by every call
            self._hs = _capi.UaisaHostStep()
            self._hs_ref = C.byref(self._hs)
            self._hss = self.L.uaisa_step_host_sparse
        hs = self._hs
        # no torch device context here: the library selects (and restores) the handle's device itself
        rc = self._hss(self.h, action.__array_interface__["data"][0], 1 if auto_reset else 0, int(max_ep_len), self._hs_ref)
        if rc:
            _capi.check(rc)
        if hv is None:
            A = self.A
            view = lambda p, n, dt: np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p), dtype=dt, count=n)  # noqa: E731
            rows_cap = max(1, A * max(self.nm, 1))
            self._rows_all = view(hs.rows, min(rows_cap, (256 << 20) // 48), self.ROW_DTYPE)
            hv = self._hv = {"reward": view(hs.reward, A, np.float32), "self_obs": view(hs.self_obs, A * 12, np.float32).reshape(A, 12),
                             "done": view(hs.done, A, np.uint8), "info": view(hs.info, A, np.uint8), "finish": view(hs.finish, A, np.uint8),
                             "count": view(hs.count, A, np.uint8), "rows": None}
        hv["rows"] = self._rows_all[: hs.n_rows]
        return hv

    @staticmethod
    def ragged_from_sparse(out: dict, nm: int):
        """Sparse host step -> the reference's ragged observation list (12 + 9*max(k,1) values per agent)."""
        A = out["self_obs"].shape[0]
        cnt = out["count"]
        obs = [np.zeros(12 + 9 * max(int(cnt[a]), 1), np.float32) for a in range(A)]
        for a in range(A):
            obs[a][:12] = out["self_obs"][a]
        for r in out["rows"]:
            obs[int(r["agent"])][12 + 9 * int(r["slot"]): 21 + 9 * int(r["slot"])] = r["v"]
        return obs


# helpers to move between the packed fp32 device state and the oracle's field names -------------------
def unpack_state(st: dict) -> dict:
    """Device state dict -> numpy fields named like oracle.STATE_FIELDS (values are exact fp32->fp64)."""
    py = st["pos_yaw"].detach().cpu().numpy().astype(np.float64)
    vp = st["vel_pitch"].detach().cpu().numpy().astype(np.float64)
    wf = st["wp_flags"].detach().cpu().numpy()
    return {"pos": py[:, :3].copy(), "vel": vp[:, :3].copy(), "yaw": py[:, 3].copy(), "pitch": vp[:, 3].copy(),
            "wp_idx": (wf & 0xFFFF).astype(np.int32), "arrive_flag": ((wf >> 16) & 1).astype(np.uint8),
            "dest_flag": ((wf >> 17) & 1).astype(np.uint8),
            "real_len": st["real_len"].detach().cpu().numpy().astype(np.float64),
            "max_dev": st["max_dev"].detach().cpu().numpy().astype(np.float64),
            "extra_len": st["extra_len"].detach().cpu().numpy().astype(np.float64)}


def pack_state(f: dict) -> dict:
    """Inverse of unpack_state (numpy fields -> torch CPU tensors ready for set_state)."""
    py = np.concatenate([np.asarray(f["pos"], np.float32), np.asarray(f["yaw"], np.float32)[:, None]], axis=1)
    vp = np.concatenate([np.asarray(f["vel"], np.float32), np.asarray(f["pitch"], np.float32)[:, None]], axis=1)
    wf = (np.asarray(f["wp_idx"], np.int32) & 0xFFFF) | (np.asarray(f["arrive_flag"], np.int32) << 16) | (np.asarray(f["dest_flag"], np.int32) << 17)
    return {"pos_yaw": torch.from_numpy(py), "vel_pitch": torch.from_numpy(vp),
            "real_len": torch.from_numpy(np.asarray(f["real_len"], np.float32)), "max_dev": torch.from_numpy(np.asarray(f["max_dev"], np.float32)),
            "extra_len": torch.from_numpy(np.asarray(f["extra_len"], np.float32)), "wp_flags": torch.from_numpy(wf.astype(np.int32))}

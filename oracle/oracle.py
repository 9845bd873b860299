"""ctypes wrapper around oracle/libuaisa_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libuaisa_oracle.so")


class UoConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "goal_threshold", "max_acc", "max_angle_change", "dt", "vel_max", "neighbor_range",
        "building_range", "building_zgate", "ctime_threshold", "delta_t", "exp_radius")] + [
        ("env_train", C.c_int32), ("nm", C.c_int32)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("uaisa_oracle.c", "uaisa_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.uo_default_config.argtypes = [C.POINTER(UoConfig)]
        L.uo_create.restype = vp
        L.uo_create.argtypes = [C.POINTER(UoConfig)] + [C.c_int] * 5 + [vp] * 8
        L.uo_destroy.argtypes = [vp]
        L.uo_set_threads.argtypes = [C.c_int]
        L.uo_get_max_threads.restype = C.c_int
        L.uo_reset.argtypes = [vp, vp]
        L.uo_observe.argtypes = [vp] * 5
        L.uo_step.argtypes = [vp] * 9
        L.uo_autoreset.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp]
        L.uo_get_state.argtypes = [vp] * 11
        L.uo_set_state.argtypes = [vp] * 11
        L.uo_cal_des.argtypes = [vp, vp]
        L.uo_get_route_len.argtypes = [vp, vp]
        L.uo_get_counters.argtypes = [vp, vp, vp]
        L.uo_rvo_select.argtypes = [vp, C.c_double, C.c_double, vp, vp]
        L.uo_pair.restype = C.c_int
        L.uo_pair.argtypes = [C.POINTER(UoConfig)] + [vp] * 6
        L.uo_pyround2.restype = C.c_double
        L.uo_pyround2.argtypes = [C.c_double]
        L.uo_npround.restype = C.c_double
        L.uo_npround.argtypes = [C.c_double, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_config(nm: int = 10, **over) -> UoConfig:
    c = UoConfig()
    lib().uo_default_config(C.byref(c))
    c.nm = nm
    for k, v in over.items():
        setattr(c, k, v)
    return c


STATE_FIELDS = (("pos", np.float64, 3), ("vel", np.float64, 3), ("yaw", np.float64, 0), ("pitch", np.float64, 0),
                ("wp_idx", np.int32, 0), ("arrive_flag", np.uint8, 0), ("dest_flag", np.uint8, 0),
                ("real_len", np.float64, 0), ("max_dev", np.float64, 0), ("extra_len", np.float64, 0))


class OracleBatch:
    """E independent envs, each a reference ``mdin`` (fp64), stepped by the C restatement."""

    def __init__(self, scn, E: int, env2scn=None, nm: int = 10, **cfg_over):
        self.L = lib()
        self.cfg = default_config(nm, **cfg_over)
        self.E, self.N, self.nm = int(E), scn.N, int(nm)
        self.A = self.E * self.N
        self.obs_dim = 12 + 9 * self.nm
        self.scn = scn
        if env2scn is None:
            env2scn = np.arange(self.E, dtype=np.int32) % scn.S
        self.env2scn = np.ascontiguousarray(env2scn, dtype=np.int32)
        arrs = [np.ascontiguousarray(scn.waypoints, np.float64), np.ascontiguousarray(scn.n_points, np.int32),
                np.ascontiguousarray(scn.radius, np.float64), np.ascontiguousarray(scn.priority, np.float64),
                np.ascontiguousarray(scn.buildings, np.float64), np.ascontiguousarray(scn.n_buildings, np.int32),
                np.ascontiguousarray(scn.map_size, np.float64), self.env2scn]
        self.h = self.L.uo_create(C.byref(self.cfg), self.E, self.N, scn.W, scn.M, scn.S, *[_p(a) for a in arrs])
        self._alloc()

    def _alloc(self):
        A, nm = self.A, self.nm
        self.obs = np.zeros((A, self.obs_dim), np.float64)
        self.count = np.zeros(A, np.int32)
        self.nbr = np.full((A, max(nm, 1)), -1, np.int32)[:, :nm].copy() if nm else np.zeros((A, 0), np.int32)
        self.reward = np.zeros(A, np.float64)
        self.done = np.zeros(A, np.uint8)
        self.info = np.zeros(A, np.uint8)
        self.finish = np.zeros(A, np.uint8)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.uo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self, agent_mask=None):
        m = None if agent_mask is None else np.ascontiguousarray(agent_mask, np.uint8)
        self.L.uo_reset(self.h, _p(m))

    def observe(self, env_mask=None):
        m = None if env_mask is None else np.ascontiguousarray(env_mask, np.uint8)
        self.L.uo_observe(self.h, _p(m), _p(self.obs), _p(self.count), _p(self.nbr))
        return self.obs, self.count, self.nbr

    def step(self, action):
        a = np.ascontiguousarray(action, np.float64).reshape(self.A, 3)
        self.L.uo_step(self.h, _p(a), _p(self.obs), _p(self.count), _p(self.nbr), _p(self.reward), _p(self.done),
                       _p(self.info), _p(self.finish))
        return self.obs, self.reward, self.done, self.info, self.finish

    def autoreset(self, done, finish, ep_len, max_ep_len: int, epoch_ended: bool):
        reset_mask = np.zeros(self.A, np.uint8)
        env_obs = np.zeros(self.E, np.uint8)
        path_cut = np.zeros(self.A, np.uint8)
        self.L.uo_autoreset(self.h, _p(np.ascontiguousarray(done, np.uint8)), _p(np.ascontiguousarray(finish, np.uint8)),
                            _p(ep_len), int(max_ep_len), int(bool(epoch_ended)), _p(reset_mask), _p(env_obs), _p(path_cut))
        return reset_mask, env_obs, path_cut

    def get_state(self) -> dict:
        out = {n: np.zeros((self.A, k) if k else (self.A,), dt) for n, dt, k in STATE_FIELDS}
        self.L.uo_get_state(self.h, *[_p(out[n]) for n, _, _ in STATE_FIELDS])
        return out

    def set_state(self, st: dict):
        arrs = []
        for n, dt, k in STATE_FIELDS:
            v = st.get(n)
            arrs.append(None if v is None else np.ascontiguousarray(v, dt).reshape((self.A, k) if k else (self.A,)))
        self.L.uo_set_state(self.h, *[_p(a) for a in arrs])

    def cal_des(self):
        d = np.zeros((self.A, 3), np.float64)
        self.L.uo_cal_des(self.h, _p(d))
        return d

    def rvo_select(self, vmax: float = 2.0, acceler: float = 0.5):
        """K7: classic RVO velocity selection (corrected restatement of reciprocal_vel_obs.cal_vel; PARITY UNPINNED) -> (velocity [A,3], env controls [A,3])."""
        v, c = np.zeros((self.A, 3), np.float64), np.zeros((self.A, 3), np.float64)
        self.L.uo_rvo_select(self.h, float(vmax), float(acceler), _p(v), _p(c))
        return v, c

    def route_len(self):
        r = np.zeros((self.scn.S, self.N), np.float64)
        self.L.uo_get_route_len(self.h, _p(r))
        return r

    def counters(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.L.uo_get_counters(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def pair(sa8, sb8, act, cfg=None):
    """rvo_inter.config_vo_circle2 known-answer entry: -> (row9, vo_flag, t, collision, md)."""
    cfg = cfg or default_config()
    sa8 = np.ascontiguousarray(sa8, np.float64)
    sb8 = np.ascontiguousarray(sb8, np.float64)
    act = np.ascontiguousarray(act, np.float64)
    row = np.zeros(9)
    t, md = C.c_double(0), C.c_double(0)
    f = lib().uo_pair(C.byref(cfg), _p(sa8), _p(sb8), _p(act), _p(row), C.byref(t), C.byref(md))
    return row, bool(f & 1), t.value, bool(f & 2), md.value


def set_threads(n: int):
    lib().uo_set_threads(int(n))


def max_threads() -> int:
    return int(lib().uo_get_max_threads())

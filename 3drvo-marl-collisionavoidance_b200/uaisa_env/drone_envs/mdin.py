"""``mdin`` -- one-environment adapter with the reference's list-of-numpy API.

Drop-in for ``uaisa_env.drone_envs.mdin.mdin`` of the reference (uaisa_env/drone_envs/mdin.py:6-49) as
``train/train_process.py`` / ``train/policy/multi_ppo.py`` / ``train/policy/post_train.py`` use it: the same
method names, argument meaning, return layout (ragged observations of length 12+9*max(k,1), python lists)
and error behaviour (``ValueError`` on a NaN/Inf observation, ir_gym.py:232-239).  All arithmetic happens in
libuaisa_b200.so on the GPU (E = 1); there is no CPU fallback.

The scenario comes from ``base_dir`` (a directory holding ``data_1.json``, like the reference's
``env_base.__init__``, env_base.py:15-47), or from ``scenario=`` (a ScenarioSet or a data_1.json dict).
For thousands of environments at once use :class:`uaisa_env.vec_env.VecMdin` instead.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..scenario import ScenarioSet
from ..vec_env import VecMdin, unpack_state

try:  # gym is optional: the reference subclasses gym.Env and builds two spaces.Box (mdin.py:6, ir_gym.py:31-32)
    import gymnasium as _gym
except Exception:  # pragma: no cover
    try:
        import gym as _gym
    except Exception:
        _gym = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))


def _box(low, high, shape=None):
    if _gym is not None:
        try:
            return _gym.spaces.Box(low, high, shape=shape, dtype=np.float32)
        except Exception:
            pass
    return _Box(low, high, shape)


class _DroneView:
    """What callers read/write on ``env.ir_gym.drone_list[i]`` (multi_ppo.py:202,207; post_train.py:72,78)."""

    def __init__(self, owner, idx):
        self._o, self.id = owner, idx
        self.rvo_vel = [0, 0, 0]          # write-only plotting attribute in the reference

    def _f(self, name):
        return self._o._state()[name][self.id]

    vel = property(lambda s: s._f("vel").copy())
    state = property(lambda s: s._f("pos").copy())
    yaw = property(lambda s: float(s._f("yaw")))
    pitch = property(lambda s: float(s._f("pitch")))
    i = property(lambda s: int(s._f("wp_idx")))
    arrive_flag = property(lambda s: bool(s._f("arrive_flag")))
    destination_arrive_flag = property(lambda s: bool(s._f("dest_flag")))
    real_route_len = property(lambda s: float(s._f("real_len")))
    max_deviation = property(lambda s: float(s._f("max_dev")))
    extra_len = property(lambda s: float(s._f("extra_len")))
    radius = property(lambda s: float(s._o.scn.radius[0, s.id]))
    priority = property(lambda s: float(s._o.scn.priority[0, s.id]))

    def cal_des_vel(self):
        return self._o.cal_des_list()[self.id]


class _Drones:
    """``ir_gym.components['drones']`` (mdin.py:46 calls ``drone_reset(id)``)."""

    def __init__(self, owner):
        self._o = owner
        self.drone_list = owner.drone_list

    def drone_reset(self, id=0):
        self._o._reset_one(int(id))

    def drones_reset(self):
        self._o._vec.reset(observe=False)
        self._o._dirty()

    def total_states(self):
        st = self._o._state()
        des = np.asarray(self._o.cal_des_list())
        n = self._o.drone_num
        return [np.concatenate([st["pos"][k], st["vel"][k], [self._o.scn.radius[0, k]], [self._o.scn.priority[0, k]], des[k], [0.0]]) for k in range(n)]

    def deviation_indicator(self):
        return self._o.indicators_deviation()

    def extar_len_indicator(self):
        return self._o.indicators_extra_len()


class ir_gym:
    """The attributes/methods of the reference's ``ir_gym`` that the trainer reaches through ``env.ir_gym``."""

    def __init__(self, scn: ScenarioSet, neighbors_region=5, neighbors_num=10, env_train=True, acceler=0.5, device="cuda", **hyper):
        self.scn = scn
        self.nr, self.nm = neighbors_region, int(neighbors_num)
        self.env_train, self.acceler = env_train, acceler
        self.drone_num = scn.N
        self.map_size = [float(x) for x in scn.map_size[0]]
        self.building_list = scn.buildings[0, : scn.n_buildings[0]].tolist()
        self.waypoints_list = [scn.waypoints[0, k, : scn.n_points[0, k]].tolist() for k in range(scn.N)]
        self.n_points_list = [int(x) for x in scn.n_points[0]]
        self.plot = False
        self.rvo_state_dim = 9
        self.observation_space = _box(-np.inf, np.inf, shape=(21,))                      # ir_gym.py:31 (nominal)
        self.action_space = _box(np.array([-1, -1, -1], np.float32), np.array([1, 1, 1], np.float32))
        self._vec = VecMdin(scn, 1, neighbors_num=self.nm, device=device, env_train=bool(env_train), acceler=acceler, **hyper)
        self._cache = None
        self.drone_list = [_DroneView(self, k) for k in range(scn.N)]
        self.components = {"drones": _Drones(self)}
        self._vec.reset(observe=False)

    # -- state cache (one small D2H per step, only when somebody looks) -------------------------------
    def _dirty(self):
        self._cache = None

    def _state(self):
        if self._cache is None:
            self._cache = unpack_state(self._vec.get_state())
        return self._cache

    # -- observation plumbing ----------------------------------------------------------------------------
    def _obs_list(self):
        v = self._vec
        bad = v.counters()["nonfinite_obs"]
        obs = np.round(v.obs.cpu().numpy().astype(np.float64), 2)     # fp32 image -> the reference's 2-dp doubles
        cnt = v.count.cpu().numpy()
        if bad or not np.isfinite(obs).all():
            raise ValueError("observation contains NaN/Inf (ir_gym.py:232-239)")
        return [obs[k, : 12 + 9 * max(int(cnt[k]), 1)].copy() for k in range(self.drone_num)]

    def _reset_one(self, idx):
        m = torch.zeros(self.drone_num, dtype=torch.uint8, device=self._vec.device)
        m[idx] = 1
        self._vec.reset(m, observe=False)
        self._dirty()

    # -- reference API -------------------------------------------------------------------------------------
    def env_reset(self):                                   # ir_gym.py:360-367
        self._vec.reset()
        self._dirty()
        return self._obs_list()

    def env_observation(self):                             # ir_gym.py:372-383
        self._vec.observe()
        self._dirty()
        return self._obs_list()

    def cal_des_list(self):                                # ir_gym.py:44-46
        d = np.round(self._vec.cal_des().cpu().numpy().astype(np.float64), 3)
        return [d[k].copy() for k in range(self.drone_num)]

    def indicators_deviation(self):                        # ir_gym.py:414-416
        return [float(x) for x in self._vec.indicators()[0].cpu().numpy()]

    def indicators_extra_len(self):                        # ir_gym.py:418-420
        return [float(x) for x in self._vec.indicators()[1].cpu().numpy()]

    def step(self, action_list):
        if not isinstance(action_list, list):
            action_list = [action_list]
        if len(action_list) != self.drone_num:
            raise ValueError("expected %d actions" % self.drone_num)
        act = np.stack([np.asarray(a, dtype=np.float64).reshape(-1) for a in action_list])
        assert act.shape == (self.drone_num, 3)            # drone.py:101
        # one latency-optimised host round trip (uaisa_step_host_sparse): host action in, ragged result out
        r = self._vec.step_host_sparse(np.ascontiguousarray(act, dtype=np.float32), auto_reset=False)
        self._dirty()
        cnt = r["count"]
        self_obs = np.round(r["self_obs"].astype(np.float64), 2)          # fp32 image -> the reference's 2-dp doubles
        obs = [np.zeros(12 + 9 * max(int(cnt[k]), 1)) for k in range(self.drone_num)]
        for k in range(self.drone_num):
            obs[k][:12] = self_obs[k]
        rows = r["rows"]
        if len(rows):
            rv = np.round(rows["v"].astype(np.float64), 2)
            for q in range(len(rows)):
                sl = int(rows["slot"][q])
                obs[int(rows["agent"][q])][12 + 9 * sl: 21 + 9 * sl] = rv[q]
        if not all(np.isfinite(o).all() for o in obs):
            raise ValueError("observation contains NaN/Inf (ir_gym.py:232-239)")
        rew = np.round(r["reward"].astype(np.float64), 3)
        done = [bool(x) for x in r["done"]]
        info = [bool(x) for x in r["info"]]
        fin = [bool(x) for x in r["finish"]]
        return obs, [x for x in rew], done, info, fin

    def render(self, *a, **k):                             # plotting is out of scope
        pass

    def close(self):
        self._vec.close()


_Base = _gym.Env if _gym is not None else object


class mdin(_Base):
    def __init__(self, world_name=None, neighbors_region=5, neighbors_num=10, base_dir=None, scenario=None, **kwargs):
        if scenario is None:
            if base_dir is None:
                raise ValueError("pass base_dir=<directory with data_1.json> or scenario=<ScenarioSet|dict> "
                                 "(the reference's default base_dir is a Windows path, env_base.py:15)")
            scenario = ScenarioSet.from_json(base_dir)
        elif isinstance(scenario, dict):
            scenario = ScenarioSet.from_dict(scenario)
        elif isinstance(scenario, (str, os.PathLike)):
            scenario = ScenarioSet.from_json(os.fspath(scenario))
        if scenario.S != 1:
            raise ValueError("mdin drives exactly one scenario; use VecMdin for batches")
        self.world_name = world_name
        self.ir_gym = ir_gym(scenario, neighbors_region, neighbors_num, **kwargs)
        self.observation_space = self.ir_gym.observation_space
        self.action_space = self.ir_gym.action_space
        self.neighbors_region = neighbors_region
        self.rvo_observation_list = []
        self.drow_rvo_flag = True

    def drone_step(self, action, **kwargs):                # mdin.py:19-30
        return self.ir_gym.step(action)

    def drone_reset(self, ifrender=False):                 # mdin.py:38-41
        return self.ir_gym.env_reset()

    def drone_reset_one(self, ifrender, id):               # mdin.py:43-46
        self.ir_gym.components["drones"].drone_reset(id)

    def drone_render(self, save=False, path=None, i=0, **kwargs):   # mdin.py:32-36: rendering is out of scope
        pass

    def drone_show(self):
        pass

    def close(self):
        self.ir_gym.close()

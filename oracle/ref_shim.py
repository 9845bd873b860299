"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference env (``/root/reference/uaisa_env``) in this
container so that (a) the C restatement in ``oracle/uaisa_oracle.c`` can be
pinned against the real thing and (b) golden vectors can be generated
(``oracle/gen_golden.py``).  ``/root/reference`` does not exist on the GPU box,
so nothing under ``tests -m gpu``, ``smoke()`` or ``bench.py`` may call this.

What the shim does:
  * stub ``gym`` (``Env``, ``spaces.Box``) -- the reference only subclasses
    ``gym.Env`` (mdin.py:6) and builds two ``spaces.Box`` (ir_gym.py:31-32);
  * mock matplotlib / imageio / PIL, which env_base.py:7 and env_plot.py:4-18
    import at module load but the hot path never touches once ``plot=False``;
  * put the reference root on ``sys.path``.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("UAISA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "uaisa_env"))


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # noqa: D401 - stub
            pass

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.dtype = low, high, dtype
                self.shape = tuple(shape) if shape is not None else np.shape(low)

        spaces = types.ModuleType("gym.spaces")
        spaces.Box = Box
        gym.Env = Env
        gym.spaces = spaces
        gym.make = None
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    for name in (
        "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.image",
        "matplotlib.offsetbox", "matplotlib.patches", "matplotlib.transforms",
        "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
        "imageio", "PIL", "PIL.Image",
    ):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)


_MDIN = None


def load_reference_mdin():
    """Return the reference's ``mdin`` class (uaisa_env/drone_envs/mdin.py:6)."""
    global _MDIN
    if _MDIN is None:
        if not reference_available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        _install_stubs()
        # The product package is also called ``uaisa_env`` (it is a drop-in).  Import the reference's
        # package of the same name in isolation, then park it under ``_ref_uaisa_env*`` so both can
        # live in one process (the reference only does module-level imports on this path).
        saved = {k: v for k, v in sys.modules.items() if k == "uaisa_env" or k.startswith("uaisa_env.")}
        for k in saved:
            del sys.modules[k]
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            import importlib

            mdin = importlib.import_module("uaisa_env.drone_envs.mdin").mdin
        finally:
            sys.path.remove(REFERENCE_ROOT)
            for k in [k for k in sys.modules if k == "uaisa_env" or k.startswith("uaisa_env.")]:
                sys.modules["_ref_" + k] = sys.modules.pop(k)
            sys.modules.update(saved)
        _MDIN = mdin
    return _MDIN


def write_scenario_dir(path: str, scenario: dict) -> str:
    """Write ``data_1.json`` + dummy occupancy grids (env_base.py:26-47 reads all three)."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "data_1.json"), "w") as f:
        json.dump(scenario, f)
    np.save(os.path.join(path, "E3d.npy"), np.zeros((1, 1, 1)))
    np.save(os.path.join(path, "E3d_safe.npy"), np.zeros((1, 1, 1)))
    return path


def make_reference_env(scenario, nm: int = 10):
    """Build a reference env for ``scenario`` (dict in data_1.json schema, or a directory)."""
    mdin = load_reference_mdin()
    if isinstance(scenario, dict):
        d = tempfile.mkdtemp(prefix="uaisa_ref_")
        write_scenario_dir(d, scenario)
    else:
        d = scenario
    # env_base.__init__ hard-codes plot=True *before* init_environment (env_base.py:21-24), and
    # env_plot is a MagicMock'd-matplotlib consumer: harmless, but switch it off afterwards.
    env = mdin(world_name="x", neighbors_num=nm, base_dir=d)
    env.ir_gym.plot = False
    return env


# ---------------------------------------------------------------------------------------------
# state access helpers (teacher forcing), following drone.py:19-76 attribute names
# ---------------------------------------------------------------------------------------------
STATE_FIELDS = ("pos", "vel", "yaw", "pitch", "wp_idx", "arrive_flag", "dest_flag",
                "real_len", "max_dev", "extra_len")


def get_state(env) -> dict:
    dl = env.ir_gym.drone_list
    return {
        "pos": np.array([np.asarray(d.state, dtype=np.float64) for d in dl]),
        "vel": np.array([np.asarray(d.vel, dtype=np.float64) for d in dl]),
        "yaw": np.array([float(d.yaw) for d in dl]),
        "pitch": np.array([float(d.pitch) for d in dl]),
        "wp_idx": np.array([int(d.i) for d in dl], dtype=np.int32),
        "arrive_flag": np.array([bool(d.arrive_flag) for d in dl], dtype=np.uint8),
        "dest_flag": np.array([bool(d.destination_arrive_flag) for d in dl], dtype=np.uint8),
        "real_len": np.array([float(d.real_route_len) for d in dl]),
        "max_dev": np.array([float(np.squeeze(d.max_deviation)) for d in dl]),
        "extra_len": np.array([float(d.extra_len) for d in dl]),
    }


def set_state(env, st: dict) -> None:
    """Inject state as float64 copies (fp32 arrays would silently demote the first step)."""
    for k, d in enumerate(env.ir_gym.drone_list):
        d.state = np.asarray(st["pos"][k], dtype=np.float64).copy()
        d.previous_state = d.state.copy()
        d.vel = np.asarray(st["vel"][k], dtype=np.float64).copy()
        d.yaw = float(st["yaw"][k])
        d.pitch = float(st["pitch"][k])
        d.i = int(st["wp_idx"][k])
        d.current_des = np.array(d.waypoints[d.i], dtype=np.float64)
        d.previous_des = np.array(d.waypoints[d.i - 1], dtype=np.float64)
        d.arrive_flag = bool(st["arrive_flag"][k])
        d.destination_arrive_flag = bool(st["dest_flag"][k])
        d.real_route_len = float(st["real_len"][k])
        d.max_deviation = float(st["max_dev"][k])
        d.extra_len = float(st["extra_len"][k])


def pad_obs(obs_list, nm: int):
    """Ragged reference observations (12+9*max(k,1)) -> dense [N, 12+9*nm] + count[N]."""
    n = len(obs_list)
    out = np.zeros((n, 12 + 9 * nm), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    for i, o in enumerate(obs_list):
        o = np.asarray(o, dtype=np.float64)
        out[i, : o.shape[0]] = o
        k = (o.shape[0] - 12) // 9
        # k==1 with an all-zero row is the reference's "no neighbour" placeholder (ir_gym.py:211-212)
        if k == 1 and not np.any(o[12:21]):
            k = 0
        cnt[i] = k
    return out, cnt

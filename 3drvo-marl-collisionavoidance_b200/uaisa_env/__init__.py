"""uaisa_env -- B200-native drop-in for the reference's ``uaisa_env`` package (hot path only).

    from uaisa_env.vec_env import VecMdin          # batched: E envs x N drones on one GPU
    from uaisa_env.drone_envs.mdin import mdin     # 1-env adapter with the reference's list/numpy API

``gym.make('mdin-v1', ...)`` works when gym or gymnasium is importable (the reference never
registers the id itself -- SURVEY.md section 8b).
"""
from .scenario import ScenarioSet  # noqa: F401


def _register():
    for modname in ("gymnasium", "gym"):
        try:
            mod = __import__(modname)
            from .drone_envs.mdin import mdin  # noqa: F401

            mod.register(id="mdin-v1", entry_point="uaisa_env.drone_envs.mdin:mdin")
        except Exception:
            continue


_register()

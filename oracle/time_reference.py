"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Runs the UNMODIFIED reference in the build container (it cannot travel to the GPU box) and records

  1. the real CPU baseline SURVEY.md section 8d / BASELINE.md section 3 ask for:
       env:  P = nproc independent processes, one reference ``mdin`` each (uaisa_env/drone_envs/mdin.py:19-30), T steps of
             training-like actions with the trainer's reset policy (multi_ppo.py:226-281) -> agent-steps/s per core and aggregate;
       PPO:  ``multi_ppo.training_loop`` (train/policy/multi_ppo.py:166-339) for 2 epochs x 300 steps with the reference
             hyper-parameters (train/train_process.py:53-63), ``use_gpu=False`` -> samples/s, rollout / update split;
     -> profiles/r2/container_reference.json   (bench.py carries it, labelled "build-container CPU", next to its own numbers)
  2. the exact env-API call trace the reference trainer issues during one short epoch (method, argument types / shapes,
     returned types / shapes, the actions themselves) -> tests/golden/trainer_call_trace.json, replayed against the drop-in
     ``uaisa_env.drone_envs.mdin`` by tests/test_trainer_trace_gpu.py.

    python oracle/time_reference.py [--steps 2000] [--procs 0] [--skip-ppo] [--skip-trace]
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import platform
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]


def _scenario(name):
    from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_world3
    w8 = os.path.join(os.environ.get("UAISA_REFERENCE_ROOT", "/root/reference"), "uaisa_env", "world", "world_8")
    return {"c1": scenario_c1, "world_3": scenario_world3, "world_8": lambda: ScenarioSet.from_json(w8),
            "c3_n16": lambda: scenario_c3(seed=0, n_scenarios=1)}[name]()


def env_worker(args):
    """One process = one reference env (Python is single-threaded)."""
    name, steps, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"; os.environ["OPENBLAS_NUM_THREADS"] = "1"; os.environ["MKL_NUM_THREADS"] = "1"
    import ref_shim
    scn = _scenario(name)
    env = ref_shim.make_reference_env(scn.to_dict(0), nm=10)
    n = env.ir_gym.drone_num
    rng = np.random.default_rng(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        env.drone_reset(False)
        ep_len = [0] * n
        t0 = time.perf_counter()
        for t in range(steps):
            acts = []
            for i in range(n):      # multi_ppo.py:197-205 with a_inc ~ U(-1,1)^3
                a_inc = np.round(rng.uniform(-1, 1, 3), 2)
                acts.append(np.round(env.ir_gym.acceler * a_inc + np.squeeze(env.ir_gym.drone_list[i].vel), 2))
            _, _, done, _, fin = env.drone_step(acts)
            ep_len = [e + 1 for e in ep_len]
            if max(done):           # multi_ppo.py:230-242
                for i in range(n):
                    if done[i]:
                        env.drone_reset_one(False, i); ep_len[i] = 0
                env.ir_gym.env_observation()
            if min(fin) or (t + 1) % 300 == 0:      # :244-264
                env.drone_reset(False); ep_len = [0] * n
            elif max(fin) or max(ep_len) > 300:     # :266-281
                for i in range(n):
                    if fin[i] or ep_len[i] > 300:
                        env.drone_reset_one(False, i); ep_len[i] = 0
                env.ir_gym.env_observation()
        el = time.perf_counter() - t0
    return n * steps / el


def time_env(name, steps, procs):
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        per = pool.map(env_worker, [(name, steps, 100 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    return {"scenario": name, "processes": procs, "steps_per_process": steps, "agent_steps_per_s_per_core": float(np.mean(per)),
            "agent_steps_per_s_aggregate": float(np.sum(per)), "wall_s": wall}


# ------------------------------------------------------------------------------------------------------------------------
def _describe(x):
    """Structural signature of a value crossing the env boundary."""
    if isinstance(x, (list, tuple)):
        return {"t": type(x).__name__, "items": [_describe(v) for v in x]}
    if isinstance(x, np.ndarray):
        return {"t": "ndarray", "shape": list(x.shape), "dtype": str(x.dtype)}
    if isinstance(x, (bool, np.bool_)):
        return {"t": "bool"}
    if isinstance(x, (int, np.integer)):
        return {"t": "int"}
    if isinstance(x, (float, np.floating)):
        return {"t": "float"}
    if x is None:
        return {"t": "None"}
    return {"t": type(x).__name__}


class _Recorder:
    """Transparent proxy around the reference env that logs what multi_ppo.training_loop touches (multi_ppo.py:183-281)."""

    def __init__(self, env, log):
        object.__setattr__(self, "_env", env); object.__setattr__(self, "_log", log)
        object.__setattr__(self, "ir_gym", _IrGymRecorder(env.ir_gym, log))

    def __getattr__(self, k):
        v = getattr(self._env, k)
        if k in ("observation_space", "action_space"):
            self._log.append({"get": k + ".shape", "value": list(v.shape)})
        return v

    def drone_step(self, action, **kw):
        out = self._env.drone_step(action, **kw)
        self._log.append({"call": "drone_step", "args": [_describe(action)], "action": [np.asarray(a, np.float64).tolist() for a in action],
                          "ret": _describe(out), "done": [bool(x) for x in out[2]], "info": [bool(x) for x in out[3]], "finish": [bool(x) for x in out[4]],
                          "obs_len": [int(np.asarray(o).shape[0]) for o in out[0]]})
        return out

    def drone_reset(self, ifrender=False):
        out = self._env.drone_reset(ifrender)
        self._log.append({"call": "drone_reset", "args": [_describe(ifrender)], "ret": _describe(out)})
        return out

    def drone_reset_one(self, ifrender, id):
        out = self._env.drone_reset_one(ifrender, id)
        self._log.append({"call": "drone_reset_one", "args": [_describe(ifrender), _describe(id)], "id": int(id), "ret": _describe(out)})
        return out


class _DroneRecorder:
    def __init__(self, d, idx, log):
        object.__setattr__(self, "_d", d); object.__setattr__(self, "_i", idx); object.__setattr__(self, "_log", log)

    def __getattr__(self, k):
        v = getattr(self._d, k)
        self._log.append({"get": "ir_gym.drone_list[i]." + k, "i": self._i, "value": _describe(np.squeeze(v))})
        return v

    def __setattr__(self, k, v):
        self._log.append({"set": "ir_gym.drone_list[i]." + k, "i": self._i, "value": _describe(v)})
        setattr(self._d, k, v)


class _IrGymRecorder:
    def __init__(self, g, log):
        object.__setattr__(self, "_g", g); object.__setattr__(self, "_log", log)
        object.__setattr__(self, "drone_list", [_DroneRecorder(d, i, log) for i, d in enumerate(g.drone_list)])

    def __getattr__(self, k):
        v = getattr(self._g, k)
        if callable(v) and k in ("env_observation", "indicators_deviation", "indicators_extra_len", "cal_des_list"):
            def f(*a, **kw):
                out = v(*a, **kw)
                e = {"call": "ir_gym." + k, "args": [_describe(x) for x in a], "ret": _describe(out)}
                if k == "env_observation":
                    e["obs_len"] = [int(np.asarray(o).shape[0]) for o in out]
                self._log.append(e)
                return out
            return f
        if k in ("drone_num", "acceler"):
            self._log.append({"get": "ir_gym." + k, "value": v})
        return v


def run_reference_ppo(scn_name, epochs, steps_per_epoch, pi_iters, v_iters, trace=None, seed=7):
    """multi_ppo.training_loop through the shim; returns (samples/s, rollout_s, update_s)."""
    import ref_shim
    ref_shim._install_stubs()
    import torch
    torch.set_num_threads(1)
    torch.cuda.synchronize = lambda *a, **k: None      # policy_rnn_ac.py:39, multi_ppo.py:153 (CPU-only container)
    if os.path.join(ref_shim.REFERENCE_ROOT, "train") not in sys.path:
        sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "train"))
    import torch.nn as nn
    from policy.multi_ppo import multi_ppo
    from policy.policy_rnn_ac import rnn_ac

    scn = _scenario(scn_name)
    env = ref_shim.make_reference_env(scn.to_dict(0), nm=10)
    test_env = ref_shim.make_reference_env(scn.to_dict(0), nm=10)
    log = trace if trace is not None else []
    penv = _Recorder(env, log) if trace is not None else env
    policy = rnn_ac(env.observation_space, env.action_space, 12, 9, 256, (256, 256), (256, 256), nn.ReLU, nn.Tanh, nn.Identity, False, "biGRU", 0)
    d = tempfile.mkdtemp(prefix="uaisa_refppo_") + "/"
    ppo = multi_ppo(penv, policy, 4e-6, 5e-5, epochs - 1, steps_per_epoch, 500, 0.99, 0.97, 0.2, pi_iters, v_iters, 0.05, False, 5, False, seed, 10 ** 9, False,
                    d, "r", None, False, False, 0, test_env, 1400, 10, False, None)
    upd = {"s": 0.0}
    orig_update = ppo.update

    def timed_update(data):
        t = time.perf_counter(); orig_update(data); upd["s"] += time.perf_counter() - t
    ppo.update = timed_update
    buf = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        ppo.training_loop()
    wall = time.perf_counter() - t0
    n = env.ir_gym.drone_num
    return {"scenario": scn_name, "drones": n, "epochs": epochs, "steps_per_epoch": steps_per_epoch, "train_pi_iters": pi_iters, "train_v_iters": v_iters,
            "samples_per_s": n * steps_per_epoch * epochs / wall, "wall_s": wall, "update_s": upd["s"], "rollout_s": wall - upd["s"],
            "torch_threads": 1, "use_gpu": False}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--skip-ppo", action="store_true")
    ap.add_argument("--skip-trace", action="store_true")
    ap.add_argument("--skip-env", action="store_true")
    args = ap.parse_args()
    procs = args.procs or (os.cpu_count() or 1)
    out_path = os.path.join(ROOT, "profiles", "r2", "container_reference.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    cpu = ""
    try:
        cpu = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass
    res.update({"where": "build container (no GPU); the unmodified /root/reference through oracle/ref_shim.py", "cpu": cpu, "cores": os.cpu_count(),
                "python": platform.python_version(), "numpy": np.__version__})
    if not args.skip_env:
        res["env"] = [time_env("c1", args.steps, procs), time_env("world_8", max(200, args.steps // 4), procs), time_env("c3_n16", max(50, args.steps // 20), procs)]
        for r in res["env"]:
            print(r)
    if not args.skip_ppo:
        res["ppo"] = run_reference_ppo("c1", 2, 300, 50, 50)
        print(res["ppo"])
    json.dump(res, open(out_path, "w"), indent=1)
    if not args.skip_trace:
        log = []
        run_reference_ppo("c1", 1, 60, 2, 2, trace=log)
        scn = _scenario("c1")
        json.dump({"scenario": scn.to_dict(0), "nm": 10, "source": "multi_ppo.training_loop (train/policy/multi_ppo.py:166-339), 1 epoch x 60 steps, seed 7",
                   "trace": log}, open(os.path.join(ROOT, "tests", "golden", "trainer_call_trace.json"), "w"))
        print("trace entries:", len(log))


if __name__ == "__main__":
    main()

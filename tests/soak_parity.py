#!/usr/bin/env python
"""Long teacher-forced lock-step runs of the CUDA path against the oracle (not collected by pytest: minutes, not seconds).
   python tests/soak_parity.py > gpurun_out/soak.txt     (GPU box; prints one summary line per case and fails on any mismatch)"""
import os, sys, time, json
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "3drvo-marl-collisionavoidance_b200")]
import numpy as np
import _util  # noqa: F401
from gpu_harness import lockstep
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_c4, scenario_world3

def swarm(n, rad, seed):
    rng = np.random.default_rng(seed); ang = 2 * np.pi * np.arange(n) / n
    st = np.stack([5 + rad * np.cos(ang), 5 + rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    en = np.stack([5 - rad * np.cos(ang), 5 - rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    return ScenarioSet.from_start_end(np.round(st, 2), np.round(en, 2), (10, 10, 5), buildings=[[5, 5, 1.5, 0.3]])

CASES = [("configs[1] 1024x4, 300 steps", scenario_c1, 1024, 300, {}),
         ("configs[1] scenario 8192x4, 120 steps", scenario_c1, 8192, 120, {}),
         ("world_3 512 envs, 300 steps", scenario_world3, 512, 300, {}),
         ("configs[2] 4 scenarios x16 drones, 512 envs, 150 steps", lambda: scenario_c3(seed=5, n_scenarios=4), 512, 150, {}),
         ("swarm of 20, 256 envs, 200 steps", lambda: swarm(20, 1.3, 9), 256, 200, {}),
         ("swarm of 12, 512 envs, 200 steps, nm=4", lambda: swarm(12, 1.0, 2), 512, 200, {"nm": 4}),
         ("configs[3] 256 drones, 4 envs, 60 steps (cluster + neighbour slices)", lambda: scenario_c4(seed=2, n_scenarios=2), 4, 60, {}),
         ("100 drones dense, 6 envs, 60 steps", lambda: scenario_c4(seed=6, n_scenarios=1, n_drones=100), 6, 60, {})]
bad = 0
for name, mk, E, T, kw in CASES:
    t0 = time.time()
    st = lockstep(mk(), E, kw.get("nm", 10), T, seed=17, epoch_len=97, block=20)
    s = st.summary()
    ok = s["flag_mismatch"] == 0 and s["count_mismatch"] == 0 and s["nbr_mismatch"] == 0 and max(s["obs_maxrel"], s["rew_maxrel"], s["state_maxrel"]) <= 1e-5
    bad += not ok
    print("%s | %s | agent-steps %d rows %d dones %d resets %d | flag/count/nbr mismatches %d/%d/%d | inexact fp32 images obs %d/%d rew %d/%d state %d/%d | max rel %.2e | %.0f s"
          % ("OK " if ok else "BAD", name, s["n_agent_steps"], s["rows"], s["dones"], s["resets"], s["flag_mismatch"], s["count_mismatch"], s["nbr_mismatch"],
             s["obs_inexact"], s["obs_vals"], s["rew_inexact"], s["rew_vals"], s["state_inexact"], s["state_vals"],
             max(s["obs_maxrel"], s["rew_maxrel"], s["state_maxrel"]), time.time() - t0), flush=True)

# ---- T-step launches (deferred re-observation, owner / helper split, 128-register dense kernels, clusters + slices) against single steps, every output bit for bit
import torch
from uaisa_env.vec_env import VecMdin
ROLL = [("configs[1] 1024x4 (owner / helper split)", scenario_c1, 1024, 120, 20), ("configs[1] scenario 16384x4", scenario_c1, 16384, 60, 20),
        ("configs[2] 4 scenarios x16 drones, 1024 envs", lambda: scenario_c3(seed=5, n_scenarios=4), 1024, 60, 20),
        ("swarm of 20, 256 envs", lambda: swarm(20, 1.3, 9), 256, 80, 16), ("configs[3] 256 drones, 6 envs", lambda: scenario_c4(seed=2, n_scenarios=2), 6, 40, 8),
        ("100 drones dense, 6 envs", lambda: scenario_c4(seed=6, n_scenarios=1, n_drones=100), 6, 40, 8),
        # BASELINE.json's full sizes
        ("configs[4] env part 65536x4", scenario_c1, 65536, 40, 20), ("configs[2] 4096x16 + cylinders", lambda: scenario_c3(seed=0, n_scenarios=8), 4096, 20, 10),
        ("configs[3] 64x256", lambda: scenario_c4(seed=0, n_scenarios=2), 64, 20, 10)]
for name, mk, E, T, chunk in ROLL:
    t0 = time.time()
    scn = mk()
    a = VecMdin(scn, E, neighbors_num=10); b = VecMdin(scn, E, neighbors_num=10)
    a.reset(); b.reset()
    A = a.A
    acts = torch.from_numpy(np.round(np.random.default_rng(3).uniform(-1, 1, (T, A, 3)), 2).astype(np.float32)).cuda()
    out = {"obs": torch.zeros((T, A, a.obs_dim), device="cuda"), "count": torch.zeros((T, A), dtype=torch.uint8, device="cuda"),
           "nbr": torch.zeros((T, A, 10), dtype=torch.int32, device="cuda"), "reward": torch.zeros((T, A), device="cuda"),
           "done": torch.zeros((T, A), dtype=torch.uint8, device="cuda"), "info": torch.zeros((T, A), dtype=torch.uint8, device="cuda"),
           "finish": torch.zeros((T, A), dtype=torch.uint8, device="cuda")}
    for s0 in range(0, T, chunk):
        a.rollout(acts[s0:s0 + chunk], {k: v[s0:s0 + chunk] for k, v in out.items()}, auto_reset=True, max_ep_len=25)
    mism = resets = rows = 0
    for t in range(T):
        b.step(acts[t])
        for k in ("reward", "done", "info", "finish"):
            mism += int((torch.nan_to_num(getattr(b, k).float(), 7.0).view(-1) != torch.nan_to_num(out[k][t].float(), 7.0).view(-1)).sum())
        b.autoreset(max_ep_len=25)
        resets += int(b.reset_mask.sum()); rows += int(b.count.sum())
        for k in ("obs", "count", "nbr"):
            mism += int((getattr(b, k) != out[k][t]).sum())
    sa, sb = a.get_state(), b.get_state()
    mism += sum(int((sa[k] != sb[k]).sum()) for k in sa)
    bad += mism != 0
    print("%s | rollout == steps | %s | agent-steps %d rows %d resets %d | mismatching values %d | %.0f s" % ("OK " if mism == 0 else "BAD", name, A * T, rows, resets, mism, time.time() - t0), flush=True)
    a.close(); b.close()
sys.exit(1 if bad else 0)

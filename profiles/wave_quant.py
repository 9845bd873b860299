#!/usr/bin/env python
"""Wave quantisation of the env kernel: agent-steps/s of a T-step rollout launch at env counts around a whole number of resident-warp waves.
   python profiles/wave_quant.py c3 3552 4096 5328"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import bench
from uaisa_env.vec_env import VecMdin
w = sys.argv[1]
T = 20
desc, scn, _ = bench.workload(w)
for E in [int(x) for x in sys.argv[2:]]:
    env = VecMdin(scn, E, neighbors_num=10); env.reset()
    A = env.A
    acts = torch.from_numpy(bench.synth_actions(np.random.default_rng(0), T, A)).cuda()
    out = {"obs": torch.empty((T, A, env.obs_dim), device="cuda"), "count": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
           "nbr": torch.empty((T, A, 10), dtype=torch.int32, device="cuda"), "reward": torch.empty((T, A), device="cuda"),
           "done": torch.empty((T, A), dtype=torch.uint8, device="cuda"), "info": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
           "finish": torch.empty((T, A), dtype=torch.uint8, device="cuda")}
    for _ in range(3):
        env.rollout(acts, out, auto_reset=True, max_ep_len=300)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    R = 10
    for _ in range(R):
        env.rollout(acts, out, auto_reset=True, max_ep_len=300)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    print("%s E=%d warps=%d: %.3f ms per %d-step launch -> %.3e agent-steps/s" % (w, E, A // 32, ms, T, A * T / (ms * 1e-3)))
    env.close()

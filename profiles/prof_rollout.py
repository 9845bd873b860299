#!/usr/bin/env python
"""A warm multi-step rollout launch for ncu (representative of bench.py's timed region, unlike a cold T=1 launch):
   ncu --set full --import-source on -k regex:uaisa_env -s 3 -c 1 -o X python profiles/prof_rollout.py [workload] [T]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import bench
from uaisa_env.vec_env import VecMdin
w = sys.argv[1] if len(sys.argv) > 1 else "c5"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
desc, scn, E = bench.workload(w)
env = VecMdin(scn, E, neighbors_num=10); env.reset()
A = env.A
acts = torch.from_numpy(bench.synth_actions(np.random.default_rng(0), T, A)).cuda()
out = {"obs": torch.empty((T, A, env.obs_dim), device="cuda"), "count": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
       "nbr": torch.empty((T, A, 10), dtype=torch.int32, device="cuda"), "reward": torch.empty((T, A), device="cuda"),
       "done": torch.empty((T, A), dtype=torch.uint8, device="cuda"), "info": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
       "finish": torch.empty((T, A), dtype=torch.uint8, device="cuda")}
for _ in range(4):
    env.rollout(acts, out, auto_reset=True, max_ep_len=300)
torch.cuda.synchronize()

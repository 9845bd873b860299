#!/usr/bin/env python
"""Run configs[3] once through the unsliced kernel and once through the sliced instantiation at slices = 1 (UAISA_FORCE_SL),
so that one ncu pass can compare them:  ncu --section ... python profiles/sl_compare.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import bench
from uaisa_env.vec_env import VecMdin
desc, scn, E = bench.workload("c4")
T = 4
for force in ((False, True) if "--both" in sys.argv else (False,)):
    if force:
        os.environ["UAISA_FORCE_SL"] = "1"
    env = VecMdin(scn, E, neighbors_num=10, threads_per_block=32, neighbor_slices=1); env.reset()
    acts = torch.from_numpy(bench.synth_actions(np.random.default_rng(0), T, env.A)).cuda()
    for _ in range(2):
        env.rollout(acts, {}, auto_reset=True)
    torch.cuda.synchronize()
    env.close()

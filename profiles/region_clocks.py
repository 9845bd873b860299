#!/usr/bin/env python
"""Per-region cycle shares of the env kernel (needs a lib built with -DUAISA_REGION_CLOCKS):
   UAISA_B200_LIB=.../lib_rgclk.so python profiles/region_clocks.py [workload] [T]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import bench
from uaisa_env.vec_env import VecMdin
w = sys.argv[1] if len(sys.argv) > 1 else "c5"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
desc, scn, E = bench.workload(w)
env = VecMdin(scn, E, neighbors_num=10); env.reset()
A = env.A
acts = torch.from_numpy(bench.synth_actions(np.random.default_rng(0), T, A)).cuda()
out = {"obs": torch.empty((T, A, env.obs_dim), device="cuda"), "count": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
       "nbr": torch.empty((T, A, 10), dtype=torch.int32, device="cuda"), "reward": torch.empty((T, A), device="cuda"),
       "done": torch.empty((T, A), dtype=torch.uint8, device="cuda"), "info": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
       "finish": torch.empty((T, A), dtype=torch.uint8, device="cuda")}
env.rollout(acts, out, auto_reset=True); torch.cuda.synchronize(); env.region_clocks()
env.rollout(acts, out, auto_reset=True); torch.cuda.synchronize()
rc = env.region_clocks(); tot = sum(rc.values())
nw = (A + 31) // 32
print(desc, "T=%d" % T)
for k, v in rc.items():
    print("%-12s %6.1f%%  %8.0f cycles/warp-step" % (k, 100.0 * v / max(tot, 1), v / nw / T))
print("total %.0f cycles/warp-step" % (tot / nw / T))

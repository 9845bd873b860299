"""Kernel breakdown of one captured PPO rollout step at 65 536 envs x 4 drones (torch.profiler over graph replays):
   python profiles/ppo_rollout_profile.py > gpurun_out/r2/ppo_rollout_profile.txt"""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_env.scenario import scenario_c1
from uaisa_env.vec_env import VecMdin
from uaisa_ppo.trainer import PPOConfig, PPOTrainer
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = VecMdin(scenario_c1(), E)
tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=40))
tr.rollout(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.rollout(); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print("==== rollout of 40 steps, %d agents: %.3f ms per step (device time)" % (env.A, tot / 40e3))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:18]:
    print("%6.1f%%  %8.1f us x %-4d %s" % (100 * e.device_time_total / tot, e.device_time_total / max(e.count, 1), e.count, e.key[:140]))

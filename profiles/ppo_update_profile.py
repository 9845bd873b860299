"""Kernel breakdown of one PPO policy / value iteration at minibatch 65 536 (torch.profiler, eager iterations of PPOTrainer):
   python profiles/ppo_update_profile.py [envs] > gpurun_out/r2/ppo_update_profile.txt"""
import os, sys
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_env.scenario import scenario_c1
from uaisa_env.vec_env import VecMdin
from uaisa_ppo.trainer import PPOConfig, PPOTrainer
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = VecMdin(scenario_c1(), E)
tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=100, graph_update=False))
tr.rollout()
tr._begin_update(); tr._u["cap"] = -1
for _ in range(3):
    tr._pi_iter(); tr._v_iter()
torch.cuda.synchronize()
for name, fn in (("policy iteration", tr._pi_iter), ("value iteration", tr._v_iter)):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in ev)
    print("==== %s: %.3f ms per iteration (device time, 5 iterations)" % (name, tot / 5e3))
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:22]:
        print("%6.1f%%  %8.1f us x %-3d %s" % (100 * e.device_time_total / tot, e.device_time_total / max(e.count, 1), e.count // 5, e.key[:150]))

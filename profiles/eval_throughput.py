#!/usr/bin/env python
"""Throughput of the batched evaluation loop (uaisa_ppo.post_train): episodes/s and env-steps/s with a random-init rnn_ac."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_env.scenario import scenario_c1
from uaisa_env.vec_env import VecMdin
from uaisa_ppo.policy import rnn_ac
from uaisa_ppo.post_train import post_train
from uaisa_ppo.fused import FusedPolicy
torch.manual_seed(0)
ac = rnn_ac().cuda()
for E in (1, 256, 4096):
    env = VecMdin(scenario_c1(), E, neighbors_num=10)
    pt = post_train(env, num_episodes=2 * E, max_ep_len=150)
    pt.policy_test(ac, policy_name="warmup") if E == 1 else None
    fused = FusedPolicy(ac, torch.device('cuda', 0))
    for g in (False, True, 'fused'):
      torch.cuda.synchronize(); t0 = time.perf_counter()
      r = pt.policy_test(fused if g == "fused" else ac, policy_name="random_init_E%d" % E, use_cuda_graph=(g is True))
      torch.cuda.synchronize(); dt = time.perf_counter() - t0
      print("graph=%s E=%d: %d episodes in %.2f s = %.0f episodes/s, %.3g agent-steps/s (%d env steps)" % (g, E, len(r["episodes"]), dt, len(r["episodes"]) / dt, r["env_steps"] * E * env.N / dt, r["env_steps"]))
    env.close()

#!/usr/bin/env python
"""A few PPO update iterations (minibatch 65 536) for ncu: the tcgen05 MLP GEMMs (uaisa_mlp_gemm_kernel), reader forward / backward, clip + Adam.
   ncu --set full --import-source on -k regex:uaisa_mlp_gemm -s 8 -c 1 -o X python profiles/prof_mlp.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_env.scenario import scenario_c1
from uaisa_env.vec_env import VecMdin
from uaisa_ppo.trainer import PPOConfig, PPOTrainer
env = VecMdin(scenario_c1(), 8192)
tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=40, graph_update=False))
tr.rollout()
tr._begin_update(); tr._u["cap"] = -1
for _ in range(3):
    tr._pi_iter(); tr._v_iter()
torch.cuda.synchronize()

#!/usr/bin/env python
"""Kernel-time breakdown of ONE eager PPO policy iteration and ONE value iteration (torch.profiler) at the bench's minibatch size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_ppo.policy import rnn_ac
from torch.profiler import profile, ProfilerActivity
dev = "cuda"; mb = 65536; n = 1 << 20
torch.manual_seed(0)
ac = rnn_ac().to(dev)
obs = torch.randn(n, 102, device=dev); cnt = torch.zeros(n, dtype=torch.uint8, device=dev)
act = torch.randn(n, 3, device=dev); logp_old = torch.randn(n, device=dev); adv = torch.randn(n, device=dev); ret = torch.randn(n, device=dev)
pi_opt = torch.optim.Adam(ac.pi.parameters(), lr=3e-4, capturable=True); vf_opt = torch.optim.Adam(ac.v.parameters(), lr=1e-3, capturable=True)
def pi_it():
    idx = torch.randint(0, n, (mb,), device=dev)
    pi_opt.zero_grad(set_to_none=True)
    _, _, logp = ac.evaluate(obs[idx], cnt[idx], act[idx], multi_cap=-1, heads="pi")
    ratio = torch.exp(logp - logp_old[idx]); a_ = adv[idx]
    loss = -torch.min(ratio * a_, torch.clamp(ratio, 0.8, 1.2) * a_).mean()
    loss.backward(); torch.nn.utils.clip_grad_norm_(list(ac.pi.parameters()), 0.5); pi_opt.step()
def v_it():
    idx = torch.randint(0, n, (mb,), device=dev)
    vf_opt.zero_grad(set_to_none=True)
    _, v, _ = ac.evaluate(obs[idx], cnt[idx], None, multi_cap=-1, heads="v")
    ((v - ret[idx]) ** 2).mean().backward(); vf_opt.step()
for _ in range(3):
    pi_it(); v_it()
torch.cuda.synchronize()
for name, fn in (("policy iteration", pi_it), ("value iteration", v_it)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn(); torch.cuda.synchronize()
    print("====", name)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))

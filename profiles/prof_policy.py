"""One forward of the full fused policy kernel on a fixed batch (what the ncu captures of uaisa_policy_forward_kernel profile).
usage: prof_policy.py [sparse|mixed|dense] [A]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_ppo.policy import rnn_ac
from uaisa_ppo.fused import FusedPolicy
kind = sys.argv[1] if len(sys.argv) > 1 else "sparse"
A = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
torch.manual_seed(0)
m = rnn_ac().cuda(); ff = FusedPolicy(m, "cuda", kernel="full")
g = torch.Generator(device="cuda").manual_seed(1)
obs = torch.zeros((A, 102), device="cuda"); obs[:, :12] = torch.randn((A, 12), device="cuda", generator=g) * 5
cnt = {"sparse": (torch.rand(A, device="cuda", generator=g) < 0.05).to(torch.uint8), "mixed": torch.randint(0, 5, (A,), device="cuda", generator=g).to(torch.uint8),
       "dense": torch.randint(0, 11, (A,), device="cuda", generator=g).to(torch.uint8)}[kind]
rows = torch.randn((A, 10, 9), device="cuda", generator=g) * 2
obs[:, 12:] = (rows * (torch.arange(10, device="cuda")[None, :] < cnt[:, None]).float()[:, :, None]).reshape(A, 90)
for _ in range(4):
    mu, v, _ = ff.forward(obs, cnt)
torch.cuda.synchronize()
print(kind, A, float(mu.abs().mean()), float(v.mean()))

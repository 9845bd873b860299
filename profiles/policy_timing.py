"""Device time of the rollout forward: torch module vs fused kernels (K6a + K6b tcgen05). usage: policy_timing.py [A]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_ppo.policy import rnn_ac
from uaisa_ppo.fused import FusedPolicy
A = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
torch.manual_seed(0)
m = rnn_ac().cuda(); fp = FusedPolicy(m, "cuda")
obs = torch.zeros((A, 102), device="cuda"); obs[:, :12] = torch.randn((A, 12), device="cuda") * 5
cnt = torch.zeros(A, dtype=torch.uint8, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        print("A=%d torch step (tf32=%s): %.3f ms" % (A, tf32, timeit(lambda: m.step(obs, cnt, 1.0, multi_cap=0))))
    feat = fp.features(obs)
    print("fused features K6a: %.3f ms" % timeit(lambda: fp.features(obs)))
    print("fused heads K6b (tcgen05): %.3f ms" % timeit(lambda: fp.heads(feat)))
    print("fused step (features + heads + sampling): %.3f ms" % timeit(lambda: fp.step(obs, cnt, 1.0, multi_cap=0)))
    flops = A * 2 * (288 * 256 + 256 * 256) * 2
    print("heads: %.1f GFLOP" % (flops / 1e9))

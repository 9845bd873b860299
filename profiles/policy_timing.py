"""Device time of the rollout forward: torch module vs the split kernels (K6a + K6b) vs the full fused kernel (uaisa_policy_fwd.cu),
for a sparse batch (every agent <= 1 neighbour row: no recurrence) and dense batches (rows up to nm).  usage: policy_timing.py [A]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_ppo.policy import rnn_ac
from uaisa_ppo.fused import FusedPolicy
A = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
m = rnn_ac().cuda(); fs = FusedPolicy(m, "cuda", kernel="split"); ff = FusedPolicy(m, "cuda", kernel="full")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def batch(kind):
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.zeros((A, 102), device="cuda"); obs[:, :12] = torch.randn((A, 12), device="cuda", generator=g) * 5
    if kind == "empty":        # no neighbour rows at all: what the rollout of configs[1] / configs[4] feeds the policy almost every step
        cnt = torch.zeros(A, dtype=torch.uint8, device="cuda")
    elif kind == "sparse":
        cnt = (torch.rand(A, device="cuda", generator=g) < 0.05).to(torch.uint8)
    elif kind == "mixed":      # configs[2]-like: most agents a few rows
        cnt = torch.randint(0, 5, (A,), device="cuda", generator=g).to(torch.uint8)
    else:
        cnt = torch.randint(0, 11, (A,), device="cuda", generator=g).to(torch.uint8)
    rows = torch.randn((A, 10, 9), device="cuda", generator=g) * 2
    mask = (torch.arange(10, device="cuda")[None, :] < cnt[:, None]).float()[:, :, None]
    obs[:, 12:] = (rows * mask).reshape(A, 90)
    return obs, cnt
with torch.no_grad():
    for kind in os.environ.get("KINDS", "empty,sparse,mixed,dense").split(","):
        obs, cnt = batch(kind)
        mean_rows = float(cnt.float().clamp(min=1).mean()); lmax_tile = float(cnt.view(-1, 128).max(dim=1).values.float().clamp(min=1).mean())
        print("---- %s: A=%d, mean steps per agent %.2f, mean steps per 128-agent tile %.2f" % (kind, A, mean_rows, lmax_tile))
        torch.backends.cuda.matmul.allow_tf32 = False
        if not os.environ.get("FUSED_ONLY"):
            print("torch module step (fp32 cuBLAS):          %.3f ms" % timeit(lambda: m.step(obs, cnt, 1.0, multi_cap=0), 5))
        if kind == "sparse" and not os.environ.get("FUSED_ONLY"):
            print("split kernels K6a + K6b + torch patch:    %.3f ms" % timeit(lambda: fs.step(obs, cnt, 1.0, multi_cap=0), 10))
        t = timeit(lambda: ff.forward(obs, cnt))
        # tensor-core flops actually issued: per tile, GRU x-part 2 dirs x L x 4 slices x (128x192x16), h-part (L-1) x 2 x 4 x (128x192x256), heads 128 x (2 x 272 x 256 + 2 x 256 x 256)
        L = cnt.view(-1, 128).max(dim=1).values.float()
        tiles = L.numel()
        fl = float((2 * L.clamp(max=1) * 4 * 2 * 128 * 192 * 16 + torch.where(L > 1, 2 * (L - 1), torch.zeros_like(L)) * 4 * 2 * 128 * 192 * 16).sum() + (torch.where(L > 1, 2 * (L - 1), torch.zeros_like(L)) * 4 * 2 * 128 * 192 * 256).sum()) + tiles * 2.0 * 128 * (2 * 272 * 256 + 2 * 256 * 256)
        print("full fused kernel (mu, v):                %.3f ms   %.1f GFLOP issued on tcgen05 -> %.1f TFLOP/s TF32" % (t, fl / 1e9, fl / t / 1e9))
        print("full fused step (+ sampling, logp):       %.3f ms" % timeit(lambda: ff.step(obs, cnt, 1.0)))

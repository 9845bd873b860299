import sys, os
sys.path[:0]=['/root/repo','/root/repo/3drvo-marl-collisionavoidance_b200']
import torch
from uaisa_env.scenario import scenario_world3
from uaisa_env.vec_env import VecMdin
from uaisa_ppo.policy import rnn_ac
env=VecMdin(scenario_world3(), 64); env.reset()
ac=rnn_ac().cuda()
act=torch.zeros((env.A,3),device='cuda')
t=torch.zeros(1,dtype=torch.long,device='cuda')
buf=torch.zeros((8,env.A),device='cuda')
ac.pi.rnn_reader.overflow=torch.zeros((),dtype=torch.int64,device='cuda')
def try_capture(name, fn):
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name,'OK')
    except Exception as e:
        print(name,'FAIL',str(e)[:100])
        torch.cuda.synchronize()
from torch.distributions.normal import Normal
with torch.no_grad():
    rd=ac.pi.rnn_reader
    try_capture('topk', lambda: torch.topk(env.count.to(torch.int32), 32))
    try_capture('features cap', lambda: ac.features(env.obs, env.count, 32))
    try_capture('features nocap-lite', lambda: rd._cell(env.obs[:, 12:21], env.obs.new_zeros((env.A,256)), False))
    feat=ac.features(env.obs, env.count, 32)
    try_capture('heads', lambda: (ac.pi.net_out(feat), ac.v.v_net(feat)))
    mu=ac.pi.net_out(feat)
    try_capture('normal sample', lambda: Normal(mu, torch.ones_like(mu), validate_args=False).sample())
    try_capture('overflow', lambda: rd.overflow.add_(((env.count > 1).sum()).to(torch.int64)))
    try_capture('ln', lambda: rd.ln(feat))

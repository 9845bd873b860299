"""-m gpu: the batched evaluation loop (uaisa_ppo.post_train) against results RECORDED FROM THE REFERENCE's own
post_train.policy_test + env (tests/golden/eval_policy_test.json, oracle/gen_eval_golden.py), same deterministic policy."""
import json
import math
import os
import re

import pytest

import _util
from _util import NoisyGoalPolicyTorch, eval_noise_table

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_policy_test.json")))
EP = re.compile(r"(Successful|Fail), Episode (\d+) \t EpRet (\S+) \t EpLen (\d+) \t EpSpeed\s+(\S+)")


def _scenario(name):
    from uaisa_env.scenario import scenario_c1, scenario_world3
    return scenario_world3() if name.startswith("world_3") else scenario_c1()


def _table(mine, ref):
    return "\n".join("%s | ours len %d arrived %s ret %.3f speed %.3f" % (" ".join(r), m["ep_len"], m["arrived"], m["ep_ret"], m["speed"]) for m, r in zip(mine, ref))


def _stats(line):
    m = re.search(r"successful rate: ([\d.]+)% average EpLen: (\S+) std length (\S+) average speed: (\S+) std speed (\S+)", line)
    return [float(x) for x in m.groups()]


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("E,graph", [(1, False), (1, True), (4, True)])
def test_policy_test_matches_the_reference_loop(name, E, graph):
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.post_train import post_train

    g = GOLD[name]
    scn = _scenario(name)
    env = VecMdin(scn, E, neighbors_num=10)
    pt = post_train(env, num_episodes=g["num_episodes"] * E, max_ep_len=g["max_ep_len"], acceler_vel=g["acceler_vel"])
    res = pt.policy_test(NoisyGoalPolicyTorch(eval_noise_table(16, 256, scn.N, g["noise_seed"], g["noise_scale"])), policy_name=name, use_cuda_graph=graph)
    ref = [EP.match(e).groups() for e in g["episodes"]]
    assert len(res["episodes"]) == len(ref) * E
    for e in range(E):      # every env replays the reference's episode sequence (same scenario, same policy, same noise rows)
        mine = [ep for ep in res["episodes"] if ep["env"] == e]
        print(_table(mine, ref))
        off = 0             # episodes whose length moved: free-running from an fp32 state can flip a marginal collision by a step
        for ep, (tag, _, ret, ln, sp) in zip(mine, ref):
            assert ep["arrived"] == (tag == "Successful") and abs(ep["ep_len"] - int(ln)) <= 2, (name, e, ep, tag, ln)
            off += ep["ep_len"] != int(ln)
            # returns: identical up to a few reward-bucket flips (3 / 1 / 0.5 / 0 / -4 per step, ir_gym.py:90-100) and the last steps
            assert (math.isinf(float(ret)) and math.isinf(ep["ep_ret"])) or abs(ep["ep_ret"] - float(ret)) <= 0.02 * abs(float(ret)) + 8.0 * (1 + abs(ep["ep_len"] - int(ln))), (ep, ret)
            assert abs(ep["speed"] - float(sp)) <= 0.02
        assert off <= 1, (name, e, off)
    a, b = _stats(res["line"]), _stats(g["line"])
    assert a[0] == b[0] and abs(a[1] - b[1]) <= 0.5 and abs(a[2] - b[2]) <= 0.5 and abs(a[3] - b[3]) <= 0.011 and abs(a[4] - b[4]) <= 0.011, (res["line"], g["line"])
    if E == 1:
        assert res["line"].split("successful rate")[0] == g["line"].split("successful rate")[0]      # same line layout
    env.close()


def test_policy_test_with_the_actor_critic():
    """The rnn_ac path (sampled with std_factor like load_policy/get_action): runs, fills the quota of every env, writes the line."""
    import tempfile
    import torch
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.policy import rnn_ac
    from uaisa_ppo.post_train import post_train

    torch.manual_seed(0)
    env = VecMdin(scenario_c1(), 64, neighbors_num=10)
    d = tempfile.mkdtemp()
    ac = rnn_ac().cuda()
    res = post_train(env, num_episodes=128, max_ep_len=30).policy_test(ac, policy_name="random_init", result_path=d, use_cuda_graph=True)
    res2 = post_train(env, num_episodes=128, max_ep_len=30).policy_test(ac, policy_name="random_init", use_cuda_graph=False)
    assert len(res2["episodes"]) == 128
    from uaisa_ppo.fused import FusedPolicy
    res3 = post_train(env, num_episodes=128, max_ep_len=30).policy_test(FusedPolicy(ac, env.device), policy_name="random_init_fused")
    assert len(res3["episodes"]) == 128 and abs(res3["average_speed"] - res2["average_speed"]) < 0.1      # same policy through the fused forward
    assert len(res["episodes"]) == 128 and all(1 <= ep["ep_len"] <= 30 for ep in res["episodes"])
    assert open(d + "/result.txt").read().startswith("policy_name: random_init  successful rate: ")
    env.close()

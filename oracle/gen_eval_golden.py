"""TEST INFRASTRUCTURE ONLY.  Golden results for the batched evaluation loop, produced by the
UNMODIFIED reference ``post_train.policy_test`` (train/policy/post_train.py:38-128) driving the unmodified reference env in the
build container:

    python oracle/gen_eval_golden.py   ->  tests/golden/eval_policy_test.json

The policy is a deterministic function of the observation, plus a shared noise table, identical on both sides (tests/_util.py: NoisyGoalPolicyNp / NoisyGoalPolicyTorch),
injected in place of ``load_policy`` -- the loop, the env and the result line are the reference's own.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import ref_shim  # noqa: E402
import _util  # noqa: E402,F401
from _util import NoisyGoalPolicyNp, eval_noise_table  # noqa: E402
from uaisa_env.scenario import scenario_c1, scenario_world3  # noqa: E402

CASES = [("c1_yaml4", scenario_c1, 5, 120, 1.0, 1, 0.1), ("c1_yaml4_len60", scenario_c1, 3, 60, 1.0, 4, 0.1), ("world_3", scenario_world3, 8, 40, 1.0, 2, 0.05), ("c1_yaml4_acc05", scenario_c1, 4, 60, 0.5, 3, 0.1)]     # name, scenario, episodes, max_ep_len, acceler_vel, noise seed, noise scale


def main():
    ref_shim._install_stubs()
    import torch
    torch.cuda.synchronize = lambda *a, **k: None
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "train"))
    from policy.post_train import post_train as ref_post_train
    out = {}
    for name, mk, n_ep, max_len, acc, seed, scale in CASES:
        scn = mk()
        env = ref_shim.make_reference_env(scn.to_dict(0), nm=10)
        pt = ref_post_train(env, num_episodes=n_ep, max_ep_len=max_len, acceler_vel=acc, inf_print=True)
        pol = NoisyGoalPolicyNp(eval_noise_table(16, 256, scn.N, seed, scale), env)
        pt.load_policy = lambda *a, **k: pol
        d = tempfile.mkdtemp(prefix="uaisa_eval_")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            pt.policy_test("drl", None, policy_name=name, result_path=d, result_name="/result.txt")
        line = open(os.path.join(d, "result.txt")).read().strip()
        eps = [ln for ln in buf.getvalue().splitlines() if "Episode" in ln]
        out[name] = {"num_episodes": n_ep, "max_ep_len": max_len, "acceler_vel": acc, "noise_seed": seed, "noise_scale": scale, "line": line, "episodes": eps}
        print(name, "|", line)
        for e in eps:
            print("   ", e)
    with open(os.path.join(ROOT, "tests", "golden", "eval_policy_test.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the batched policy (SURVEY.md section 8a row a-P1), produced by the
UNMODIFIED reference ``rnn_ac`` (train/policy/policy_rnn_ac.py) in the build container:

    python oracle/gen_policy_golden.py   ->  tests/golden/policy_rnn_ac.npz

The weights are not shipped: both sides build ``uaisa_ppo.policy.rnn_ac`` under ``torch.manual_seed(1234)`` (CPU
generator, same torch build on both boxes) and the reference module loads that state_dict with strict=True.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import ref_shim  # noqa: E402

SEED = 1234


def make_inputs(nm=10, B=96, seed=0):
    rng = np.random.default_rng(seed)
    obs = np.zeros((B, 12 + 9 * nm), np.float32)
    cnt = rng.choice([0, 0, 0, 1, 1, 2, 3, 5, 10], B).astype(np.int64)
    obs[:, :12] = np.round(rng.normal(0, 3, (B, 12)), 2)
    for b in range(B):
        obs[b, 12:12 + 9 * cnt[b]] = np.round(rng.normal(0, 2, 9 * cnt[b]), 2)
    act = np.round(rng.uniform(-1, 1, (B, 3)), 2).astype(np.float32)
    return obs, cnt, act


def main():
    ref_shim._install_stubs()
    torch.cuda.synchronize = lambda *a, **k: None          # policy_rnn_ac.py:39 calls it unconditionally
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "train"))
    from gym.spaces import Box
    from policy.policy_rnn_ac import rnn_ac as ref_rnn_ac
    from uaisa_ppo.policy import rnn_ac

    torch.manual_seed(SEED)
    mine = rnn_ac()
    ref = ref_rnn_ac(Box(-1, 1, shape=(21,)), Box(np.array([-1, -1, -1.0]), np.array([1, 1, 1.0])), 12, 9, 256, (256, 256), (256, 256),
                     nn.ReLU, nn.Tanh, nn.Identity, False, 'biGRU', 0)
    ref.load_state_dict(mine.state_dict(), strict=True)
    obs, cnt, act = make_inputs()
    lst = [torch.from_numpy(obs[b, :12 + 9 * max(int(cnt[b]), 1)]) for b in range(obs.shape[0])]
    with torch.no_grad():
        pi, logp = ref.pi(lst, torch.from_numpy(act))     # Actor.forward, policy_rnn_ac.py:178-188 (batched list path)
        v = ref.v(lst)
        mu1 = torch.stack([ref.pi._distribution(o).mean for o in lst])   # single-observation path, :102-127
    assert (mu1 - pi.mean).abs().max() < 1e-5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "policy_rnn_ac.npz"), obs=obs, count=cnt, act=act, mu=pi.mean.numpy(),
                        std=pi.stddev.numpy(), logp=logp.numpy(), v=v.numpy(), seed=np.int64(SEED))
    print("policy golden written: B=%d, max|mu|=%.3f" % (obs.shape[0], float(pi.mean.abs().max())))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the PPO update (SURVEY.md section 8a row a-P3), produced by the UNMODIFIED
reference ``multi_ppo`` (train/policy/multi_ppo.py:341-404: compute_loss_pi / compute_loss_v / update) in the build container:

    python oracle/gen_ppo_golden.py   ->  tests/golden/policy_ppo_update.npz

A fixed buffer of B = 300 samples (one robot's epoch: ragged observations with 0..5 neighbour rows, 2-dp actions, advantages,
returns, old log-probs) goes through the reference's losses and through ONE ``update`` (3 policy + 3 value full-batch
iterations, clip_grad_norm_ 2.0, Adam).  Recorded: both losses, the approx-KL, the gradient of the first policy / value
iteration (per-tensor L2 norms + a strided sample) and the parameters after the update (same digest).  The weights are not
shipped: both sides build ``uaisa_ppo.policy.rnn_ac`` under ``torch.manual_seed(1234)`` (CPU generator).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
import ref_shim  # noqa: E402

SEED = 1234
PI_LR, VF_LR, PI_ITERS, V_ITERS, TARGET_KL = 1e-3, 1e-3, 3, 3, 0.5
STRIDE = 97      # digest: every 97th element of every tensor


def make_buffer(nm=10, B=300, seed=5):
    rng = np.random.default_rng(seed)
    obs = np.zeros((B, 12 + 9 * nm), np.float32)
    cnt = rng.choice([0, 0, 0, 0, 1, 1, 2, 3, 5], B).astype(np.int64)
    obs[:, :12] = np.round(rng.normal(0, 3, (B, 12)), 2)
    for b in range(B):
        obs[b, 12:12 + 9 * cnt[b]] = np.round(rng.normal(0, 2, 9 * cnt[b]), 2)
    act = np.round(rng.uniform(-1, 1, (B, 3)), 2).astype(np.float32)
    adv = rng.normal(0, 2, B).astype(np.float32)
    ret = rng.normal(-3, 5, B).astype(np.float32)
    return obs, cnt, act, adv, ret


def digest(named):
    """{name: tensor} -> (names, per-tensor L2 norms, concatenated strided samples)"""
    names, norms, samp = [], [], []
    for k, t in named:
        a = t.detach().reshape(-1).double().numpy()
        names.append(k); norms.append(np.sqrt((a * a).sum())); samp.append(a[::STRIDE])
    return names, np.array(norms), np.concatenate(samp)


def main():
    ref_shim._install_stubs()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "train"))
    from gym.spaces import Box
    from policy.multi_ppo import multi_ppo
    from policy.policy_rnn_ac import rnn_ac as ref_rnn_ac
    from uaisa_env.scenario import scenario_c1
    from uaisa_ppo.policy import rnn_ac

    torch.manual_seed(SEED)
    mine = rnn_ac()
    ref = ref_rnn_ac(Box(-1, 1, shape=(21,)), Box(np.array([-1, -1, -1.0]), np.array([1, 1, 1.0])), 12, 9, 256, (256, 256), (256, 256),
                     nn.ReLU, nn.Tanh, nn.Identity, False, 'biGRU', 0)
    ref.load_state_dict(mine.state_dict(), strict=True)
    obs, cnt, act, adv, ret = make_buffer()
    lst = [torch.from_numpy(obs[b, :12 + 9 * max(int(cnt[b]), 1)].copy()) for b in range(obs.shape[0])]
    with torch.no_grad():                     # old log-probs: the initial policy's, shifted a little so that ratio != 1 and KL != 0
        _, logp0 = ref.pi(lst, torch.from_numpy(act))
    rng = np.random.default_rng(9)
    logp_old = (logp0.numpy() + rng.normal(0, 0.05, logp0.shape[0])).astype(np.float32)
    data = dict(obs=lst, act=torch.from_numpy(act), ret=torch.from_numpy(ret), adv=torch.from_numpy(adv), logp=torch.from_numpy(logp_old))

    env = ref_shim.make_reference_env(scenario_c1().to_dict(0), nm=10)
    ppo = multi_ppo(env, ref, PI_LR, VF_LR, 1, 300, 500, 0.99, 0.97, 0.2, PI_ITERS, V_ITERS, TARGET_KL, False, 5, False, 7, 10 ** 9, False,
                    "/tmp/uaisa_ppo_golden/", "r", None, False, False, 0, env, 1400, 10, False, None)
    ppo.robot_num = 1
    # ---- the two losses and their first gradients (multi_ppo.py:379-404)
    named = [(k, p) for k, p in ref.named_parameters()]
    ppo.pi_optimizer.zero_grad(); ppo.vf_optimizer.zero_grad()
    loss_pi, info = ppo.compute_loss_pi(data)
    loss_pi.backward()
    gpi_names, gpi_norm, gpi_samp = digest([(k, p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in named])
    ppo.pi_optimizer.zero_grad(); ppo.vf_optimizer.zero_grad()
    loss_v = ppo.compute_loss_v(data)
    loss_v.backward()
    _, gv_norm, gv_samp = digest([(k, p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in named])
    for p in ref.parameters():
        p.grad = None
    # ---- one update (multi_ppo.py:341-376)
    with contextlib.redirect_stdout(io.StringIO()):
        ppo.update([data])
    names, p_norm, p_samp = digest(named)
    with torch.no_grad():
        _, logp1 = ref.pi(lst, torch.from_numpy(act))
        v1 = ref.v(lst)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "policy_ppo_update.npz"), obs=obs, count=cnt, act=act, adv=adv, ret=ret, logp_old=logp_old,
                        loss_pi=np.float64(loss_pi.item()), kl=np.float64(info["kl"]), loss_v=np.float64(loss_v.item()),
                        names=np.array(names), grad_pi_norm=gpi_norm, grad_pi_sample=gpi_samp, grad_v_norm=gv_norm, grad_v_sample=gv_samp,
                        param_norm=p_norm, param_sample=p_samp, logp_after=logp1.numpy(), v_after=v1.numpy(), seed=np.int64(SEED),
                        hyper=np.array([PI_LR, VF_LR, PI_ITERS, V_ITERS, TARGET_KL, STRIDE], np.float64))
    print("ppo golden: loss_pi %.6f kl %.6f loss_v %.6f |grad_pi| %.4f" % (loss_pi.item(), info["kl"], loss_v.item(), float(np.sqrt((gpi_norm ** 2).sum()))))


if __name__ == "__main__":
    main()

"""-m gpu: rollout-side kernels (action transform, GAE) against the reference formulas, and a PPO epoch end to end."""
import numpy as np
import pytest
import torch

import _util  # noqa: F401

pytestmark = pytest.mark.gpu


def ref_finish_path(rew, val, gamma, lam):
    """multi_PPObuf.finish_path with last_val = 0 (multi_ppo.py:68-77), discount_cumsum written out (:21-36)."""
    rews = np.append(rew, 0.0); vals = np.append(val, 0.0)
    deltas = rews[:-1] + gamma * vals[1:] - vals[:-1]

    def dc(x, d):
        out = np.zeros_like(x); run = 0.0
        for i in range(len(x) - 1, -1, -1):
            run = x[i] + d * run; out[i] = run
        return out
    return dc(deltas, gamma * lam), dc(rews, gamma)[:-1]


def test_gae_kernel_matches_finish_path():
    from uaisa_ppo.trainer import gae

    rng = np.random.default_rng(0)
    T, A = 57, 300
    rew = rng.normal(0, 3, (T, A)).astype(np.float32); val = rng.normal(0, 2, (T, A)).astype(np.float32)
    cut = (rng.uniform(0, 1, (T, A)) < 0.08).astype(np.uint8)
    adv, ret = gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(cut).cuda(), 0.99, 0.97)
    adv, ret = adv.cpu().numpy(), ret.cpu().numpy()
    for a in range(0, A, 17):
        start = 0
        for t in range(T):
            if cut[t, a] or t == T - 1:
                ra, rr = ref_finish_path(rew[start:t + 1, a].astype(np.float64), val[start:t + 1, a].astype(np.float64), 0.99, 0.97)
                assert np.allclose(adv[start:t + 1, a], ra, rtol=2e-5, atol=2e-5) and np.allclose(ret[start:t + 1, a], rr, rtol=2e-5, atol=2e-5)
                start = t + 1


def test_action_transform_matches_trainer_formula():
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin, unpack_state

    env = VecMdin(scenario_c1(), 50); env.reset()
    rng = np.random.default_rng(1)
    for _ in range(5):
        env.step(torch.from_numpy(np.round(rng.uniform(-1, 1, (env.A, 3)), 2).astype(np.float32)).cuda())
    a_inc = rng.normal(0, 0.6, (env.A, 3)).astype(np.float32)
    action, a_r = env.action_transform(torch.from_numpy(a_inc).cuda())
    vel = unpack_state(env.get_state())["vel"]
    ar = np.round(a_inc, 2)                                        # float32 array rounding, multi_ppo.py:197
    ab = np.round(0.5 * ar + vel, 2)                               # :203-205 (float64)
    assert np.array_equal(a_r.cpu().numpy(), ar)
    assert np.array_equal(action.cpu().numpy(), ab.astype(np.float32))
    env.close()


def test_ppo_epoch_runs_and_learns_something():
    from uaisa_env.scenario import scenario_world3
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer
    from uaisa_env.vec_env import VecMdin

    env = VecMdin(scenario_world3(), 64)
    tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=24, train_pi_iters=4, train_v_iters=4, minibatch=2048, vf_lr=1e-3, pi_lr=3e-4))
    before = {k: v.clone() for k, v in tr.ac.state_dict().items()}
    s1 = tr.train_epoch()
    assert np.isfinite(s1["loss_v"]) and s1["pi_iters"] >= 1
    changed = [k for k, v in tr.ac.state_dict().items() if not torch.equal(v, before[k])]
    assert any(k.startswith("pi.net_out") for k in changed) and any(k.startswith("v.v_net") for k in changed) and any("rnn_reader" in k for k in changed)
    lv = [tr.train_epoch()["loss_v"] for _ in range(3)]
    assert all(np.isfinite(x) for x in lv)
    # buffer bookkeeping: every path is cut at the epoch end (multi_ppo.py:244-264)
    assert bool(tr.buf["cut"][-1].all())
    import io
    f = io.BytesIO(); tr.save_model(f); f.seek(0)
    ck = torch.load(f, map_location="cpu", weights_only=False)
    assert set(ck) == {"model_state", "pi_optimizer", "vf_optimizer"}
    env.close()

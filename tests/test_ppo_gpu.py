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


def _digest(named, stride):
    norms, samp = [], []
    for _, t in named:
        a = t.detach().reshape(-1).double().cpu().numpy()
        norms.append(np.sqrt((a * a).sum())); samp.append(a[::stride])
    return np.array(norms), np.concatenate(samp)


def _golden_trainer(z, graph_update=False):
    """PPOTrainer over a 1-env x 4-drone VecMdin whose buffer is overwritten with the golden's 300 samples (T = 75, A = 4)."""
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.policy import rnn_ac
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer

    pi_lr, vf_lr, pi_it, v_it, tkl, _ = z["hyper"]
    torch.manual_seed(int(z["seed"]))
    ac = rnn_ac()                                   # CPU generator: the weights the golden was recorded with
    env = VecMdin(scenario_c1(), 1)
    cfg = PPOConfig(steps_per_epoch=75, pi_lr=float(pi_lr), vf_lr=float(vf_lr), train_pi_iters=int(pi_it), train_v_iters=int(v_it), target_kl=float(tkl),
                    minibatch=300, use_fused_policy=False, use_cuda_graph=False, graph_update=graph_update,
                    tc_mlp=False)        # fp32 MLPs: the golden is the reference's fp32 arithmetic (the tcgen05 path has its own tests, test_mlp_tc_gpu.py)
    tr = PPOTrainer(env, ac=ac, cfg=cfg)
    dev = tr.dev
    tr.buf["obs"].view(300, -1).copy_(torch.from_numpy(z["obs"]).to(dev)); tr.buf["count"].view(300).copy_(torch.from_numpy(z["count"].astype(np.uint8)).to(dev))
    tr.buf["act"].view(300, 3).copy_(torch.from_numpy(z["act"]).to(dev)); tr.buf["logp"].view(300).copy_(torch.from_numpy(z["logp_old"]).to(dev))
    adv, ret = torch.from_numpy(z["adv"]).to(dev), torch.from_numpy(z["ret"]).to(dev)
    return tr, env, adv, ret


def test_update_matches_reference_multi_ppo_golden():
    """Losses, approx-KL, first gradients and the parameters after ONE reference update (multi_ppo.py:341-404: 3 + 3 full-batch iterations,
    clip_grad_norm_ 2.0 -- active here, |g| = 3.85 -- and Adam) on a fixed 300-sample buffer with ragged observations, recorded from the
    unmodified reference by oracle/gen_ppo_golden.py."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_ppo_update.npz"))
    stride = int(z["hyper"][5])
    tr, env, adv, ret = _golden_trainer(z)
    named = [(k, dict(tr.ac.named_parameters())[str(k)]) for k in z["names"]]
    F = tr.flat
    tr._begin_update(adv, ret)
    tr._u["cap"] = 0
    tr._pi_fb()
    assert abs(float(F.g[F.n + 1]) - float(z["loss_pi"])) <= 2e-5 * max(1.0, abs(float(z["loss_pi"])))
    assert abs(float(F.g[F.n]) - float(z["kl"])) <= 2e-6
    gn, gs = _digest([(k, p.grad) for k, p in named], stride)
    assert np.allclose(gn, z["grad_pi_norm"], rtol=2e-4, atol=1e-6) and np.allclose(gs, z["grad_pi_sample"], rtol=2e-3, atol=2e-6)
    F.g.zero_()
    tr._v_fb()
    assert abs(float(tr._u["loss_v"]) - float(z["loss_v"])) <= 2e-5 * abs(float(z["loss_v"]))
    gn, gs = _digest([(k, p.grad) for k, p in named], stride)
    assert np.allclose(gn, z["grad_v_norm"], rtol=2e-4, atol=1e-6) and np.allclose(gs, z["grad_v_sample"], rtol=2e-3, atol=2e-5)
    F.g.zero_()
    st = tr.update(adv, ret)
    assert st["pi_iters"] == int(z["hyper"][2]) and not st["early_stop"]
    pn, ps = _digest(named, stride)
    lr = float(z["hyper"][0])
    # Adam moves every parameter by ~lr per iteration whatever the gradient's size, so elements whose gradient is at rounding-noise level may
    # differ by a fraction of lr between fp32 GPU and fp32 CPU arithmetic; everything else agrees to 1e-5
    assert np.allclose(pn, z["param_norm"], rtol=2e-5)
    d = np.abs(ps - z["param_sample"])
    assert np.mean(d <= 0.02 * lr) >= 0.98 and d.max() <= 6.5 * lr, (np.mean(d <= 0.02 * lr), d.max())
    with torch.no_grad():
        _, v1, logp1 = tr.ac.evaluate(tr._obs, tr._cnt, tr._act)
    assert np.allclose(logp1.cpu().numpy(), z["logp_after"], rtol=1e-3, atol=2e-3) and np.allclose(v1.cpu().numpy(), z["v_after"], rtol=1e-3, atol=2e-3)
    env.close()


def test_graphed_update_equals_eager_update():
    """The CUDA-graph iterations and the eager iterations apply the same rule (same clip over ALL parameters, same Adam, the warm-up before
    the capture leaves no trace): identical buffer + seed -> identical parameters."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_ppo_update.npz"))
    z = dict(z)
    z["count"] = np.minimum(z["count"], 1)            # (the graphed path is taken when no sample has >= 2 rows)
    outs = []
    for graphed in (False, True):
        tr, env, adv, ret = _golden_trainer(z, graph_update=graphed)
        torch.manual_seed(99)
        st = tr.update(adv, ret)
        assert ("graph" in tr.update_path.split(";")[0].lower()) == graphed, tr.update_path
        st2 = tr.update(adv, ret)
        outs.append((tr.flat.p.clone(), tr.pi_opt.step.item(), tr.vf_opt.step.item(), st["pi_iters"], st2["pi_iters"]))
        env.close()
    assert outs[0][1:] == outs[1][1:] == (6, 6, 3, 3)
    # Same rule -> same parameters.  The reader backward accumulates its CTAs' partial sums with float atomics, so a gradient carries rounding noise of ~1e-7 of
    # its TERMS' size that depends on the order the CTAs retire in; the value head's terms are large here (returns ~ -500), so an element whose terms cancel can
    # change sign between two runs of the SAME path, and Adam moves such an element by a whole +-lr per iteration whatever the gradient's size.  Seen once in
    # ~10 full-suite runs (one element in a few hundred off by 0.9 lr).  A difference in the rule -- clip threshold, Adam state, a trace of the warm-up iterations --
    # moves nearly EVERY element by >= 1e-5, so: all but a handful of elements equal to 2e-6, none further apart than the iterations' total reach.
    d = (outs[0][0] - outs[1][0]).abs()
    lr_max = max(float(z["hyper"][0]), float(z["hyper"][1]))
    frac = float((d <= 2e-6).float().mean())
    assert frac >= 0.998 and float(d.max()) <= 6.5 * lr_max, (float(d.max()), frac, int((d > 2e-6).sum()), torch.nonzero(d > 2e-6).flatten()[:8].tolist())


def test_clip_adam_kernel_matches_torch_and_early_stop_latches():
    """uaisa_opt_prepare + uaisa_clip_adam against clip_grad_norm_ + torch.optim.Adam on a flat bucket; the KL latch stops later iterations."""
    import ctypes as C
    from uaisa_env import _capi
    L = _capi.lib()
    dev = torch.device("cuda")
    g0 = torch.Generator(device="cpu").manual_seed(3)
    n, lo, cnt = 100003, 1000, 90000
    p = torch.randn(n, generator=g0).to(dev); ref_p = p.clone().requires_grad_(True)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev); ctl = torch.zeros(2, dtype=torch.int32, device=dev)
    partials = torch.zeros(L.uaisa_opt_partials_len(), dtype=torch.float64, device=dev)
    kl = torch.zeros(1, device=dev)
    opt = torch.optim.Adam([ref_p], lr=1e-2)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(5):
        g = (torch.randn(n, generator=g0) * (0.02 if it % 2 else 0.001)).to(dev)     # norm 6.3 (clipped) / 0.3 (not clipped)
        ref_p.grad = g.clone(); ref_p.grad[:lo] = 0; ref_p.grad[lo + cnt:] = 0        # the reference optimiser only sees its own range ...
        total = torch.linalg.vector_norm(g)                                            # ... but the clip norm covers the whole bucket
        coef = torch.clamp(2.0 / (total + 1e-6), max=1.0)
        ref_p.grad.mul_(coef)
        before = ref_p.detach().clone()
        opt.step()
        ref_p.data[:lo] = before[:lo]; ref_p.data[lo + cnt:] = before[lo + cnt:]
        assert L.uaisa_opt_prepare(P(g), n, 1, P(partials), P(kl), 0.05, P(ctl), P(step), s) == 0
        assert L.uaisa_clip_adam(P(p), P(g), P(m), P(v), n, lo, cnt, P(partials), 2.0, 1e-2, 0.9, 0.999, 1e-8, P(step), P(ctl), s) == 0
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=2e-6), (it, float((p - ref_p.detach()).abs().max()))
    assert step.item() == 5 and ctl.tolist() == [0, 5]
    kl.fill_(0.06)                                   # KL above the target: latch, nothing is applied any more (multi_ppo.py:362-364)
    snap = p.clone()
    for it in range(2):
        if it == 1:
            kl.fill_(0.0)                            # the latch holds for the rest of the update
        g = torch.randn(n, generator=g0).to(dev)
        assert L.uaisa_opt_prepare(P(g), n, 1, P(partials), P(kl), 0.05, P(ctl), P(step), s) == 0
        assert L.uaisa_clip_adam(P(p), P(g), P(m), P(v), n, lo, cnt, P(partials), 2.0, 1e-2, 0.9, 0.999, 1e-8, P(step), P(ctl), s) == 0
    assert torch.equal(p, snap) and step.item() == 5 and ctl.tolist() == [1, 5]


def test_fused_forward_logp_and_value_are_close_to_the_update_path():
    """The rollout's logp_old / value come from the fused forward (TF32 tcgen05 heads), the update evaluates with torch in fp32: bound the
    mismatch that feeds the PPO ratio and the approx-KL early stop (target_kl 0.05)."""
    from uaisa_env.scenario import scenario_c3
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.fused import FusedPolicy
    from uaisa_ppo.policy import rnn_ac

    env = VecMdin(scenario_c3(seed=3, n_drones=8, n_buildings=2), 256); env.reset()
    rng = np.random.default_rng(0)
    for _ in range(12):
        env.step(torch.from_numpy(np.round(rng.uniform(-1, 1, (env.A, 3)), 2).astype(np.float32)).cuda())
    torch.manual_seed(5)
    ac = rnn_ac().cuda()
    fp = FusedPolicy(ac, env.device)
    torch.manual_seed(11)
    a, v, logp = fp.step(env.obs, env.count)
    with torch.no_grad():
        _, v_t, logp_t = ac.evaluate(env.obs, env.count, a)
    assert float((logp - logp_t).abs().max()) < 2e-2 and float((logp - logp_t).abs().mean()) < 3e-3
    assert float((v - v_t).abs().max()) < 2e-2
    env.close()


def test_fused_reader_forward_backward_matches_torch_autograd():
    """uaisa_reader_fwd / uaisa_reader_bwd (the reader of update minibatches without multi-row samples as one forward and one backward kernel, with the
    minibatch gathered inside the kernel) against the torch ops they replace: features to 1e-5, every parameter gradient to 1e-4 of its scale."""
    from uaisa_ppo.policy import rnn_ac

    torch.manual_seed(3)
    ac = rnn_ac().cuda()
    rd = ac.pi.rnn_reader
    rng = np.random.default_rng(2)
    n, B = 5000, 3111
    obs = np.zeros((n, 102), np.float32)
    obs[:, :12] = np.round(rng.normal(0, 4, (n, 12)), 2); obs[:, 0:3] += 15
    cnt = (rng.uniform(0, 1, n) < 0.4).astype(np.uint8)
    obs[:, 12:21] = np.round(rng.normal(0, 2, (n, 9)), 2) * cnt[:, None]
    o, c = torch.from_numpy(obs).cuda(), torch.from_numpy(cnt).cuda()
    idx = torch.from_numpy(rng.integers(0, n, B)).cuda()
    wgt = torch.from_numpy(rng.normal(0, 1, (B, 268)).astype(np.float32)).cuda()
    res = {}
    for fused in (False, True):
        rd.fused_step0 = fused
        ac.zero_grad(set_to_none=True)
        feat = ac.features(o, c, multi_cap=-1, index=idx)
        assert (feat.grad_fn is not None) and (("ReaderStep0" in type(feat.grad_fn).__name__) == fused)
        (feat * wgt).sum().backward()
        res[fused] = (feat.detach().clone(), {k: p.grad.detach().clone() for k, p in rd.named_parameters() if p.grad is not None})
    rd.fused_step0 = True
    assert torch.allclose(res[True][0], res[False][0], rtol=1e-5, atol=2e-5), float((res[True][0] - res[False][0]).abs().max())
    assert set(res[True][1]) >= set(res[False][1])
    for k, g_ref in res[False][1].items():
        g = res[True][1][k]
        scale = float(g_ref.abs().max())
        assert float((g - g_ref).abs().max()) <= 2e-4 * scale + 1e-6, (k, float((g - g_ref).abs().max()), scale)
    # weight_hh never enters a single step from h0 = 0: no gradient (torch) / zero gradient (kernel)
    for k, g in res[True][1].items():
        if "weight_hh" in k:
            assert float(g.abs().max()) == 0.0
    # without an index: all rows
    f1 = ac.features(o, c, multi_cap=-1)
    rd.fused_step0 = False
    f0 = ac.features(o, c, multi_cap=-1)
    rd.fused_step0 = True
    assert torch.allclose(f1, f0, rtol=1e-5, atol=2e-5)


def test_unrolled_rollout_fills_the_buffers_like_the_per_step_rollout():
    """PPOConfig.unroll_rollout (the whole epoch as one graph, env / policy kernels writing straight into the epoch buffers) against the per-step graph
    with its buffer copies: same seeds -> identical observations, counts, actions, rewards, values, log-probs and path cuts."""
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer
    bufs = []
    for unroll in (True, False):
        env = VecMdin(scenario_c1(), 48, neighbors_num=10)
        tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=21, max_ep_len=9, unroll_rollout=unroll), seed=11)
        assert (tr._single_graph and unroll) == unroll
        torch.manual_seed(5)
        tr.rollout()
        first = {k: v.clone() for k, v in tr.buf.items()}
        tr.rollout()                                    # a second epoch continues from the first one's final observation
        bufs.append((first, {k: v.clone() for k, v in tr.buf.items()}, int(tr.stats["collisions"])))
        assert (tr._graph_unrolled is not None) == unroll
        tr.close(); env.close()
    for ep in (0, 1):
        for k in bufs[0][ep]:
            a, b = bufs[0][ep][k], bufs[1][ep][k]
            assert torch.equal(torch.nan_to_num(a.float(), 7.0), torch.nan_to_num(b.float(), 7.0)), (ep, k)
    assert bufs[0][2] == bufs[1][2] and bufs[0][2] > 0
    assert int(bufs[0][0]["cut"].sum()) > 0


@pytest.mark.parametrize("mask_some", [False, True])
def test_native_update_iteration_matches_the_autograd_iteration(mask_some):
    """PPOConfig.native_update (reader / MLP / loss forward and backward as a fixed sequence of our own kernels, no autograd tape) against the same
    iteration through torch autograd over the same kernels: the whole gradient bucket, the approx-KL, both losses; then parameters after an update."""
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer
    res = []
    for native in (True, False):
        env = VecMdin(scenario_c1(), 40, neighbors_num=10)
        tr = PPOTrainer(env, cfg=PPOConfig(steps_per_epoch=20, max_ep_len=9, train_pi_iters=3, train_v_iters=3, minibatch=1000, pi_lr=1e-3, vf_lr=1e-3, target_kl=10.0,
                                           native_update=native, graph_update=False), seed=5)
        torch.manual_seed(17)
        tr.rollout()
        if mask_some:                                      # inf / nan rewards of the reference (ir_gym.py:88): masked samples
            tr.buf["rew"][3, 5] = float("inf"); tr.buf["rew"][7, 11] = float("nan")
        tr._begin_update(); tr._u["cap"] = -1
        assert tr._native() == native
        F = tr.flat
        F.g.zero_(); tr._pi_fb(); g_pi = F.g.clone()
        tr._u["off"].zero_()
        F.g.zero_(); tr._v_fb(); g_v = F.g.clone(); lv = float(tr._u["loss_v"])
        torch.manual_seed(23)
        p0 = F.p.clone()
        st = tr.update()
        res.append((g_pi, g_v, lv, F.p.clone() - p0, st["pi_iters"]))
        tr.close(); env.close()
    (gp1, gv1, lv1, d1, it1), (gp0, gv0, lv0, d0, it0) = res
    for a, b, name in ((gp1, gp0, "pi gradients + kl + loss"), (gv1, gv0, "v gradients")):
        assert float(b.abs().max()) > 0
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-9, (name, float((a - b).abs().max()), float(b.abs().max()))
    assert abs(lv1 - lv0) <= 1e-5 * max(1.0, abs(lv0)) and it1 == it0 == 3
    cos = float((d1 * d0).sum() / (d1.norm() * d0.norm()))
    assert cos > 0.999, cos

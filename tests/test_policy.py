"""The batched actor-critic (SURVEY 8a row a-P1) against golden outputs of the reference rnn_ac
(oracle/gen_policy_golden.py); CPU everywhere, CUDA with -m gpu; plus state_dict compatibility with shipped checkpoints."""
import os

import numpy as np
import pytest
import torch

import _util  # noqa: F401
from uaisa_ppo.policy import rnn_ac

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_rnn_ac.npz"))


def _model(device):
    torch.manual_seed(int(G["seed"]))
    return rnn_ac().to(device)


def _check(device, tol):
    m = _model(device)
    obs, cnt, act = (torch.from_numpy(G[k]).to(device) for k in ("obs", "count", "act"))
    with torch.no_grad():
        dist, v, logp = m.evaluate(obs, cnt, act)
    assert np.abs(dist.mean.cpu().numpy() - G["mu"]).max() < tol
    assert np.abs(dist.stddev.cpu().numpy() - G["std"]).max() < 1e-6
    assert np.abs(v.cpu().numpy() - G["v"]).max() < tol
    assert np.abs(logp.cpu().numpy() - G["logp"]).max() < 50 * tol
    return m


def test_policy_matches_reference_golden_cpu():
    m = _check("cpu", 2e-5)
    assert sum(p.numel() for p in m.parameters()) == 680991            # SURVEY 8a a-P1
    keys = set(m.state_dict().keys())
    for k in ("pi.log_std", "pi.rnn_reader.rnn_net.weight_ih_l0_reverse", "pi.rnn_reader.ln.weight", "pi.net_out.4.bias",
              "v.v_net.0.weight", "v.rnn_reader.rnn_net.bias_hh_l0"):
        assert k in keys


def test_dense_batching_equals_per_agent_evaluation():
    """The compaction of multi-row agents must not change anybody's result."""
    m = _model("cpu")
    obs, cnt = torch.from_numpy(G["obs"]), torch.from_numpy(G["count"])
    with torch.no_grad():
        full = m.features(obs, cnt)
        for b in range(0, obs.shape[0], 7):
            one = m.features(obs[b:b + 1], cnt[b:b + 1])
            assert torch.allclose(one[0], full[b], atol=1e-5)


def test_static_capacity_path_is_exact_when_it_fits_and_counts_overflow():
    m = _model("cpu")
    obs, cnt = torch.from_numpy(G["obs"]), torch.from_numpy(G["count"])
    n_multi = int((cnt > 1).sum())
    rd = m.pi.rnn_reader
    with torch.no_grad():
        ref = m.features(obs, cnt)
        rd.overflow = torch.zeros((), dtype=torch.int64)
        assert torch.allclose(m.features(obs, cnt, multi_cap=n_multi + 3), ref, atol=1e-6) and int(rd.overflow) == 0
        m.features(obs, cnt, multi_cap=n_multi - 2)
        assert int(rd.overflow) == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/train/model_save"), reason="reference checkpoints not present")
def test_loads_shipped_reference_checkpoint_strict():
    ck = torch.load("/root/reference/train/model_save/r8_0/r8_0_check_point_0.pt", map_location="cpu", weights_only=False)
    m = rnn_ac()
    m.load_state_dict(ck["model_state"], strict=True)


@pytest.mark.gpu
def test_policy_matches_reference_golden_cuda():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    _check("cuda", 5e-5)


@pytest.mark.gpu
def test_fused_features_and_tcgen05_heads_match_torch():
    """K6a / K6b (csrc/uaisa_policy.cu) against the torch module with the same weights: features to fp32 accuracy, heads to
    TF32 accuracy (10-bit mantissa inputs, fp32 accumulation)."""
    from uaisa_ppo.fused import FusedPolicy

    torch.backends.cuda.matmul.allow_tf32 = False
    m = _model("cuda")
    fp = FusedPolicy(m, "cuda", kernel="split")
    rng = np.random.default_rng(3)
    for A in (96, 128, 1000, 4133):
        obs = np.zeros((A, 102), np.float32)
        obs[:, :12] = np.round(rng.normal(0, 5, (A, 12)), 2); obs[:, 0:3] += 20
        cnt = rng.choice([0, 0, 1, 1, 2, 4], A).astype(np.uint8)
        for b in range(A):
            obs[b, 12:12 + 9 * cnt[b]] = np.round(rng.normal(0, 2, 9 * int(cnt[b])), 2)
        o, c = torch.from_numpy(obs).cuda(), torch.from_numpy(cnt).cuda()
        with torch.no_grad():
            ref_feat = m.features(o, c)
            dist, ref_v, _ = m.evaluate(o, c)
        feat = fp.features(o)
        single = (c <= 1)
        assert torch.allclose(feat[single, :268], ref_feat[single], atol=2e-5, rtol=1e-5), (feat[single, :268] - ref_feat[single]).abs().max()
        assert bool((feat[:, 268:] == 0).all())
        featp = feat.clone(); featp[:, :268] = ref_feat
        mu, v = fp.heads(featp)
        torch.cuda.synchronize()
        assert (mu - dist.mean).abs().max() < 5e-3 and (v - ref_v).abs().max() < 5e-3 * max(1.0, float(ref_v.abs().max())), ((mu - dist.mean).abs().max(), (v - ref_v).abs().max())
        a, vv, logp = fp.step(o, c)
        assert a.shape == (A, 3) and torch.isfinite(logp).all() and (vv - ref_v).abs().max() < 5e-3 * max(1.0, float(ref_v.abs().max()))


def test_eval_parity_policy_forms_agree():
    """The numpy form (run inside the reference's policy_test when the golden is made) and the torch form (run by the batched loop)
    of the scripted evaluation policy are the same function: identical float32 outputs on 2-dp observations."""
    import numpy as np
    import torch
    from _util import goal_policy_np, goal_policy_torch
    rng = np.random.default_rng(3)
    obs = np.round(rng.normal(0, 2, (4000, 21)), 2)
    obs[:200, 3:6] = 0.0                      # standing still
    obs[200:300, 8:11] = 0.0                  # no desired velocity (at the waypoint)
    a = np.stack([goal_policy_np(o) for o in obs])
    b = goal_policy_torch(torch.from_numpy(obs.astype(np.float32))).numpy()
    ra, rb = np.round(a, 2), np.round(b, 2)
    assert (ra != rb).mean() == 0 and np.abs(a - b).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("nm", [10, 3, 16])
def test_full_forward_kernel_matches_torch_for_every_row_count(nm):
    """K6 complete (csrc/uaisa_policy_fwd.cu): ONE kernel with the masked bi-GRU recurrence, LayerNorm and both tcgen05 heads against the torch module
    (itself pinned to the reference's rnn_ac by the golden above) for every neighbour-row count 0..nm, tiles with and without multi-row agents, ragged
    last tile.  TF32 operands (10-bit mantissa), fp32 accumulation."""
    from uaisa_ppo.fused import FusedPolicy

    torch.backends.cuda.matmul.allow_tf32 = False
    m = _model("cuda")
    fp = FusedPolicy(m, "cuda")
    rng = np.random.default_rng(7)
    D = 12 + 9 * nm
    for A, choices in ((128, [0, 1]), (256, [0, 0, 1, 2]), (1000, list(range(nm + 1))), (4133, [0, 0, 0, 1, 1, 2, 3, nm]), (77, [nm])):
        obs = np.zeros((A, D), np.float32)
        obs[:, :12] = np.round(rng.normal(0, 5, (A, 12)), 2); obs[:, 0:3] += 20
        cnt = rng.choice(choices, A).astype(np.uint8)
        if A == 4133:
            cnt[:1024] = np.minimum(cnt[:1024], 1)             # the first 8 tiles take the no-recurrence path
        for b in range(A):
            obs[b, 12:12 + 9 * cnt[b]] = np.round(rng.normal(0, 2, 9 * int(cnt[b])), 2)
        o, c = torch.from_numpy(obs).cuda(), torch.from_numpy(cnt).cuda()
        with torch.no_grad():
            ref_feat = m.features(o, c)
            dist, ref_v, _ = m.evaluate(o, c)
        fp.stats.zero_()
        mu, v, feat = fp.forward(o, c, want_features=True)
        torch.cuda.synchronize()
        err_f = (feat - ref_feat).abs().max().item()
        assert err_f < 1.5e-2, (A, nm, err_f)                                  # LayerNorm output is O(1); TF32 recurrence over <= nm steps
        assert (feat - ref_feat).abs().mean().item() < 1e-3
        assert (mu - dist.mean).abs().max().item() < 1e-2 and (v - ref_v).abs().max().item() < 1e-2 * max(1.0, float(ref_v.abs().max())), ((mu - dist.mean).abs().max(), (v - ref_v).abs().max())
        st = fp.stats.tolist()
        # a tile runs the recurrence when it holds a multi-row agent; with UAISA_POLICY_CLUSTER=2 the two tiles of a cluster walk the same
        # weight-tile sequence, so both run it when either does
        span = 128 * max(1, int(os.environ.get("UAISA_POLICY_CLUSTER", "1")))
        tiles_multi = int(sum(len(range(t, min(t + span, A), 128)) for t in range(0, A, span) if int(cnt[t:t + span].max()) > 1))
        assert st[0] == tiles_multi and st[1] == (int(cnt.max()) if tiles_multi else 0), (st, tiles_multi)
        single = torch.from_numpy(cnt <= 1).cuda()
        if bool(single.any()):      # no recurrence for these agents: the only TF32 rounding is in the 9-wide input projection
            assert (feat[single] - ref_feat[single]).abs().max().item() < 4e-3
    # graph capture + in-place refresh after a weight change
    a, vv, logp = fp.step(o, c)
    assert a.shape == (A, 3) and torch.isfinite(logp).all()
    with torch.no_grad():
        for p_ in m.parameters():
            p_.mul_(1.01)
    fp.refresh()
    with torch.no_grad():
        dist2, ref_v2, _ = m.evaluate(o, c)
    mu2, v2, _ = fp.forward(o, c)
    assert (mu2 - dist2.mean).abs().max().item() < 1e-2 and (v2 - ref_v2).abs().max().item() < 1e-2 * max(1.0, float(ref_v2.abs().max()))


@pytest.mark.gpu
def test_sample_kernel_equals_the_torch_expression():
    """uaisa_policy_sample (one launch) against the element-wise torch expression of rnn_ac.step it replaces (policy_rnn_ac.py:57-69): the same
    draws give the same actions bit for bit and log-probabilities within one fp32 rounding of the 3-term sum; values / log-probabilities land in
    the caller's buffers."""
    from uaisa_ppo.fused import FusedPolicy

    m = _model("cuda")
    fp = FusedPolicy(m, "cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for A, sf, ls in ((1, 1.0, (-1.0, -1.0, -1.0)), (1000, 0.7, (-1.0, 0.3, -2.5)), (70001, 1.0, (-30.0, 5.0, 0.0))):     # the last one hits both clamps
        with torch.no_grad():
            m.pi.log_std.copy_(torch.tensor(ls, device="cuda"))
        mu = torch.tanh(torch.randn((A, 3), device="cuda", generator=g))
        eps = torch.randn((A, 3), device="cuda", generator=g)
        out_lp = torch.full((A,), 7.0, device="cuda")
        a, logp = fp.sample(mu, sf, out_logp=out_lp, eps=eps)
        std = torch.clamp(sf * torch.exp(m.pi.log_std) + 1e-6, min=1e-4, max=10.0)
        a_ref = mu + std * eps
        lp_ref = (-((a_ref - mu) ** 2) / (2 * std * std) - torch.log(std) - 0.9189385332046727).sum(-1)
        assert logp.data_ptr() == out_lp.data_ptr()
        assert torch.equal(a, a_ref), (A, (a - a_ref).abs().max())
        assert (logp - lp_ref).abs().max().item() <= 2e-6 * max(1.0, float(lp_ref.abs().max())), (A, (logp - lp_ref).abs().max())
    # step(): values and log-probabilities written straight into rollout-buffer rows
    D = 12 + 9 * 3
    obs = torch.zeros((300, D), device="cuda"); obs[:, :12] = torch.randn((300, 12), device="cuda", generator=g)
    cnt = torch.zeros(300, dtype=torch.uint8, device="cuda")
    buf_v, buf_lp = torch.zeros((2, 300), device="cuda"), torch.zeros((2, 300), device="cuda")
    a, v, logp = fp.step(obs, cnt, 1.0, out_v=buf_v[1], out_logp=buf_lp[1])
    mu, v_ref, _ = fp.forward(obs, cnt)
    assert v.data_ptr() == buf_v[1].data_ptr() and torch.equal(buf_v[1], v_ref) and float(buf_v[0].abs().max()) == 0.0
    assert torch.isfinite(buf_lp[1]).all() and float(buf_lp[1].abs().max()) > 0 and float(buf_lp[0].abs().max()) == 0.0


@pytest.mark.gpu
def test_tiles_without_neighbour_rows_take_the_constant_reader_output():
    """A 128-agent tile in which NO agent has a neighbour row skips the GRU (no weight tile, no MMA, no gate): the reader's output is the constant h0 packed with
    the weights.  Same values BIT FOR BIT as the general path (an agent without neighbours feeds one zero row, ir_gym.py:211-212, so its accumulators are
    exactly 0 there), whatever the observation holds in the unused row slots; and equal to the torch module within the TF32 bound."""
    from uaisa_ppo.fused import FusedPolicy

    torch.backends.cuda.matmul.allow_tf32 = False
    m = _model("cuda")
    fp = FusedPolicy(m, "cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    A, D = 300, 12 + 9 * 3                                            # tiles 0, 1 full, tile 2 ragged
    obs = torch.zeros((A, D), device="cuda"); obs[:, :12] = torch.randn((A, 12), device="cuda", generator=g) * 4
    cnt0 = torch.zeros(A, dtype=torch.uint8, device="cuda")
    fp.stats.zero_()
    mu0, v0, f0 = fp.forward(obs, cnt0, want_features=True)           # every tile takes the constant path
    assert fp.stats.tolist() == [0, 0]
    with torch.no_grad():
        ref_f = m.features(obs, cnt0); dist, ref_v, _ = m.evaluate(obs, cnt0)
    assert (f0 - ref_f).abs().max().item() < 1e-3 and (mu0 - dist.mean).abs().max().item() < 1e-2 and (v0 - ref_v).abs().max().item() < 1e-2 * max(1.0, float(ref_v.abs().max()))
    # one agent with a row in every tile: the tiles run the GRU step now; the agents without rows must not change by a bit
    obs1, cnt1 = obs.clone(), cnt0.clone()
    for a in (127, 128, 299):
        cnt1[a] = 1; obs1[a, 12:21] = torch.randn(9, device="cuda", generator=g)
    mu1, v1, f1 = fp.forward(obs1, cnt1, want_features=True)
    keep = cnt1 == 0
    assert torch.equal(mu0[keep], mu1[keep]) and torch.equal(v0[keep], v1[keep]) and torch.equal(f0[keep], f1[keep])
    assert not torch.equal(f0[127], f1[127])
    # stale data in the row slots of agents without neighbours is never read -- in either path
    obs2 = obs.clone(); obs2[:, 12:] = torch.randn((A, D - 12), device="cuda", generator=g) * 3
    mu2, v2, f2 = fp.forward(obs2, cnt0, want_features=True)
    assert torch.equal(mu0, mu2) and torch.equal(v0, v2) and torch.equal(f0, f2)
    obs3 = obs1.clone(); obs3[keep, 12:] = obs2[keep, 12:]
    mu3, v3, f3 = fp.forward(obs3, cnt1, want_features=True)
    assert torch.equal(mu1, mu3) and torch.equal(v1, v3) and torch.equal(f1, f3)
    # ... nor in a tile that runs the recurrence (an agent with 3 rows next to agents without any): both directions feed the zero row
    obs4, cnt4 = obs1.clone(), cnt1.clone()
    cnt4[5] = 3; obs4[5, 12:] = torch.randn(D - 12, device="cuda", generator=g)
    fp.stats.zero_()
    mu4, v4, f4 = fp.forward(obs4, cnt4, want_features=True)
    assert fp.stats.tolist() == [1, 3]
    keep4 = cnt4 == 0
    obs5 = obs4.clone(); obs5[keep4, 12:] = obs2[keep4, 12:]
    mu5, v5, f5 = fp.forward(obs5, cnt4, want_features=True)
    assert torch.equal(mu4, mu5) and torch.equal(v4, v5) and torch.equal(f4, f5)
    with torch.no_grad():
        ref4 = m.features(obs4, cnt4)
    assert (f4 - ref4).abs().max().item() < 1.5e-2 and (f4[keep4] - ref4[keep4]).abs().max().item() < 4e-3

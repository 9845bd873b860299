"""-m gpu: the reference's quirk states (SURVEY.md section 8a "quirks"), constructed on purpose and injected through uaisa_set_state:
exactly collinear pairs (unclamped arccos -> NaN beta -> outside the cone, vel_obs3D.py:54-61), co-located distinct drones (excluded like
"self", rvo_inter.py:92), roots t = -1 (vel_obs3D.py:174-175 -> 1/(t+0.2) = -1.25), des_vel = 0 at the destination (inf / nan reward,
ir_gym.py:88) -- plus a >= 1 M agent-step soak and the ep_len-after-reset rule.  The oracle these are compared with is itself pinned on
the same kinds of pairs against the unmodified reference (tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest
import torch

import _util  # noqa: F401
import oracle
from gpu_harness import Stats, lockstep
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_world3
from uaisa_env.vec_env import VecMdin, pack_state, unpack_state

pytestmark = pytest.mark.gpu


def _one_step_compare(scn, E, nm, st, act):
    """Load `st` on both sides, step once with `act`, compare everything.  -> (Stats, oracle outputs)"""
    env = VecMdin(scn, E, neighbors_num=nm); ob = oracle.OracleBatch(scn, E, nm=nm)
    env.reset(); ob.reset()
    env.set_state(pack_state(st))
    st32 = unpack_state(env.get_state())
    ob.set_state(st32)
    env.step(torch.from_numpy(act.astype(np.float32)).cuda())
    oobs, orew, odone, oinfo, ofin = ob.step(act.astype(np.float32).astype(np.float64))
    s = Stats()
    g = {k: getattr(env, k).cpu().numpy() for k in ("obs", "count", "nbr", "reward", "done", "info", "finish")}
    s.flag_mismatch += int((g["done"] != odone).sum() + (g["info"] != oinfo).sum() + (g["finish"] != ofin).sum())
    s.count_mismatch += int((g["count"] != ob.count.astype(np.uint8)).sum()); s.nbr_mismatch += int((g["nbr"] != ob.nbr).sum())
    s.add_obs(g["obs"], oobs); s.add_rew(g["reward"], orew)
    gst, ost = unpack_state(env.get_state()), ob.get_state()
    for k in ("pos", "vel", "yaw", "pitch", "real_len", "max_dev"):
        s.add_state(gst[k].astype(np.float32), ost[k])
    env.close()
    return s, dict(obs=oobs.copy(), count=ob.count.copy(), reward=orew.copy(), done=odone.copy())


def _pairs_scenario(E):
    """E independent 2-drone envs on one big map (nothing leaves it), far-away goals."""
    return ScenarioSet.from_start_end([[20, 20, 20], [22, 20, 20]], [[180, 180, 180], [170, 185, 175]], (200, 200, 200))


def _base_state(E, N):
    A = E * N
    return {"pos": np.zeros((A, 3)), "vel": np.zeros((A, 3)), "yaw": np.zeros(A), "pitch": np.zeros(A), "wp_idx": np.ones(A, np.int32),
            "arrive_flag": np.zeros(A, np.uint8), "dest_flag": np.zeros(A, np.uint8), "real_len": np.zeros(A), "max_dev": np.zeros(A), "extra_len": np.zeros(A)}


def test_exactly_collinear_pairs_nan_beta():
    """rel, the velocities and the action all parallel on a binary lattice: dot / (|A| |B|) lands on 1 +- 2^-52 and the reference's
    unclamped arccos returns NaN for a large share of such pairs (SURVEY: 43 234 of 200 000) -> the pair is OUTSIDE the cone."""
    E = 4096
    rng = np.random.default_rng(0)
    scn = _pairs_scenario(E)
    st = _base_state(E, 2)
    base = rng.integers(-3, 4, (E, 3)).astype(np.float64)
    base[np.all(base == 0, axis=1)] = [1, 2, 2]
    rel = base * rng.choice([0.25, 0.5, 0.75, 1.0], (E, 1))
    pa = 50.0 + rng.integers(0, 40, (E, 3)) * 0.25
    st["pos"][0::2] = pa; st["pos"][1::2] = pa + rel
    st["vel"][0::2] = rel * rng.choice([0.125, 0.25, 0.5], (E, 1)); st["vel"][1::2] = rel * rng.choice([-0.25, 0.125, 0.25], (E, 1))
    act = np.zeros((2 * E, 3))
    act[0::2] = rel * rng.choice([0.25, 0.5, 1.0], (E, 1)); act[1::2] = -rel * rng.choice([0.25, 0.5], (E, 1))
    # what the device will see: fp32 images (all lattice values here are exact in fp32)
    s, o = _one_step_compare(scn, E, 10, st, act)
    # Lattice inputs also make exact DECIMAL TIES out of exact trigonometric identities (e.g. speed 0.75*sqrt(2) at yaw 292.5, pitch -67.5 deg gives
    # v_y = -0.375 mathematically): which side of the tie a 2-dp observation lands on depends on the last ulp of the libm sin / cos, so a couple of
    # observation entries may move by one quantum (0.01).  Everything discrete -- flags, counts, neighbour indices, i.e. the cone decisions this test is
    # about -- and the state must still agree.
    assert s.flag_mismatch == 0 and s.count_mismatch == 0 and s.nbr_mismatch == 0, s.summary()
    assert s.state_maxrel <= 1e-5 and s.rew_maxrel <= 1e-5 and s.obs_inexact <= 6, s.summary()
    # the category really occurs: pre-move pairs evaluated through the oracle's pair function with NaN beta are reported outside (vo False)
    n_nan = 0
    for e in range(0, E, 8):
        sa = np.concatenate([st["pos"][2 * e], st["vel"][2 * e], [0.2, 5.0]]); sb = np.concatenate([st["pos"][2 * e + 1], st["vel"][2 * e + 1], [0.2, 5.0]])
        a = act[2 * e]
        relv = sb[:3] - sa[:3]; B = sa[:3] + 2 * a - 0.5 * (2 * sa[:3] + sa[3:6] + sb[3:6])
        nA, nB = np.linalg.norm(relv), np.linalg.norm(B)
        if nB > 0 and np.dot(sa[3:6], relv) > 0 and nA > 0.4:
            with np.errstate(invalid="ignore"):
                n_nan += int(np.isnan(np.arccos(np.dot(relv, B) / (nA * nB))))
    assert n_nan > 10, n_nan


def test_colocated_distinct_drones_are_excluded_like_self():
    E = 256
    rng = np.random.default_rng(1)
    scn = ScenarioSet.from_start_end([[20, 20, 20], [22, 20, 20], [24, 20, 20]], [[180, 180, 180], [170, 185, 175], [160, 170, 190]], (200, 200, 200))
    st = _base_state(E, 3)
    p = 60.0 + np.round(rng.uniform(0, 30, (E, 3)), 2)
    st["pos"][0::3] = p; st["pos"][1::3] = p                       # drones 0 and 1 at EXACTLY the same position
    st["pos"][2::3] = p + np.round(rng.uniform(0.5, 3.0, (E, 3)), 2)
    v = np.round(rng.uniform(-1, 1, (3 * E, 3)), 2)
    st["vel"][:] = v
    act = np.round(0.5 * np.round(rng.uniform(-1, 1, (3 * E, 3)), 2) + v, 2)
    s, o = _one_step_compare(scn, E, 10, st, act)
    s.assert_parity()
    # co-located drones never collide with each other in the reference (the pair is skipped), and they stay co-located only if they move alike
    env = VecMdin(scn, E, neighbors_num=10); env.reset(); env.set_state(pack_state(st))
    env.observe()
    nbr = env.nbr.cpu().numpy().reshape(E, 3, -1)
    assert not np.any(nbr[:, 0] == 1) and not np.any(nbr[:, 1] == 0)     # they never list each other (reset observation: action 0)
    env.close()


def test_negative_roots_and_destination_rewards():
    """Pairs found by random search through the oracle's pair function whose expected collision time is -1 (both roots negative) are fed to
    the kernel as 2-drone envs; half of the envs additionally park drone 1 on its destination (des_vel = 0 -> inf / nan reward)."""
    rng = np.random.default_rng(2)
    found = []
    cfg = oracle.default_config(10)
    tries = 0
    # both roots negative <=> the relative velocity 2a - va - vb points AWAY from the other drone, while "inside the cone" looks at
    # 2a - (va + vb)/2: take va + vb = sigma * u along the line of sight u and 2a = (va + vb)/2 + lambda * u with 0 < lambda < sigma/2
    while len(found) < 300 and tries < 20000:
        tries += 1
        pa = np.round(rng.uniform(60, 70, 3), 2)
        u = rng.normal(0, 1, 3); u /= np.linalg.norm(u)
        rel = np.round(u * rng.uniform(0.5, 3.0), 2)
        sg = rng.uniform(1.0, 3.0); lam = rng.uniform(0.05, sg / 2 - 0.05)
        va = np.round(0.5 * sg * u + rng.normal(0, 0.03, 3), 2); vb = np.round(0.5 * sg * u + rng.normal(0, 0.03, 3), 2)
        a = np.round(0.25 * sg * u + 0.5 * lam * u + rng.normal(0, 0.01, 3), 2)
        row, vo, t, col, md = oracle.pair(np.concatenate([pa, va, [np.float32(0.2), 5.0]]), np.concatenate([pa + rel, vb, [np.float32(0.2), 5.0]]), a, cfg)
        if vo and t == -1.0:
            found.append((pa, rel, va, vb, a))
    assert len(found) >= 100, (len(found), tries)
    E = len(found)
    ends = [[180, 180, 180], [170, 185, 175]]
    scn = ScenarioSet.from_start_end([[20, 20, 20], [22, 20, 20]], ends, (200, 200, 200))
    st = _base_state(E, 2)
    act = np.zeros((2 * E, 3))
    for e, (pa, rel, va, vb, a) in enumerate(found):
        st["pos"][2 * e] = pa; st["pos"][2 * e + 1] = pa + rel; st["vel"][2 * e] = va; st["vel"][2 * e + 1] = vb
        # the env re-reads the action as controls (acc, yaw/90, pitch/90) for the MOVE, but PASS A evaluates the cone on the pre-move state with the raw action
        act[2 * e] = a; act[2 * e + 1] = np.round(rng.uniform(-1, 1, 3), 2)
    s, o = _one_step_compare(scn, E, 10, st, act)
    s.assert_parity()
    # second batch: drone 1 sits on its destination with both flags latched: des_vel = 0, reward inf or nan exactly like the reference
    st2 = {k: v.copy() for k, v in st.items()}
    st2["pos"][1::2] = np.asarray(ends[1], np.float64); st2["vel"][1::2] = 0.0
    st2["arrive_flag"][1::2] = 1; st2["dest_flag"][1::2] = 1
    act2 = act.copy(); act2[1::2][: E // 2] = 0.0
    s2, o2 = _one_step_compare(scn, E, 10, st2, act2)
    s2.assert_parity()
    r1 = o2["reward"][1::2]
    assert np.isinf(r1).sum() > 0 and np.isnan(r1).sum() > 0, (np.isinf(r1).sum(), np.isnan(r1).sum())


def test_soak_one_million_agent_steps():
    """>= 1 M teacher-forced agent-steps over warp-local, CTA and many-neighbour geometries inside the -m gpu suite (the 10 M run of
    tests/soak_parity.py stays a script)."""
    import test_gpu_parity as tg
    total = Stats()
    for mk, E, nm, T in ((scenario_c1, 3072, 10, 45), (scenario_world3, 2048, 10, 40), (lambda: scenario_c3(seed=5, n_scenarios=4), 384, 10, 40),
                         (lambda: tg.swarm(12, 1.0, 2), 512, 4, 40)):
        lockstep(mk(), E, nm, T, seed=23, epoch_len=31, block=10, stats=total)
    total.assert_parity()
    assert total.n_agent_steps >= 1_000_000 and total.rows > 10000 and total.dones > 1000 and total.resets > 1000, total.summary()
    assert total.obs_inexact <= 1e-5 * total.obs_vals and total.rew_inexact <= 1e-4 * total.rew_vals, total.summary()


def test_reset_starts_a_new_episode_for_the_reset_policy():
    """uaisa_reset zeroes the episode length the trainer reset policy counts (multi_ppo.py:168, 238, 261): a reset() after some auto-reset steps
    must not inherit stale lengths (the max_ep_len cut would fire early)."""
    scn = ScenarioSet.from_start_end([[20, 20, 20], [40, 20, 20], [60, 20, 20]], [[180, 180, 180], [170, 185, 175], [160, 170, 190]], (200, 200, 200))   # nobody collides, leaves or arrives
    E, T, M = 64, 20, 12
    rng = np.random.default_rng(4)
    acts = torch.from_numpy(np.round(rng.uniform(-0.2, 0.2, (T, E * scn.N, 3)), 2).astype(np.float32)).cuda()
    a = VecMdin(scn, E); b = VecMdin(scn, E)
    a.reset(); b.reset()
    a.rollout(acts[:7], {}, auto_reset=True, max_ep_len=M)          # ep_len = 7 everywhere (slow drones: nobody is done)
    a.reset()
    outs = []
    for env in (a, b):
        cnt = torch.zeros((T, env.A), dtype=torch.uint8, device="cuda"); rew = torch.zeros((T, env.A), device="cuda")
        obs = torch.zeros((T, env.A, env.obs_dim), device="cuda")
        env.rollout(acts, {"count": cnt, "reward": rew, "obs": obs}, auto_reset=True, max_ep_len=M)
        outs.append((cnt, torch.nan_to_num(rew, 7.0), obs, env.get_state()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert all(torch.equal(outs[0][3][k], outs[1][3][k]) for k in outs[0][3])
    # masked reset: only the reset agents start a new episode
    m = torch.zeros(a.A, dtype=torch.uint8, device="cuda"); m[::3] = 1
    a.reset(); a.rollout(acts[:5], {}, auto_reset=True, max_ep_len=M); a.reset(m)
    b.reset(); b.rollout(acts[:5], {}, auto_reset=True, max_ep_len=M); b.reset(m)
    pc = [torch.zeros(x.A, dtype=torch.uint8, device="cuda") for x in (a, b)]
    for t in range(9):
        for env in (a, b):
            env.step(acts[t]); env.autoreset(max_ep_len=M)
        assert torch.equal(a.path_cut, b.path_cut) and torch.equal(a.reset_mask, b.reset_mask)
        if t == 7:      # drones 1 and 2 of every env have now run 5 + 8 = 13 > 12 steps, drone 0 only 8
            assert bool(a.reset_mask.view(E, 3)[:, 1:].all()) and not bool(a.reset_mask.view(E, 3)[:, 0].any())
    a.close(); b.close()
    del pc


@pytest.mark.parametrize("name,E,steps", [("c2", 1024, 12), ("c3", 4096, 5), ("c4", 64, 4), ("c5", 65536, 5)])
def test_full_size_lockstep_on_diverse_data(name, E, steps):
    """BASELINE.json's full sizes with DIVERSE data: every env has its own action stream (and, for configs[2] / [3], one of several scenarios); the
    whole batch is compared with the oracle every step (teacher-forced), not one representative env."""
    from uaisa_env.scenario import scenario_c4
    scn = {"c2": scenario_c1, "c5": scenario_c1, "c3": lambda: scenario_c3(seed=0, n_scenarios=8), "c4": lambda: scenario_c4(seed=0, n_scenarios=2)}[name]()
    st = lockstep(scn, E, 10, steps, seed=31, epoch_len=0, block=2, max_ep_len=3)
    st.assert_parity()
    assert st.n_agent_steps == E * scn.N * steps and st.resets > 0
    assert st.obs_inexact <= 1e-5 * st.obs_vals and st.rew_inexact <= 1e-4 * st.rew_vals, st.summary()

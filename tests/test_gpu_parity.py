"""-m gpu: the CUDA path (through the C ABI) against the oracle on identical seeded inputs.
Contract (BASELINE.json north_star): done/info/finish flags, counts and neighbour indices bit-exact;
positions, velocities, rewards and observations within 1e-5 relative (fp32 state vs fp64 reference)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import _util  # noqa: F401
import oracle
from gpu_harness import Stats, lockstep, f32
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_c4, scenario_world3
from uaisa_env.vec_env import VecMdin, pack_state, unpack_state

pytestmark = pytest.mark.gpu
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("policy_"))


def swarm(n=12, rad=1.0, seed=0):
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(n) / n
    st = np.stack([5 + rad * np.cos(ang), 5 + rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    en = np.stack([5 - rad * np.cos(ang), 5 - rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    return ScenarioSet.from_start_end(np.round(st, 2), np.round(en, 2), (10, 10, 5), buildings=[[5, 5, 1.5, 0.3]])


def golden_scn(name):
    z = np.load([p for p in GOLDEN if os.path.basename(p) == name + ".npz"][0])
    return ScenarioSet.from_dict(json.loads(str(z["scenario_json"])))


@pytest.mark.parametrize("name,E,nm,steps", [
    ("c1", 64, 10, 150), ("world3", 85, 10, 150), ("world_8", 32, 10, 150), ("world_8", 16, 2, 100),
    ("swarm12", 24, 3, 120), ("swarm12", 24, 10, 120), ("c3", 16, 10, 120), ("swarm20", 8, 16, 80)])
def test_lockstep_teacher_forced(name, E, nm, steps):
    scn = {"c1": scenario_c1, "world3": scenario_world3, "world_8": lambda: golden_scn("world_8"),
           "swarm12": lambda: swarm(12, 1.0, 3), "swarm20": lambda: swarm(20, 1.3, 4),
           "c3": lambda: scenario_c3(seed=2, n_scenarios=4)}[name]()
    st = lockstep(scn, E, nm, steps, seed=hash((name, nm)) % 1000, epoch_len=70, block=20)
    st.assert_parity()
    assert st.dones > 0 and st.resets > 0
    if name.startswith("swarm"):
        assert st.rows > 200
    # in practice the fp32 images are bit-identical; keep a hard ceiling on the inexact fraction
    assert st.obs_inexact <= 1e-4 * st.obs_vals and st.rew_inexact <= 1e-3 * st.rew_vals, st.summary()


def test_dense_256():
    """configs[3] shape: one CTA per 256-drone env, neighbour lists truncated to nm."""
    scn = scenario_c4(seed=1, n_scenarios=2, n_drones=256)
    st = lockstep(scn, 3, 10, 12, seed=5, block=3)
    st.assert_parity()
    assert st.rows > 0


@pytest.mark.parametrize("N,tpb,l2", [(300, 0, False), (200, 64, False), (130, 32, False), (256, 128, False), (200, 64, True), (256, 0, True)])
def test_cluster_mode_envs_larger_than_one_cta(N, tpb, l2, monkeypatch):
    """N > 128: one env is spread over a thread-block cluster; the columns are exchanged through distributed shared memory
    (st.shared::cluster + cluster barrier) or, with UAISA_CLUSTER_L2, through a double-buffered array in L2."""
    if l2:
        monkeypatch.setenv("UAISA_CLUSTER_L2", "1")
    scn = scenario_c4(seed=3, n_scenarios=1, n_drones=N)
    st = lockstep(scn, 3, 10, 6, seed=9, block=2, threads_per_block=tpb, max_ep_len=4)
    st.assert_parity()
    assert st.rows > 0 and st.resets > 0


@pytest.mark.parametrize("N,tpb,slices", [(256, 32, 3), (256, 32, 4), (256, 64, 2), (300, 0, 0), (64, 64, 2), (100, 0, 0), (40, 64, 2)])
def test_neighbour_slices(N, tpb, slices):
    """Helper threads share each ego's neighbour loop (dense airspace, few envs): thread-block clusters (N > 128, also
    with a ragged last CTA) and single-CTA envs, explicit and auto-selected geometries; parity incl. the top-nm lists merged
    from the slices."""
    scn = scenario_c4(seed=4, n_scenarios=1, n_drones=N) if N > 128 else swarm(N, N / 12.0, 5)
    if N <= 128 and tpb and tpb < N:
        tpb = ((N + 31) // 32) * 32
        slices = min(slices, 128 // tpb)
    st = lockstep(scn, 3, 10, 6, seed=11, block=2, threads_per_block=tpb, neighbor_slices=slices, max_ep_len=4)
    st.assert_parity()
    assert st.rows > 0 and st.resets > 0


def test_edge_shapes():
    # single drone, no neighbours; nm = 0; N=3 (CTA not a multiple of N); explicit CTA sizes
    one = ScenarioSet.from_start_end([[1, 1, 1]], [[3, 3, 2]], (5, 5, 3))
    lockstep(one, 5, 10, 40, seed=1).assert_parity()
    lockstep(scenario_world3(), 7, 0, 40, seed=2).assert_parity()
    lockstep(scenario_world3(), 50, 10, 40, seed=3, threads_per_block=32).assert_parity()
    lockstep(swarm(12, 1.0, 7), 30, 10, 40, seed=4, threads_per_block=64).assert_parity()
    lockstep(scenario_c1(), 9, 10, 40, seed=5, env_train=0).assert_parity()
    lockstep(scenario_c1(), 37, 10, 60, seed=6, threads_per_block=64, neighbor_slices=2).assert_parity()      # owner / helper split, two warp pairs per CTA
    lockstep(scenario_c1(), 37, 10, 60, seed=6, neighbor_slices=1).assert_parity()                            # ... and the same batch unsplit


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_reference_golden(path):
    """CUDA path vs outputs RECORDED FROM THE REFERENCE itself: the device is loaded with the fp32
    image of each recorded pre-state, so differences are bounded by that input rounding (a 2-dp
    rounded observation may move by one quantum when the rounding of the state crosses a tie)."""
    z = np.load(path)
    scn = ScenarioSet.from_dict(json.loads(str(z["scenario_json"])))
    nm = int(z["nm"])
    env = VecMdin(scn, 1, neighbors_num=nm)
    T = z["action"].shape[0]
    flag_bad = 0; obs_bad = 0; rew_bad = 0; nvals = 0
    for t in range(T):
        pre = {k: z["pre_" + k][t] for k in ("pos", "vel", "yaw", "pitch", "wp_idx", "arrive_flag", "dest_flag", "real_len", "max_dev", "extra_len")}
        env.set_state(pack_state(pre))
        env.step(torch.from_numpy(z["action"][t].astype(np.float32)).cuda())
        o = env.obs.cpu().numpy().astype(np.float64); r = env.reward.cpu().numpy().astype(np.float64)
        flag_bad += int((env.done.cpu().numpy() != z["done"][t]).sum() + (env.finish.cpu().numpy() != z["finish"][t]).sum()
                        + (env.info.cpu().numpy() != z["info"][t]).sum() + (env.count.cpu().numpy() != z["count"][t]).sum())
        same_cnt = env.count.cpu().numpy() == z["count"][t]
        d = np.abs(o - z["obs"][t])[same_cnt]
        assert d.size == 0 or d.max() <= 0.0100001 + 1e-5 * np.abs(z["obs"][t]).max(), (t, d.max())
        obs_bad += int((d > 1e-5 * np.maximum(1, np.abs(z["obs"][t][same_cnt]))).sum()); nvals += d.size
        ref_r = z["reward"][t]
        fin = np.isfinite(ref_r)
        rew_bad += int((np.abs(r[fin] - ref_r[fin]) > 2e-3 + 1e-5 * np.abs(ref_r[fin])).sum()) + int((np.isfinite(r) != fin).sum())
    env.close()
    n = T * scn.N
    assert flag_bad <= max(1, n // 500), flag_bad
    assert obs_bad <= max(2, nvals // 2000), (obs_bad, nvals)
    assert rew_bad <= max(1, n // 200), rew_bad


F32_GOLDEN = [p for p in GOLDEN if os.path.basename(p).startswith("f32_")]


@pytest.mark.parametrize("path", F32_GOLDEN, ids=[os.path.basename(p)[:-4] for p in F32_GOLDEN])
def test_against_reference_golden_exact(path):
    """CUDA path vs outputs recorded from the UNMODIFIED reference with fp32-representable pre-states and actions (oracle/gen_golden.py, f32 cases:
    the reference was teacher-forced with float64 copies of fp32 values).  No input-rounding slack is left, so the contract
    applies literally: every flag, count and ragged length identical; observations, rewards and post-states within 1e-5 relative of the
    reference's doubles (in practice: their fp32 images); also the post-reset re-observation."""
    z = np.load(path)
    scn = ScenarioSet.from_dict(json.loads(str(z["scenario_json"])))
    nm = int(z["nm"])
    env = VecMdin(scn, 1, neighbors_num=nm)
    T = z["action"].shape[0]
    st = Stats()
    n_inexact = 0
    for t in range(T):
        pre = {k: z["pre_" + k][t] for k in ("pos", "vel", "yaw", "pitch", "wp_idx", "arrive_flag", "dest_flag", "real_len", "max_dev", "extra_len")}
        for k in ("pos", "vel", "yaw", "pitch", "real_len", "max_dev", "extra_len"):
            assert np.array_equal(pre[k].astype(np.float32).astype(np.float64), pre[k]), (t, k)      # the fixture really is fp32-representable
        assert np.array_equal(z["action"][t].astype(np.float32).astype(np.float64), z["action"][t])
        env.set_state(pack_state(pre))
        env.step(torch.from_numpy(z["action"][t].astype(np.float32)).cuda())
        g = {k: getattr(env, k).cpu().numpy() for k in ("obs", "count", "reward", "done", "info", "finish")}
        st.flag_mismatch += int((g["done"] != z["done"][t]).sum() + (g["finish"] != z["finish"][t]).sum() + (g["info"] != z["info"][t]).sum())
        st.count_mismatch += int((g["count"] != z["count"][t]).sum())
        st.add_obs(g["obs"], z["obs"][t]); st.add_rew(g["reward"], z["reward"][t])
        post = unpack_state(env.get_state())
        for k in ("pos", "vel", "yaw", "pitch", "real_len", "max_dev"):
            st.add_state(post[k].astype(np.float32), z["post_" + k][t])
        st.flag_mismatch += int((post["wp_idx"] != z["post_wp_idx"][t]).sum() + (post["arrive_flag"] != z["post_arrive_flag"][t]).sum() + (post["dest_flag"] != z["post_dest_flag"][t]).sum())
        m = z["reset_mask"][t]
        if m.any():      # reset + env_observation from the device's own fp32 post-state, as recorded
            env.reset(torch.from_numpy(m.astype(np.uint8)).cuda())
            st.add_obs(env.obs.cpu().numpy(), z["obs_after_reset"][t])
            st.count_mismatch += int((env.count.cpu().numpy() != z["count_after_reset"][t]).sum())
    env.close()
    st.assert_parity(1e-5)
    n_inexact = st.obs_inexact + st.rew_inexact
    assert n_inexact <= 2, st.summary()       # (fp32 images of the reference's doubles, bit for bit, up to a libm ulp at a decimal tie)


@pytest.mark.parametrize("geom", ["swarm12", "wl4", "wl4_nosplit", "wl4_split64", "wl4_small", "c3_16", "swarm20_nm16", "cluster200", "sliced256", "sliced64"])
def test_rollout_equals_repeated_step_and_fused_autoreset(geom):
    """A T-step uaisa_rollout with the fused reset policy == T x (uaisa_step + uaisa_autoreset), bit for bit, in every geometry of the kernel
    (warp-local, CTA + re-observe compaction, thread-block cluster, neighbour slices).  The launch is cut into uneven chunks so that both the
    deferred re-observation (steps t < T-1 of a launch: evaluated inside PASS A of step t+1) and the in-step one (last step) are compared."""
    scn, E, nm, T, kw = {
        "swarm12": (lambda: swarm(12, 1.2, 9), 40, 10, 30, {}),
        "wl4": (scenario_c1, 75, 10, 40, {}),                                                        # small batch: owner / helper split chosen automatically
        "wl4_nosplit": (scenario_c1, 75, 10, 40, {"neighbor_slices": 1}),                            # the same through the unsplit warp-local kernel
        "wl4_split64": (scenario_c1, 203, 10, 40, {"neighbor_slices": 2, "threads_per_block": 64}),   # two owner / helper warp pairs per CTA, ragged last CTA
        "wl4_small": (lambda: swarm(4, 0.8, 2), 33, 3, 40, {}),
        "c3_16": (lambda: scenario_c3(seed=5, n_scenarios=3, n_drones=16, n_buildings=6), 24, 10, 30, {}),
        "swarm20_nm16": (lambda: swarm(20, 1.3, 4), 13, 16, 30, {}),
        "cluster200": (lambda: scenario_c4(seed=2, n_scenarios=1, n_drones=200), 3, 10, 14, {"threads_per_block": 64}),
        "sliced256": (lambda: scenario_c4(seed=3, n_scenarios=2, n_drones=256), 3, 10, 14, {"threads_per_block": 64, "neighbor_slices": 2}),
        "sliced64": (lambda: scenario_c4(seed=4, n_scenarios=1, n_drones=64), 5, 10, 20, {"threads_per_block": 64, "neighbor_slices": 2}),
    }[geom]
    scn = scn()
    rng = np.random.default_rng(0)
    A = E * scn.N
    acts = torch.from_numpy(np.round(rng.uniform(-1, 1, (T, A, 3)), 2).astype(np.float32)).cuda()
    e2s = (np.arange(E) % scn.S).astype(np.int32) if getattr(scn, "S", 1) > 1 else None
    a = VecMdin(scn, E, env2scenario=e2s, neighbors_num=nm, **kw); b = VecMdin(scn, E, env2scenario=e2s, neighbors_num=nm, **kw)
    a.reset(); b.reset()
    out = {"obs": torch.zeros((T, A, a.obs_dim), device="cuda"), "count": torch.zeros((T, A), dtype=torch.uint8, device="cuda"),
           "nbr": torch.zeros((T, A, nm), dtype=torch.int32, device="cuda"), "reward": torch.zeros((T, A), device="cuda"),
           "done": torch.zeros((T, A), dtype=torch.uint8, device="cuda"), "info": torch.zeros((T, A), dtype=torch.uint8, device="cuda"),
           "finish": torch.zeros((T, A), dtype=torch.uint8, device="cuda")}
    s = 0
    for chunk in (7, 1, 2, T):                       # launches of 7, 1, 2 and the remaining steps
        n = min(chunk, T - s)
        if n <= 0:
            break
        a.rollout(acts[s:s + n], {k: v[s:s + n] for k, v in out.items()}, auto_reset=True, max_ep_len=12)
        s += n
    resets = rows = 0
    for t in range(T):
        b.step(acts[t])
        for k in ("reward", "done", "info", "finish"):
            assert torch.equal(getattr(b, k).view(-1), out[k][t].view(-1)) or (k == "reward" and torch.equal(torch.nan_to_num(b.reward, 7.0), torch.nan_to_num(out[k][t], 7.0))), (t, k)
        b.autoreset(max_ep_len=12)
        resets += int(b.reset_mask.sum()); rows += int(b.count.sum())
        for k in ("obs", "count", "nbr"):
            assert torch.equal(getattr(b, k), out[k][t]), (geom, t, k)
    assert resets > 0, "the reset policy never fired: the test compares nothing"
    if geom in ("swarm12", "swarm20_nm16", "c3_16", "cluster200", "sliced256"):
        assert rows > 0, "no neighbour row in any re-observed env"
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    a.close(); b.close()


@pytest.mark.parametrize("geom", ["wl4", "c3_16", "swarm12"])
def test_step_autoreset_is_step_then_autoreset(geom):
    """uaisa_step_autoreset (one launch: what the PPO rollout issues per step) == uaisa_step + uaisa_autoreset, every output incl. reset_mask / path_cut."""
    scn = {"wl4": scenario_c1, "c3_16": lambda: scenario_c3(seed=5, n_scenarios=3, n_drones=16, n_buildings=6), "swarm12": lambda: swarm(12, 1.2, 9)}[geom]()
    E, T = 60, 45
    a = VecMdin(scn, E, neighbors_num=10); b = VecMdin(scn, E, neighbors_num=10)
    a.reset(); b.reset()
    rng = np.random.default_rng(3)
    n_reset = n_cut = 0
    for t in range(T):
        act = torch.from_numpy(np.round(rng.uniform(-1, 1, (a.A, 3)), 2).astype(np.float32)).cuda()
        ended = (t % 20 == 19)
        a.step_autoreset(act, max_ep_len=9, epoch_ended=ended)
        b.step(act)
        rew, done, info, fin = (x.clone() for x in (b.reward, b.done, b.info, b.finish))
        b.autoreset(max_ep_len=9, epoch_ended=ended)
        for k, ref in (("reward", rew), ("done", done), ("info", info), ("finish", fin), ("obs", b.obs), ("count", b.count), ("nbr", b.nbr),
                       ("reset_mask", b.reset_mask), ("path_cut", b.path_cut)):
            assert torch.equal(torch.nan_to_num(getattr(a, k).float(), 7.0), torch.nan_to_num(ref.float(), 7.0)), (geom, t, k)
        n_reset += int(b.reset_mask.sum()); n_cut += int(b.path_cut.sum())
    assert n_reset > 0 and n_cut > 0
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    a.close(); b.close()


def test_env_sharding_is_bitwise_deterministic():
    """a shard of envs computes exactly what the same env ids compute in the full batch."""
    scn = scenario_c3(seed=3, n_scenarios=3, n_drones=8, n_buildings=4)
    E, nm, T = 48, 10, 25
    e2s = (np.arange(E) % 3).astype(np.int32)
    rng = np.random.default_rng(1)
    acts = np.round(rng.uniform(-1, 1, (T, E, scn.N, 3)), 2).astype(np.float32)
    full = VecMdin(scn, E, env2scenario=e2s, neighbors_num=nm); full.reset()
    lo = VecMdin(scn, 16, env2scenario=e2s[:16], neighbors_num=nm); lo.reset()
    hi = VecMdin(scn, 32, env2scenario=e2s[16:], neighbors_num=nm, threads_per_block=64); hi.reset()
    for t in range(T):
        full.step(torch.from_numpy(acts[t].reshape(-1, 3)).cuda()); full.autoreset(max_ep_len=10)
        lo.step(torch.from_numpy(acts[t, :16].reshape(-1, 3).copy()).cuda()); lo.autoreset(max_ep_len=10)
        hi.step(torch.from_numpy(acts[t, 16:].reshape(-1, 3).copy()).cuda()); hi.autoreset(max_ep_len=10)
        for k in ("obs", "reward", "done", "finish", "count", "nbr"):
            cat = torch.cat([getattr(lo, k), getattr(hi, k)], 0)
            f = getattr(full, k)
            assert torch.equal(torch.nan_to_num(cat.float(), 7.0), torch.nan_to_num(f.float(), 7.0)), (t, k)
    for e in (full, lo, hi):
        e.close()


def test_masks_and_api_errors():
    scn = scenario_world3()
    env = VecMdin(scn, 10, neighbors_num=10, count_pairs=True)
    env.reset()
    act = torch.full((env.A, 3), 0.3, device="cuda")
    env.step(act)
    before = env.obs.clone()
    # observe only env 3: other rows untouched
    m = torch.zeros(env.E, dtype=torch.uint8, device="cuda"); m[3] = 1
    env.obs.fill_(-7.0)
    env.observe(m)
    assert (env.obs[9:12] != -7.0).all() and (env.obs[:9] == -7.0).all() and (env.obs[12:] == -7.0).all()
    # reset one drone of env 5: state back at its start, env 5 re-observed
    am = torch.zeros(env.A, dtype=torch.uint8, device="cuda"); am[16] = 1
    env.obs.fill_(-7.0)
    env.reset(am)
    st = unpack_state(env.get_state())
    assert np.array_equal(st["pos"][16], scn.waypoints[0, 1, 0]) and np.all(st["vel"][16] == 0) and st["wp_idx"][16] == 1
    assert (env.obs[15:18] != -7.0).all() and (env.obs[:15] == -7.0).all()
    c = env.counters()
    assert c["in_range_pairs"] > 0 and c["nonfinite_obs"] == 0
    assert env.launch_count >= 4
    with pytest.raises(ValueError):
        env.step(torch.zeros((env.A, 3), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        env.step(torch.zeros((env.A, 3)))
    env.close()
    bad = scenario_world3(); bad.n_points[0, 0] = 1
    with pytest.raises(Exception):
        VecMdin(bad, 2)
    del before


def test_step_host_matches_device_step():
    scn = scenario_c1()
    a = VecMdin(scn, 33, neighbors_num=10); b = VecMdin(scn, 33, neighbors_num=10)
    a.reset(); b.reset()
    rng = np.random.default_rng(3)
    out = {"obs": np.zeros((a.A, a.obs_dim), np.float32), "count": np.zeros(a.A, np.uint8), "nbr": np.zeros((a.A, 10), np.int32),
           "reward": np.zeros(a.A, np.float32), "done": np.zeros(a.A, np.uint8), "info": np.zeros(a.A, np.uint8), "finish": np.zeros(a.A, np.uint8)}
    for t in range(10):
        act = np.round(rng.uniform(-1, 1, (a.A, 3)), 2).astype(np.float32)
        a.step_host(act, out)
        b.step(torch.from_numpy(act).cuda())
        for k in out:
            assert np.array_equal(out[k], getattr(b, k).cpu().numpy(), equal_nan=(k == "reward")), (t, k)
    a.close(); b.close()


@pytest.mark.parametrize("N,E,copy_engine", [(12, 20, False), (4, 67, False), (16, 9, False), (256, 2, False), (12, 20, True)])
def test_step_host_sparse_matches_device_step(N, E, copy_engine, monkeypatch):
    """uaisa_step_host_sparse (one CUDA graph, ragged results) carries exactly the information of the dense step: the
    single-kernel path that writes mapped host memory (warp-local, CTA and cluster geometries) and the copy-engine path
    used for large batches."""
    if copy_engine:
        monkeypatch.setenv("UAISA_HOST_COPY_ENGINE", "1")
    scn = swarm(N, max(1.0, N / 12.0), 11) if N <= 128 else scenario_c4(seed=0, n_scenarios=1)      # N = 256: thread-block cluster geometry
    a = VecMdin(scn, E, neighbors_num=10); b = VecMdin(scn, E, neighbors_num=10)
    a.reset(); b.reset()
    l0 = a.launch_count
    rng = np.random.default_rng(4)
    rows_seen = 0
    for t in range(40):
        st = unpack_state(b.get_state())
        act = _util.training_like_actions(rng, st["vel"]).astype(np.float32) if t % 2 else np.round(rng.uniform(-1, 1, (a.A, 3)), 2).astype(np.float32)
        out = a.step_host_sparse(act, auto_reset=True, max_ep_len=15)
        b.rollout(torch.from_numpy(act).cuda().view(1, -1, 3), {"obs": b.obs.view(1, b.A, -1), "count": b.count.view(1, -1), "nbr": b.nbr.view(1, b.A, -1),
                                                                 "reward": b.reward.view(1, -1), "done": b.done.view(1, -1), "info": b.info.view(1, -1),
                                                                 "finish": b.finish.view(1, -1)}, auto_reset=True, max_ep_len=15)
        obs = b.obs.cpu().numpy(); cnt = b.count.cpu().numpy(); nbr = b.nbr.cpu().numpy()
        assert np.array_equal(out["reward"], b.reward.cpu().numpy(), equal_nan=True)
        for k in ("done", "info", "finish"):
            assert np.array_equal(out[k], getattr(b, k).cpu().numpy()), (t, k)
        assert np.array_equal(out["count"], cnt) and np.array_equal(out["self_obs"], obs[:, :12])
        assert len(out["rows"]) == int(cnt.sum())
        for r in out["rows"]:
            ag, sl = int(r["agent"]), int(r["slot"])
            assert sl < cnt[ag] and np.array_equal(r["v"], obs[ag, 12 + 9 * sl: 21 + 9 * sl]) and int(r["nbr"]) == nbr[ag, sl]
        rows_seen += len(out["rows"])
        if t % 8 == 7:      # the device state behind a polled host step is complete and visible to other streams right away
            sa, sb = a.get_state(), b.get_state()
            assert all(torch.equal(sa[k], sb[k]) for k in sa), t
        rag = VecMdin.ragged_from_sparse(out, 10)
        assert all(rag[k].shape[0] == 12 + 9 * max(int(cnt[k]), 1) and np.array_equal(rag[k], obs[k, : rag[k].shape[0]]) for k in range(a.A))
    assert rows_seen > (30 if N <= 128 else 4), rows_seen
    assert a.launch_count - l0 == (2 * 40 if copy_engine else 40)      # env kernel (+ pack kernel) per call
    a.close(); b.close()


@pytest.mark.parametrize("name,E", [("c2", 1024), ("c3", 4096), ("c4", 64), ("c5", 65536)])
def test_full_size_properties(name, E):
    """BASELINE.json's full sizes, through size-independent properties: (1) every env that shares a scenario and an
    action stream computes bitwise the same thing wherever it sits in the batch (CTA / warp / lane position);
    (2) that common result equals the oracle's for one such env; (3) a T-step rollout equals T single steps."""
    scn = {"c2": scenario_c1, "c5": scenario_c1, "c3": lambda: scenario_c3(seed=0, n_scenarios=1), "c4": lambda: scenario_c4(seed=0, n_scenarios=1)}[name]()
    N, T, nm = scn.N, 6, 10
    rng = np.random.default_rng(7)
    acts1 = np.round(rng.uniform(-1, 1, (T, 1, N, 3)), 2).astype(np.float32)
    acts = np.broadcast_to(acts1, (T, E, N, 3)).reshape(T, E * N, 3).copy()
    env = VecMdin(scn, E, neighbors_num=nm); env.reset()
    ob = oracle.OracleBatch(scn, 1, nm=nm); ob.reset()
    ep = np.zeros(N, np.int32)
    A = E * N
    for t in range(T):
        ob.set_state({k: v[:N] for k, v in unpack_state(env.get_state()).items()})
        env.step(torch.from_numpy(acts[t]).cuda())
        oo, orw, od, oi, of = ob.step(acts1[t, 0].astype(np.float64))
        for k in ("obs", "reward", "done", "info", "finish", "count", "nbr"):
            g = getattr(env, k).view(E, N, -1)
            assert torch.equal(torch.nan_to_num(g.float(), 7.0), torch.nan_to_num(g[:1].float(), 7.0).expand_as(g)), (t, k)
        assert np.array_equal(env.done.view(E, N)[E - 1].cpu().numpy(), od) and np.array_equal(env.count.view(E, N)[E // 2].cpu().numpy(), ob.count.astype(np.uint8))
        assert np.array_equal(env.obs.view(E, N, -1)[E - 1].cpu().numpy(), f32(oo))
        assert np.allclose(env.reward.view(E, N)[0].cpu().numpy(), f32(orw), rtol=1e-5, atol=1e-5, equal_nan=True)
        env.autoreset(max_ep_len=4)
        ob.set_state({k: v[:N] for k, v in unpack_state(env.get_state()).items()})    # oracle follows the device state (teacher forcing)
    # rollout == repeated steps at full size (checksums of every output)
    env2 = VecMdin(scn, E, neighbors_num=nm); env2.reset()
    env3 = VecMdin(scn, E, neighbors_num=nm); env3.reset()
    rew = torch.zeros((T, A), device="cuda"); cnt = torch.zeros((T, A), dtype=torch.uint8, device="cuda")
    env2.rollout(torch.from_numpy(acts).cuda(), {"reward": rew, "count": cnt}, auto_reset=True, max_ep_len=4)
    for t in range(T):
        env3.step(torch.from_numpy(acts[t]).cuda())
        assert torch.equal(torch.nan_to_num(env3.reward, 7.0), torch.nan_to_num(rew[t], 7.0)), t
        env3.autoreset(max_ep_len=4)
        assert torch.equal(env3.count, cnt[t]), t
    s2, s3 = env2.get_state(), env3.get_state()
    assert all(torch.equal(s2[k], s3[k]) for k in s2)
    for e in (env, env2, env3):
        e.close()

"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): the C
restatement in lock-step against the live, unmodified reference with fresh seeds."""
import json
import os

import numpy as np
import pytest

import _util  # noqa: F401
import oracle
import ref_shim
from uaisa_env.scenario import ScenarioSet

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")

from test_oracle_golden import close  # noqa: E402


def _lockstep(scn_dict, nm, steps, seed, block=25):
    import lockstep

    scn = ScenarioSet.from_dict(scn_dict)
    tr = lockstep.run_reference(scn_dict, nm, steps, seed, block=block)
    ob = oracle.OracleBatch(scn, 1, nm=nm)
    ob.reset()
    for t in range(steps):
        obs, rew, done, info, fin = ob.step(tr["action"][t])
        assert np.array_equal(done, tr["done"][t]) and np.array_equal(info, tr["info"][t]) and np.array_equal(fin, tr["finish"][t]), t
        assert np.array_equal(ob.count, tr["count"][t]), t
        assert close(obs, tr["obs"][t], 1e-8) and close(rew, tr["reward"][t], 1e-8), t
        m = tr["reset_mask"][t]
        if m.any():
            ob.reset(m)
            o2, c2, _ = ob.observe()
            assert close(o2, tr["obs_after_reset"][t], 1e-8) and np.array_equal(c2, tr["count_after_reset"][t]), t
    return tr


@pytest.mark.parametrize("world,seed", [("world_3", 11), ("world_8", 12), ("world_4", 13), ("world_2", 14)])
def test_shipped_worlds(world, seed):
    d = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "uaisa_env", "world", world, "data_1.json")))
    _lockstep(d, 10, 300, seed)


def test_dense_swarm_truncation():
    from gen_golden import swarm

    tr = _lockstep(swarm(10, 1.0, 5).to_dict(), 2, 200, 21, block=20)
    assert tr["count"].max() == 2 and tr["count"].sum() > 200


def test_pair_function_known_answers():
    """config_vo_circle2 directly (rvo_inter.py:116-196) on random pairs incl. collinear ones."""
    env = ref_shim.make_reference_env(json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "uaisa_env", "world", "world_3", "data_1.json"))))
    rvo = env.ir_gym.rvo
    rng = np.random.default_rng(0)
    n_full = 0
    for it in range(4000):
        pa = rng.uniform(0, 4, 3); rel = rng.normal(0, 1.0, 3) * rng.choice([0.3, 1.0, 3.0])
        va = rng.normal(0, 0.7, 3); vb = rng.normal(0, 0.7, 3)
        if it % 5 == 0:
            va = rel * rng.uniform(0.1, 1.0)      # exactly collinear: exercises the unclamped arccos
        act = np.round(va + rng.normal(0, 0.3, 3), 2) if it % 3 else np.zeros(3)
        sa = np.concatenate([pa, va, [0.2, 5.0]]); sb = np.concatenate([pa + rel, vb, [0.2, rng.choice([5.0, 3.0])]])
        ref = rvo.config_vo_circle2(list(sa) + [0, 0, 0, 0], list(sb) + [0, 0, 0, 0], act)
        row, vo, t, col, md = oracle.pair(sa, sb, act)
        assert vo == bool(ref[1]) and col == bool(ref[3]), it
        assert close(row, np.array(ref[0], dtype=np.float64), 1e-10), (it, row, ref[0])
        assert close(t, ref[2], 1e-10) and close(md, ref[4], 1e-10), it
        n_full += row[6] != 0
    assert n_full > 500


def test_reference_loads_a_scenario_directory_written_by_the_product(tmp_path):
    """ScenarioSet.save_scenario_dir / save_drone_paths_yaml produce what the reference's own loaders read (env_base.py:26-47,
    train/drone_paths.yaml): the unmodified reference env built from the written directory has the same drones, routes, buildings
    and map, and its reset observation equals the oracle's."""
    from uaisa_env.scenario import scenario_c3

    scn = scenario_c3(seed=4, n_scenarios=1, n_drones=6, n_buildings=3)
    d = scn.save_scenario_dir(str(tmp_path / "world_x"))
    back = ScenarioSet.from_json(d)
    assert np.array_equal(back.waypoints, scn.waypoints) and np.array_equal(back.buildings, scn.buildings) and np.array_equal(back.map_size, scn.map_size)
    env = ref_shim.make_reference_env(d, nm=10)
    g = env.ir_gym
    assert g.drone_num == scn.N and [float(x) for x in g.map_size] == [float(x) for x in scn.map_size[0]]
    assert np.allclose(np.array(g.building_list, dtype=np.float64).reshape(-1, 4), scn.buildings[0, : scn.n_buildings[0]])
    for k, dr in enumerate(g.drone_list):
        assert np.allclose(np.array(dr.waypoints, dtype=np.float64), scn.waypoints[0, k, : scn.n_points[0, k]])
    obs_ref, cnt_ref = ref_shim.pad_obs(env.drone_reset(False), 10)
    ob = oracle.OracleBatch(scn, 1, nm=10)
    ob.reset()
    obs, cnt, _ = ob.observe()
    assert close(obs, obs_ref, 1e-8) and np.array_equal(cnt, cnt_ref)
    y = scn.save_drone_paths_yaml(str(tmp_path / "drone_paths.yaml"))
    s2 = ScenarioSet.from_drone_paths_yaml(y, map_size=scn.map_size[0])
    assert np.array_equal(s2.waypoints[0, :, 0], scn.waypoints[0, :, 0]) and s2.N == scn.N

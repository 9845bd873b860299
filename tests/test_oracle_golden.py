"""CPU: the C restatement (oracle/uaisa_oracle.c) against golden traces recorded from the
unmodified reference (oracle/gen_golden.py).  Flags/counts bit-exact, floats to 1e-9 (in practice
<= 5e-16: only numpy-SIMD-vs-glibc libm ulps differ)."""
import glob
import json
import os

import numpy as np
import pytest

import _util  # noqa: F401  (path setup)
import oracle
from uaisa_env.scenario import ScenarioSet

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("policy_"))
STATE = ("pos", "vel", "yaw", "pitch", "wp_idx", "arrive_flag", "dest_flag", "real_len", "max_dev", "extra_len")
TOL = 1e-9


def close(a, b, tol=TOL):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore"):
        ok = same | (np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))
    return bool(np.all(ok))


def load(path):
    z = np.load(path)
    scn = ScenarioSet.from_dict(json.loads(str(z["scenario_json"])))
    return z, scn, int(z["nm"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_teacher_forced(path):
    z, scn, nm = load(path)
    ob = oracle.OracleBatch(scn, 1, nm=nm)
    ob.reset()
    obs, cnt, _ = ob.observe()
    assert close(obs, z["obs0"]) and np.array_equal(cnt, z["count0"])
    T = z["action"].shape[0]
    for t in range(T):
        ob.set_state({k: z["pre_" + k][t] for k in STATE})
        obs, rew, done, info, fin = ob.step(z["action"][t])
        assert np.array_equal(done, z["done"][t]), (t, "done")
        assert np.array_equal(info, z["info"][t]), (t, "info")
        assert np.array_equal(fin, z["finish"][t]), (t, "finish")
        assert np.array_equal(ob.count, z["count"][t]), (t, "count")
        assert close(obs, z["obs"][t]), (t, "obs")
        assert close(rew, z["reward"][t]), (t, "reward", rew, z["reward"][t])
        st = ob.get_state()
        for k in STATE:
            assert close(st[k], z["post_" + k][t]), (t, k)
        m = z["reset_mask"][t]
        if m.any():
            ob.reset(m)
            o2, c2, _ = ob.observe()
            assert close(o2, z["obs_after_reset"][t]) and np.array_equal(c2, z["count_after_reset"][t]), (t, "observe")


FREE = [p for p in GOLDEN if not os.path.basename(p).startswith("f32_")]      # (the f32_ traces re-load a rounded state every step: not one free trajectory)


@pytest.mark.parametrize("path", FREE, ids=[os.path.basename(p)[:-4] for p in FREE])
def test_oracle_free_running(path):
    """No teacher forcing: the oracle reproduces the whole recorded trajectory from reset."""
    z, scn, nm = load(path)
    ob = oracle.OracleBatch(scn, 1, nm=nm)
    ob.reset()
    T = z["action"].shape[0]
    for t in range(T):
        obs, rew, done, info, fin = ob.step(z["action"][t])
        assert np.array_equal(done, z["done"][t]) and np.array_equal(fin, z["finish"][t]), t
        assert np.array_equal(ob.count, z["count"][t]), t
        assert close(obs, z["obs"][t], 1e-7) and close(rew, z["reward"][t], 1e-7), t
        m = z["reset_mask"][t]
        if m.any():
            ob.reset(m)
    st = ob.get_state()
    for k in STATE:
        ref = z["post_" + k][-1].copy()
        if z["reset_mask"][-1].any():
            continue
        assert close(st[k], ref, 1e-7), k


def test_neighbour_indices_match_rows():
    """nbr[a, r] names the drone whose relative position is row r's [3:6] (new output)."""
    z, scn, nm = load([p for p in GOLDEN if "swarm12_nm10" in p][0])
    ob = oracle.OracleBatch(scn, 1, nm=nm)
    checked = 0
    for t in range(z["action"].shape[0]):
        ob.set_state({k: z["pre_" + k][t] for k in STATE})
        obs, *_ = ob.step(z["action"][t])
        pos = ob.get_state()["pos"]
        for a in range(ob.A):
            for r in range(ob.count[a]):
                j = ob.nbr[a, r]
                assert j >= 0 and j != a
                rel = np.round(pos[j] - pos[a], 2)
                assert np.allclose(obs[a, 12 + 9 * r + 3: 12 + 9 * r + 6], rel, atol=1e-12)
                checked += 1
            assert np.all(ob.nbr[a, ob.count[a]:] == -1)
    assert checked > 500

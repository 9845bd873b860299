"""-m gpu: the 1-env ``mdin`` adapter (reference list-of-numpy API) replays traces recorded from the reference."""
import glob
import json
import os

import numpy as np
import pytest

import _util  # noqa: F401

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, json.loads(str(z["scenario_json"])), int(z["nm"])


@pytest.mark.parametrize("name,horizon", [("world_3", 40), ("c1_yaml4", 60), ("world_8", 25), ("swarm12_nm3", 12)])
def test_free_running_replay_of_reference_trace(name, horizon):
    """Free-running from reset with the recorded actions and resets: identical flags and ragged shapes,
    observations / rewards equal to the reference's own doubles over a short horizon (fp32 state drift)."""
    from uaisa_env.drone_envs.mdin import mdin

    z, scn, nm = _load(name)
    env = mdin(world_name="x", neighbors_num=nm, scenario=scn)
    obs = env.drone_reset(False)
    assert len(obs) == scn["drone_num"] and all(o.shape == (21,) for o in obs)
    assert np.allclose(np.stack(obs), z["obs0"][:, :21], atol=1e-9)
    assert env.observation_space.shape == (21,) and env.action_space.shape == (3,)
    assert env.ir_gym.drone_num == scn["drone_num"] and env.ir_gym.acceler == 0.5
    nbad = 0
    for t in range(horizon):
        o, r, d, i, f = env.drone_step([a for a in z["action"][t]])
        assert isinstance(o, list) and isinstance(d, list) and isinstance(d[0], bool)
        assert d == [bool(x) for x in z["done"][t]] and f == [bool(x) for x in z["finish"][t]] and i == [bool(x) for x in z["info"][t]], t
        for k in range(len(o)):
            kk = int(z["count"][t][k])
            assert o[k].shape == (12 + 9 * max(kk, 1),), (t, k)
            nbad += int((np.abs(o[k] - z["obs"][t][k, : o[k].shape[0]]) > 1e-9).sum())
        rr = np.array(r, dtype=np.float64); ref = z["reward"][t]
        fin = np.isfinite(ref)
        nbad += int((np.abs(rr[fin] - ref[fin]) > 1e-9).sum()) + int((np.isfinite(rr) != fin).sum())
        m = z["reset_mask"][t]
        if z["full_reset"][t]:
            env.drone_reset(False)
        elif m.any():
            for k in np.nonzero(m)[0]:
                env.drone_reset_one(False, int(k))
            o2 = env.ir_gym.env_observation()
            assert [x.shape[0] for x in o2] == [12 + 9 * max(int(c), 1) for c in z["count_after_reset"][t]]
    assert nbad <= 3, nbad      # a 2-dp value may move by one quantum once the fp32 state has drifted
    v = env.ir_gym.drone_list[0].vel
    assert v.shape == (3,) and len(env.ir_gym.indicators_deviation()) == scn["drone_num"] and len(env.ir_gym.cal_des_list()) == scn["drone_num"]
    env.ir_gym.drone_list[0].rvo_vel = np.zeros(3)
    env.close()


def test_nonfinite_observation_raises_like_the_reference():
    from uaisa_env.drone_envs.mdin import mdin
    from uaisa_env.scenario import scenario_world3

    env = mdin(scenario=scenario_world3())
    env.drone_reset(False)
    env.drone_step([np.array([0.5, 0.1, 0.1])] * 3)
    with pytest.raises(ValueError):   # a NaN action propagates through np.clip into the velocity and the observation
        env.drone_step([np.array([np.nan, 0.0, 0.0]), np.array([0.1, 0.0, 0.0]), np.array([0.1, 0.0, 0.0])])
    env.close()
    with pytest.raises(ValueError):
        mdin()

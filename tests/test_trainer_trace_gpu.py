"""-m gpu: the drop-in claim, executed.  tests/golden/trainer_call_trace.json is the exact sequence of env-API calls the UNMODIFIED
reference trainer (multi_ppo.training_loop, train/policy/multi_ppo.py:166-339) issued during one epoch in the build container -- method,
argument types / shapes, returned types / shapes, the actions it passed (oracle/time_reference.py records it through a proxy).  Here the same
sequence drives the product's ``uaisa_env.drone_envs.mdin``: every call must be accepted and return the same structure; flags and ragged
observation lengths must match the reference's (free-running from reset on the recorded actions)."""
import json
import os

import numpy as np
import pytest

import _util  # noqa: F401

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _describe(x):
    if isinstance(x, (list, tuple)):
        return {"t": type(x).__name__, "items": [_describe(v) for v in x]}
    if isinstance(x, np.ndarray):
        return {"t": "ndarray", "shape": list(x.shape), "dtype": str(x.dtype)}
    if isinstance(x, (bool, np.bool_)):
        return {"t": "bool"}
    if isinstance(x, (int, np.integer)):
        return {"t": "int"}
    if isinstance(x, (float, np.floating)):
        return {"t": "float"}
    if x is None:
        return {"t": "None"}
    return {"t": type(x).__name__}


def _num(d):
    """int / float -> "num": the reference's max_deviation is the python int 0 until the first step and a float afterwards (drone.py:66, 261)."""
    if isinstance(d, dict):
        if d.get("t") in ("int", "float") or (d.get("t") == "ndarray" and int(np.prod(d.get("shape", [2]))) == 1):      # (... or a 1-element array: np.linalg.norm results kept un-squeezed)
            return {"t": "num"}
        return {k: _num(v) for k, v in d.items()}
    if isinstance(d, list):
        return [_num(v) for v in d]
    return d


def test_reference_trainer_call_sequence_replays_against_the_drop_in():
    from uaisa_env.drone_envs.mdin import mdin

    tr = json.load(open(os.path.join(GOLDEN, "trainer_call_trace.json")))
    env = mdin(world_name="x", neighbors_num=tr["nm"], scenario=tr["scenario"])
    n_calls = {"drone_step": 0, "drone_reset": 0, "drone_reset_one": 0, "ir_gym.env_observation": 0, "ir_gym.indicators_deviation": 0, "ir_gym.indicators_extra_len": 0}
    shape_bad = 0
    for e in tr["trace"]:
        if "get" in e:
            k = e["get"]
            if k == "observation_space.shape":
                assert list(env.observation_space.shape) == e["value"]
            elif k == "action_space.shape":
                assert list(env.action_space.shape) == e["value"]
            elif k == "ir_gym.drone_num":
                assert env.ir_gym.drone_num == e["value"]
            elif k == "ir_gym.acceler":
                assert env.ir_gym.acceler == e["value"]
            elif k.startswith("ir_gym.drone_list[i]."):
                v = getattr(env.ir_gym.drone_list[e["i"]], k.rsplit(".", 1)[1])
                assert _describe(np.squeeze(v)) == e["value"], (k, _describe(np.squeeze(v)), e["value"])
            continue
        if "set" in e:
            setattr(env.ir_gym.drone_list[e["i"]], e["set"].rsplit(".", 1)[1], np.zeros(3))
            continue
        c = e["call"]
        n_calls[c] += 1
        if c == "drone_step":
            out = env.drone_step([np.asarray(a, np.float64) for a in e["action"]])
            d = _describe(out)
            ref = e["ret"]
            assert d["t"] == ref["t"] and len(d["items"]) == 5
            for q in (1, 2, 3, 4):     # reward / done / info / finish lists: identical structure (python floats are np.float64 in both)
                assert _num(d["items"][q]) == _num(ref["items"][q]), (q, d["items"][q], ref["items"][q])
            assert [x["dtype"] for x in d["items"][0]["items"]] == [x["dtype"] for x in ref["items"][0]["items"]]
            assert [bool(x) for x in out[2]] == e["done"] and [bool(x) for x in out[3]] == e["info"] and [bool(x) for x in out[4]] == e["finish"]
            shape_bad += int([int(o.shape[0]) for o in out[0]] != e["obs_len"])
        elif c == "drone_reset":
            out = env.drone_reset(False)
            assert _describe(out) == e["ret"]
        elif c == "drone_reset_one":
            assert env.drone_reset_one(False, e["id"]) is None
        elif c == "ir_gym.env_observation":
            out = env.ir_gym.env_observation()
            assert _describe(out)["t"] == "list" and [x["dtype"] for x in _describe(out)["items"]] == [x["dtype"] for x in e["ret"]["items"]]
            shape_bad += int([int(o.shape[0]) for o in out] != e["obs_len"])
        elif c == "ir_gym.indicators_deviation":
            assert _num(_describe(env.ir_gym.indicators_deviation())) == _num(e["ret"])
        elif c == "ir_gym.indicators_extra_len":
            assert _num(_describe(env.ir_gym.indicators_extra_len())) == _num(e["ret"])
    assert n_calls["drone_step"] == 60 and n_calls["drone_reset"] >= 2 and n_calls["ir_gym.indicators_deviation"] == 60
    assert shape_bad == 0
    env.close()

"""CPU: the C-ABI library loads and exports every symbol include/uaisa_b200.h declares; host-side
argument validation works without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _util  # noqa: F401
from uaisa_env import _capi
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_c4, scenario_world3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "uaisa_b200.h")).read()
    declared = set(re.findall(r"\b(uaisa_[a-z_]+)\s*\(", hdr))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    L = _capi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.uaisa_version() >= 100


def test_default_config_and_validation():
    L = _capi.lib()
    c = _capi.default_config()
    assert c.nm == 10 and c.goal_threshold == 0.4 and c.neighbor_range == 10.0 and c.env_train == 1
    h = C.c_void_p()
    c.n_drones = 0
    assert L.uaisa_create(C.byref(c), C.byref(h)) == -1 and b"n_drones" in L.uaisa_last_error()
    c = _capi.default_config(); c.nm = 99
    assert L.uaisa_create(C.byref(c), C.byref(h)) == -1
    assert L.uaisa_create(None, C.byref(h)) == -1
    assert L.uaisa_step(None, None, None, None, None, None, None, None, None, None) == -1
    assert L.uaisa_destroy(None) == 0
    assert C.sizeof(_capi.UaisaConfig) == 10 * 4 + 11 * 8


def test_scenario_loaders(tmp_path):
    s = scenario_c1()
    assert s.N == 4 and s.W == 2 and s.n_buildings[0] == 0 and tuple(s.map_size[0]) == (50, 50, 10)
    d = s.to_dict()
    import json
    p = tmp_path / "data_1.json"
    p.write_text(json.dumps(d))
    s2 = ScenarioSet.from_json(str(tmp_path))
    assert np.array_equal(s.waypoints, s2.waypoints) and np.array_equal(s.n_points, s2.n_points)
    y = tmp_path / "drone_paths.yaml"
    y.write_text("start_points:\n  - [3, 3, 3]\n  - [45, 3, 5]\nend_points:\n  - [45, 45, 8]\n  - [3, 48, 3]\n")
    s3 = ScenarioSet.from_drone_paths_yaml(str(y))
    assert s3.N == 2 and np.array_equal(s3.waypoints[0, 1, 1], [3, 48, 3])
    c3 = scenario_c3(seed=0, n_scenarios=2)
    assert c3.S == 2 and c3.N == 16 and c3.n_buildings.min() >= 1 and c3.n_points.max() <= c3.W
    st = ScenarioSet.stack([scenario_c1(), ScenarioSet.from_start_end([[1, 1, 1]] * 4, [[2, 2, 2]] * 4, (5, 5, 5), buildings=[[1, 1, 1, 1], [2, 2, 2, 2]])])
    assert st.S == 2 and st.M == 2
    with pytest.raises(ValueError):
        ScenarioSet.stack([scenario_c1(), scenario_world3()])

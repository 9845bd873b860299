"""CPU: offline scenario tooling (uaisa_env.world.planner) -- the counterparts of the reference's
uaisa_env/world/{grid_3D_safe_zone,theta_star_3D,line_sight_partial_3D,path_planning_main,random_start_end}.py."""
import json
import os

import numpy as np
import pytest

import _util  # noqa: F401
from uaisa_env.scenario import ScenarioSet
from uaisa_env.world import line_of_sight, occupancy_from_buildings, plan_scenario, random_cylinder_map, random_start_end, theta_star


def _seg_clear_of_cylinders(p, q, buildings, margin=0.0):
    t = np.linspace(0, 1, 200)[:, None]
    pts = p[None] * (1 - t) + q[None] * t
    for bx, by, bh, br in buildings:
        inside = (np.hypot(pts[:, 0] - bx, pts[:, 1] - by) <= br + margin) & (pts[:, 2] <= bh)
        if inside.any():
            return False
    return True


def test_occupancy_and_line_of_sight():
    E, Es = occupancy_from_buildings([[5.0, 5.0, 4.0, 1.0]], (10, 10, 6), fence=False)
    assert E[5, 5, 0] == 1 and E[5, 5, 3] == 1 and E[5, 5, 4] == 0 and E[0, 0, 0] == 0
    assert Es[5, 5, 4] == 0.5 and Es[1, 1, 1] == 0.0 and (Es >= E).all()
    assert not line_of_sight(Es, (1, 5, 1), (9, 5, 1)) and line_of_sight(Es, (1, 5, 5.5), (9, 5, 5.5)) and line_of_sight(Es, (1, 1, 1), (9, 1, 1))
    Ef, _ = occupancy_from_buildings([], (6, 6, 4), fence=True)
    assert Ef[0].all() and Ef[:, :, -1].all() and Ef[2, 2, 2] == 0


def test_theta_star_goes_around_a_wall_and_straight_when_free():
    E = np.zeros((12, 12, 6))
    p = theta_star(E, (1, 1, 1), (10, 9, 4))
    assert p.shape[0] == 2 and np.array_equal(p[0], [1, 1, 1]) and np.array_equal(p[-1], [10, 9, 4])      # any-angle: one straight leg
    E[6, 0:9, :] = 1.0                                                                                   # a wall with a gap at y >= 9
    p = theta_star(E, (1, 1, 1), (10, 2, 1))
    assert p is not None and 3 <= p.shape[0] <= 6
    assert all(line_of_sight(E, a, b) for a, b in zip(p[:-1], p[1:]))
    direct = np.linalg.norm(p[-1] - p[0]); length = sum(np.linalg.norm(b - a) for a, b in zip(p[:-1], p[1:]))
    assert direct < length < 2.6 * direct
    E[6, :, :] = 1.0
    assert theta_star(E, (1, 1, 1), (10, 2, 1)) is None


def test_plan_scenario_with_cylinders_yields_loadable_obstacle_aware_routes(tmp_path):
    buildings = [[5.0, 5.0, 5.0, 0.8], [3.0, 6.5, 5.0, 0.6], [7.0, 3.5, 4.0, 0.7]]
    ang = 2 * np.pi * np.arange(8) / 8
    starts = np.round(np.stack([5 + 4 * np.cos(ang), 5 + 4 * np.sin(ang), np.full(8, 3.0)], 1), 2)
    ends = np.round(np.stack([5 - 4 * np.cos(ang), 5 - 4 * np.sin(ang), np.full(8, 2.0)], 1), 2)
    scn = plan_scenario(starts, ends, (10, 10, 6), buildings=buildings, save_dir=str(tmp_path))
    assert scn.N == 8 and scn.n_points.min() >= 2 and scn.n_points.max() >= 3          # somebody had to go around the centre cylinder
    for k in range(8):
        w = scn.waypoints[0, k, : scn.n_points[0, k]]
        assert np.allclose(w[0], starts[k]) and np.allclose(w[-1], ends[k])
        assert all(_seg_clear_of_cylinders(a, b, buildings, margin=0.2) for a, b in zip(w[:-1], w[1:])), k      # drone radius 0.2 clears every cylinder
    d = json.load(open(os.path.join(str(tmp_path), "data_1.json")))
    assert set(d) == {"drone_num", "map_size", "waypoints_list", "n_points_list", "building_list"}
    s2 = ScenarioSet.from_json(str(tmp_path))
    assert np.array_equal(s2.waypoints, scn.waypoints) and np.load(os.path.join(str(tmp_path), "E3d_safe.npy")).shape == (10, 10, 6)
    # the env accepts it (host-side validation only; the GPU suite steps planned scenarios in test_planner_gpu)
    scn.validate()


def test_random_map_and_random_start_end():
    starts, ends = random_start_end(4, (20, 20, 5), 6.0, seed=3)
    allp = np.concatenate([starts, ends])
    dmat = np.linalg.norm(allp[:, None] - allp[None], axis=-1) + np.eye(8) * 100
    assert dmat.min() >= 6.0
    E, Es, bl = random_cylinder_map((20, 20, 8), starts + [0, 0, 1], ends + [0, 0, 1], seed=5, k_sigma=1.5)
    assert len(bl) > 3 and all(3 <= b[2] <= 8 and b[3] in (0.0, 1.0) for b in bl)
    for p in np.concatenate([starts, ends]):
        i, j = int(np.ceil(p[0])), int(np.ceil(p[1]))
        assert not any(abs(b[0] - 0.5 - i) <= 1 and abs(b[1] - 0.5 - j) <= 1 for b in bl)          # the (2 n_low + 1)^2 block around it is clear
    with pytest.raises(ValueError):
        random_start_end(50, (5, 5, 2), 4.0, seed=0, max_tries=2000)

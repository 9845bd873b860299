"""Scenario constants for the batched env: loaders for the reference's on-disk formats and the
synthetic generators behind BASELINE.json's configs.

Formats kept from the reference:
  * ``data_1.json``  -- ``env_base.load_data`` (uaisa_env/drone_envs/env_base.py:26-39): keys
    ``drone_num, map_size, waypoints_list[N][W_i][3], n_points_list[N], building_list[M][x,y,h,r]``.
  * ``drone_paths.yaml`` -- ``train/drone_paths.yaml:1-12``: ``start_points`` / ``end_points``
    (N x [x,y,z]); the reference only feeds it to its offline planner
    (uaisa_env/world/path_planning_main.py:20-46); here it yields straight 2-point routes.

A :class:`ScenarioSet` is S scenarios of N drones each, padded to W waypoints and M buildings, in
float64 -- scenario constants are shared by many envs, stay L2-resident on the GPU and are kept in
double so a json scenario is represented exactly as the reference sees it.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

DEFAULT_RADIUS = 0.2    # drone.py:15  radius=0.2
DEFAULT_PRIORITY = 5.0  # drone.py:14  priority=5


@dataclass
class ScenarioSet:
    waypoints: np.ndarray    # [S,N,W,3] float64, rows >= n_points are padding (copies of the last point)
    n_points: np.ndarray     # [S,N] int32
    radius: np.ndarray       # [S,N] float64
    priority: np.ndarray     # [S,N] float64
    buildings: np.ndarray    # [S,M,4] float64 (x, y, h, r)
    n_buildings: np.ndarray  # [S] int32
    map_size: np.ndarray     # [S,3] float64

    @property
    def S(self) -> int:
        return int(self.waypoints.shape[0])

    @property
    def N(self) -> int:
        return int(self.waypoints.shape[1])

    @property
    def W(self) -> int:
        return int(self.waypoints.shape[2])

    @property
    def M(self) -> int:
        return int(self.buildings.shape[1])

    def validate(self) -> "ScenarioSet":
        S, N, W = self.S, self.N, self.W
        assert self.waypoints.shape == (S, N, W, 3) and self.waypoints.dtype == np.float64
        assert self.n_points.shape == (S, N) and self.n_points.dtype == np.int32
        if np.any(self.n_points < 2) or np.any(self.n_points > W):
            raise ValueError("every drone needs 2 <= n_points <= W waypoints (drone.py:46-48 reads waypoints[1])")
        assert self.radius.shape == (S, N) and self.priority.shape == (S, N)
        assert self.buildings.shape[0] == S and self.buildings.shape[2] == 4
        assert self.n_buildings.shape == (S,) and self.map_size.shape == (S, 3)
        if np.any(self.n_buildings < 0) or np.any(self.n_buildings > self.M):
            raise ValueError("n_buildings out of range")
        return self

    # -- data_1.json ---------------------------------------------------------------------------
    @staticmethod
    def from_dict(d: dict, radius: float = DEFAULT_RADIUS, priority: float = DEFAULT_PRIORITY) -> "ScenarioSet":
        n = int(d["drone_num"])
        wl = d["waypoints_list"][:n]
        npl = [int(x) for x in d["n_points_list"][:n]]
        if len(wl) != n or len(npl) != n:
            raise ValueError("waypoints_list / n_points_list shorter than drone_num")
        W = max(max(npl), max(len(w) for w in wl), 2)
        wp = np.zeros((1, n, W, 3), dtype=np.float64)
        for k, w in enumerate(wl):
            w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
            if npl[k] != w.shape[0]:
                # the reference takes destination = waypoints[-1] but counts with n_points
                # (env_drones.py:34-41 vs drone.py:117); they agree in every shipped scenario.
                raise ValueError("n_points_list[%d]=%d != len(waypoints)=%d" % (k, npl[k], w.shape[0]))
            wp[0, k, : w.shape[0]] = w
            wp[0, k, w.shape[0]:] = w[-1]
        bl = np.asarray(d.get("building_list", []), dtype=np.float64).reshape(-1, 4)
        M = max(int(bl.shape[0]), 1)
        b = np.zeros((1, M, 4), dtype=np.float64)
        b[0, : bl.shape[0]] = bl
        return ScenarioSet(
            waypoints=wp, n_points=np.asarray([npl], dtype=np.int32),
            radius=np.full((1, n), radius, dtype=np.float64), priority=np.full((1, n), priority, dtype=np.float64),
            buildings=b, n_buildings=np.asarray([bl.shape[0]], dtype=np.int32),
            map_size=np.asarray([d["map_size"]], dtype=np.float64).reshape(1, 3),
        ).validate()

    @staticmethod
    def from_json(path: str, **kw) -> "ScenarioSet":
        if os.path.isdir(path):
            path = os.path.join(path, "data_1.json")
        with open(path, "r") as f:
            return ScenarioSet.from_dict(json.load(f), **kw)

    def to_dict(self, s: int = 0) -> dict:
        """Back to the data_1.json schema (one scenario)."""
        return {
            "drone_num": self.N,
            "map_size": [float(x) for x in self.map_size[s]],
            "waypoints_list": [self.waypoints[s, k, : self.n_points[s, k]].tolist() for k in range(self.N)],
            "n_points_list": [int(x) for x in self.n_points[s]],
            "building_list": self.buildings[s, : self.n_buildings[s]].tolist(),
        }

    def save_scenario_dir(self, path: str, s: int = 0) -> str:
        """Write scenario ``s`` as a directory the REFERENCE can load (``env_base.load_data``, env_base.py:26-47):
        ``data_1.json`` plus the two occupancy grids it opens next to it (``E3d.npy`` / ``E3d_safe.npy``, only used by the
        reference's plotting and offline planner; written as empty 1 m grids of the map's extent)."""
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "data_1.json"), "w") as f:
            json.dump(self.to_dict(s), f)
        shape = tuple(int(max(1, np.ceil(x))) for x in self.map_size[s])
        for name in ("E3d.npy", "E3d_safe.npy"):
            np.save(os.path.join(path, name), np.zeros(shape, dtype=np.int8))
        return path

    def save_drone_paths_yaml(self, path: str, s: int = 0) -> str:
        """Start and end points of scenario ``s`` in the ``drone_paths.yaml`` schema (train/drone_paths.yaml:1-12), the input of
        the reference's offline planner (path_planning_main.py:20-46)."""
        import yaml

        starts = [self.waypoints[s, k, 0].tolist() for k in range(self.N)]
        ends = [self.waypoints[s, k, self.n_points[s, k] - 1].tolist() for k in range(self.N)]
        with open(path, "w") as f:
            yaml.safe_dump({"start_points": starts, "end_points": ends}, f)
        return path

    # -- drone_paths.yaml ----------------------------------------------------------------------
    @staticmethod
    def from_drone_paths_yaml(path: str, map_size: Sequence[float] = (50, 50, 10), **kw) -> "ScenarioSet":
        import yaml

        with open(path, "r") as f:
            y = yaml.safe_load(f)
        return ScenarioSet.from_start_end(y["start_points"], y["end_points"], map_size, **kw)

    @staticmethod
    def from_start_end(starts, ends, map_size, buildings=(), **kw) -> "ScenarioSet":
        starts = np.asarray(starts, dtype=np.float64).reshape(-1, 3)
        ends = np.asarray(ends, dtype=np.float64).reshape(-1, 3)
        n = starts.shape[0]
        d = {
            "drone_num": n, "map_size": list(map_size),
            "waypoints_list": [[starts[k].tolist(), ends[k].tolist()] for k in range(n)],
            "n_points_list": [2] * n, "building_list": [list(b) for b in buildings],
        }
        return ScenarioSet.from_dict(d, **kw)

    # -- combination ---------------------------------------------------------------------------
    @staticmethod
    def stack(sets: Iterable["ScenarioSet"]) -> "ScenarioSet":
        sets = list(sets)
        N = sets[0].N
        if any(s.N != N for s in sets):
            raise ValueError("all scenarios of a batch must have the same drone count")
        W = max(s.W for s in sets)
        M = max(s.M for s in sets)
        wps, bls = [], []
        for s in sets:
            w = np.zeros((s.S, N, W, 3))
            w[:, :, : s.W] = s.waypoints
            w[:, :, s.W:] = s.waypoints[:, :, -1:, :]
            wps.append(w)
            b = np.zeros((s.S, M, 4))
            b[:, : s.M] = s.buildings
            bls.append(b)
        cat = lambda xs: np.concatenate(xs, axis=0)  # noqa: E731
        return ScenarioSet(cat(wps), cat([s.n_points for s in sets]), cat([s.radius for s in sets]),
                           cat([s.priority for s in sets]), cat(bls), cat([s.n_buildings for s in sets]),
                           cat([s.map_size for s in sets])).validate()


# ----------------------------------------------------------------------------------------------
# Synthetic scenarios for BASELINE.json configs (SURVEY.md section 8d "Synthetic inputs")
# ----------------------------------------------------------------------------------------------
C1_STARTS = [[3, 3, 3], [45, 3, 5], [27, 48, 3], [43, 25, 6]]    # train/drone_paths.yaml:1-5
C1_ENDS = [[45, 45, 8], [3, 48, 3], [26, 3, 3], [5, 25, 3]]      # train/drone_paths.yaml:8-12


def scenario_c1() -> ScenarioSet:
    """configs[0]/[1]/[4]: the 4-drone train/drone_paths.yaml scenario, straight routes, 50x50x10."""
    return ScenarioSet.from_start_end(C1_STARTS, C1_ENDS, (50, 50, 10))


def scenario_world3() -> ScenarioSet:
    """Shipped uaisa_env/world/world_3/data_1.json (N=3, 10x10x5, no buildings)."""
    starts = [[5.0, 0.5, 1.0], [9.5, 9.5, 1.0], [9.5, 0.5, 1.0]]
    ends = [[5.0, 9.5, 4.0], [0.5, 0.5, 4.0], [0.5, 9.5, 4.0]]
    return ScenarioSet.from_start_end(starts, ends, (10, 10, 5))


def scenario_c3(seed: int = 0, n_scenarios: int = 1, n_drones: int = 16, n_buildings: int = 8, planned: bool = False) -> ScenarioSet:
    """configs[2]: N=16 on a ring of radius 4 about (5,5), z 5 -> antipodal goals at z 2, one or two
    intermediate waypoints, M cylinders (x,y~U(2,8), h~U(3,5), r~U(0.5,1.0)) clear of every waypoint.
    planned=True: the same starts / goals / cylinder statistics, but the routes come from the offline Theta* planner
    (uaisa_env.world.plan_scenario): obstacle-aware legs that keep 0.35 from every cylinder."""
    rng = np.random.default_rng(seed)
    out = []
    if planned:
        from .world import plan_scenario
        for _ in range(n_scenarios):
            ang = 2 * np.pi * (np.arange(n_drones) + rng.uniform(0, 1)) / n_drones
            starts = np.round(np.stack([5 + 4 * np.cos(ang), 5 + 4 * np.sin(ang), np.full(n_drones, 4.0)], axis=1), 2)
            ends = np.round(np.stack([5 - 4 * np.cos(ang), 5 - 4 * np.sin(ang), np.full(n_drones, 2.0)], axis=1), 2)
            ends_all = np.concatenate([starts, ends])
            bl = []
            while len(bl) < n_buildings:
                x, y = rng.uniform(2.5, 7.5, 2)
                h, r = rng.uniform(3, 5), rng.uniform(0.4, 0.8)
                if np.min(np.hypot(ends_all[:, 0] - x, ends_all[:, 1] - y)) < r + 0.8:
                    continue
                bl.append([round(float(x), 2), round(float(y), 2), round(float(h), 2), round(float(r), 2)])
            out.append(plan_scenario(starts, ends, (10, 10, 5), buildings=bl))
        return ScenarioSet.stack(out)
    for _ in range(n_scenarios):
        ang = 2 * np.pi * (np.arange(n_drones) + rng.uniform(0, 1)) / n_drones
        starts = np.stack([5 + 4 * np.cos(ang), 5 + 4 * np.sin(ang), np.full(n_drones, 5.0)], axis=1)
        ends = np.stack([5 - 4 * np.cos(ang), 5 - 4 * np.sin(ang), np.full(n_drones, 2.0)], axis=1)
        wl, npl = [], []
        for k in range(n_drones):
            nmid = int(rng.integers(1, 3))
            pts = [starts[k]]
            for m in range(nmid):
                f = (m + 1) / (nmid + 1)
                mid = starts[k] * (1 - f) + ends[k] * f + rng.uniform(-1.0, 1.0, 3) * np.array([1, 1, 0.3])
                pts.append(np.clip(mid, [0.5, 0.5, 0.5], [9.5, 9.5, 4.5]))
            pts.append(ends[k])
            wl.append(np.round(np.asarray(pts), 2).tolist())
            npl.append(len(pts))
        allw = np.concatenate([np.asarray(w) for w in wl], axis=0)
        bl = []
        tries = 0
        margin = 0.3   # clearance between a cylinder's surface and every waypoint; relaxed if the map is too crowded
        while len(bl) < n_buildings:
            tries += 1
            if tries % 2000 == 0:
                margin *= 0.5
            x, y = rng.uniform(2, 8, 2)
            h, r = rng.uniform(3, 5), rng.uniform(0.5, 1.0)
            if np.min(np.hypot(allw[:, 0] - x, allw[:, 1] - y)) < r + margin:
                continue
            bl.append([round(float(x), 2), round(float(y), 2), round(float(h), 2), round(float(r), 2)])
        out.append(ScenarioSet.from_dict({"drone_num": n_drones, "map_size": [10, 10, 5], "waypoints_list": wl,
                                          "n_points_list": npl, "building_list": bl}))
    return ScenarioSet.stack(out)


def scenario_c4(seed: int = 0, n_scenarios: int = 1, n_drones: int = 256, min_sep: float = 0.5) -> ScenarioSet:
    """configs[3]: N=256 dense airspace, 50x50x10, uniform starts/goals with >= min_sep separation."""
    rng = np.random.default_rng(seed)
    out = []

    def sample(n):
        pts = np.zeros((0, 3))
        while pts.shape[0] < n:
            c = rng.uniform([1, 1, 1], [49, 49, 9], size=(n, 3))
            for p in c:
                if pts.shape[0] == 0 or np.min(np.linalg.norm(pts - p, axis=1)) >= min_sep:
                    pts = np.vstack([pts, p])
                    if pts.shape[0] == n:
                        break
        return np.round(pts, 2)

    for _ in range(n_scenarios):
        out.append(ScenarioSet.from_start_end(sample(n_drones), sample(n_drones), (50, 50, 10)))
    return ScenarioSet.stack(out)

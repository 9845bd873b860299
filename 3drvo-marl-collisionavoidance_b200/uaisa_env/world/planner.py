"""Offline scenario tooling: obstacle-aware multi-waypoint routes for the env's scenarios.

What the reference does offline (uaisa_env/world/): ``grid_3D_safe_zone.py:14`` draws random cylinders into an occupancy grid and pads them
with a safety zone (occupied = 1, free cell next to an occupied one = 0.5); ``theta_star_3D.py:5`` runs Theta* (A* over the 26-neighbourhood whose
parent pointer skips to the grand-parent whenever ``line_sight_partial_3D.py:3`` sees a free line) with cost
``F = kg*G + kh*H + ke*E_safe[cell]``; ``path_planning_main.py:172-205`` reads ``drone_paths.yaml`` and writes ``data_1.json`` + the grids;
``random_start_end.py`` draws start / end pairs.  Those files are module-level scripts with hard-coded Windows paths and matplotlib calls; this is
the same tool chain as importable functions (same inputs, same outputs and file formats, own implementation -- in particular a binary-heap open
list instead of the reference's O(n) array scans, and a sampled supercover line-of-sight test).  It runs on the host: routes are scenario
CONSTANTS, built once; the live path needs only the formats.

Conventions: grids are indexed ``E[ix, iy, iz]`` over 1 m cells covering ``[0, X] x [0, Y] x [0, Z]`` (cell i spans [i, i+1)); buildings are the
env's cylinders ``[x, y, h, r]`` (rvo_inter.py:198-209 tests ``z <= h`` and ``d2D <= r_drone + r``).
"""
from __future__ import annotations

import heapq
import math
from typing import Optional, Sequence

import numpy as np

from ..scenario import ScenarioSet


# ---------------------------------------------------------------------------------------------------------------- maps
def occupancy_from_buildings(buildings: Sequence[Sequence[float]], map_size: Sequence[float], clearance: float = 0.0, fence: bool = True):
    """Cylinders -> (E3d, E3d_safe): E3d marks every cell whose CENTRE lies inside a cylinder grown by ``clearance``; E3d_safe additionally gives
    0.5 to free cells that touch an occupied one (grid_3D_safe_zone.py:119-146) and, with ``fence``, closes the map's outer shell
    (:152-166 "electronic fence")."""
    X, Y, Z = (int(math.ceil(v)) for v in map_size)
    E = np.zeros((X, Y, Z), np.float64)
    cx, cy = np.meshgrid(np.arange(X) + 0.5, np.arange(Y) + 0.5, indexing="ij")
    for bx, by, bh, br in buildings:
        foot = np.hypot(cx - bx, cy - by) <= br + clearance
        top = int(min(Z, math.ceil(bh + clearance)))
        if top > 0:
            E[:, :, :top][foot] = 1.0
    Es = E.copy()
    occ = E > 0
    pad = np.zeros((X + 2, Y + 2, Z + 2), bool)
    pad[1:-1, 1:-1, 1:-1] = occ
    near = np.zeros_like(occ)
    for dx in (0, 1, 2):
        for dy in (0, 1, 2):
            for dz in (0, 1, 2):
                near |= pad[dx:dx + X, dy:dy + Y, dz:dz + Z]
    Es[near & ~occ] = 0.5
    if fence:
        for A in (E, Es):
            A[[0, -1], :, :] = 1.0; A[:, [0, -1], :] = 1.0; A[:, :, [0, -1]] = 1.0
    return E, Es


def random_cylinder_map(map_size: Sequence[int], starts, ends, h: Optional[float] = None, k_sigma: float = 2.0, n_low: int = 1, seed: Optional[int] = None):
    """Random cylinder field in the manner of grid_3D_safe_zone.py:14-118: cell (i, j) becomes a cylinder centre when a standard normal draw
    exceeds ``k_sigma`` (density ~2.3 % at 2), heights ~ N(0.8 h, 0.5 h) clipped to [3, Z], radii 0 or 1 cells, and a (2 n_low + 1)^2 block around
    every start and end point is kept clear.  -> (E3d, E3d_safe, building_list [x, y, h, r])."""
    rng = np.random.default_rng(seed)
    X, Y, Z = (int(v) for v in map_size)
    starts = np.asarray(starts, np.float64).reshape(-1, 3); ends = np.asarray(ends, np.float64).reshape(-1, 3)
    if h is None:
        h = float(max(ends[:, 2].max(), starts[:, 2].max()))
    centre = rng.normal(0.0, 1.0, (X, Y)) > k_sigma
    for p in np.concatenate([starts, ends]):
        i, j = int(math.ceil(p[0])), int(math.ceil(p[1]))
        centre[max(0, i - n_low - 1): i + n_low + 2, max(0, j - n_low - 1): j + n_low + 2] = False
    buildings = []
    for i, j in np.argwhere(centre):
        hh = int(np.clip(round(rng.normal(0.8 * h, 0.5 * h)), 3, Z))
        buildings.append([float(i) + 0.5, float(j) + 0.5, float(hh), float(rng.integers(0, 2))])
    E, Es = occupancy_from_buildings(buildings, (X, Y, Z), fence=True)
    return E, Es, buildings


def random_start_end(n_pairs: int, map_size: Sequence[int], min_distance: float, seed: Optional[int] = None, max_tries: int = 100000):
    """Start / end pairs on the integer lattice, every two points of different pairs and the two ends of one pair at least ``min_distance``
    apart (random_start_end.py:18-52).  -> (starts [n, 3], ends [n, 3]); write them with ScenarioSet.save_drone_paths_yaml."""
    rng = np.random.default_rng(seed)
    X, Y, Z = (int(v) for v in map_size)
    pts = []
    draw = lambda: np.array([rng.integers(0, X), rng.integers(0, Y), rng.integers(0, Z)], np.float64)  # noqa: E731
    far = lambda p, q: np.linalg.norm(p - q) >= min_distance  # noqa: E731
    tries = 0
    clear = lambda p: all(far(p, a) and far(p, b) for a, b in pts)  # noqa: E731

    def place(extra=None):
        nonlocal tries
        while True:
            tries += 1
            if tries > max_tries:
                raise ValueError("could not place %d pairs %.1f apart in a %dx%dx%d map" % (n_pairs, min_distance, X, Y, Z))
            p = draw()
            if clear(p) and (extra is None or far(p, extra)):
                return p
    while len(pts) < n_pairs:
        s = place()
        pts.append((s, place(s)))
    return np.stack([p[0] for p in pts]), np.stack([p[1] for p in pts])


# ---------------------------------------------------------------------------------------------------------------- line of sight + Theta*
def line_of_sight(E: np.ndarray, p, q, block: float = 1.0) -> bool:
    """True when the straight segment p -> q (continuous coordinates) crosses no cell with E >= ``block`` (role of line_sight_partial_3D.py:3).
    The segment is sampled every quarter cell and each sample's cell is tested together with the cells a point within 1e-9 of a face also touches."""
    p = np.asarray(p, np.float64); q = np.asarray(q, np.float64)
    n = int(math.ceil(np.linalg.norm(q - p) * 4.0)) + 1
    t = np.linspace(0.0, 1.0, n + 1)[:, None]
    pts = p[None, :] * (1.0 - t) + q[None, :] * t
    hi = np.array(E.shape) - 1
    for eps in (-1e-9, 1e-9):
        idx = np.clip(np.floor(pts + eps).astype(np.int64), 0, hi)
        if np.any(E[idx[:, 0], idx[:, 1], idx[:, 2]] >= block):
            return False
    return True


def cylinder_line_of_sight(buildings, margin: float):
    """-> los(p, q): the exact-geometry counterpart of :func:`line_of_sight` when the obstacles are known cylinders [x, y, h, r]: the segment is
    sampled every 0.05 and must stay ``margin`` away from every cylinder (horizontally beyond r + margin, or above h + margin)."""
    b = np.asarray(buildings, np.float64).reshape(-1, 4)

    def los(p, q):
        p = np.asarray(p, np.float64); q = np.asarray(q, np.float64)
        if b.shape[0] == 0:
            return True
        n = int(math.ceil(np.linalg.norm(q - p) / 0.05)) + 1
        t = np.linspace(0.0, 1.0, n + 1)[:, None]
        pts = p[None, :] * (1.0 - t) + q[None, :] * t
        d = np.hypot(pts[:, None, 0] - b[None, :, 0], pts[:, None, 1] - b[None, :, 1])
        return not bool(((d <= b[None, :, 3] + margin) & (pts[:, None, 2] <= b[None, :, 2] + margin)).any())
    return los


def theta_star(E_safe: np.ndarray, start, goal, K: Optional[Sequence[float]] = None, block: float = 1.0, los=None):
    """3-D Theta* on cell corners (theta_star_3D.py:5-124): nodes are integer lattice points, 26 neighbours, cost ``F = kg*G + kh*H + ke*E_safe`` with the
    reference's default K = (1, 1.25, |goal - start|) (path_planning_main.py:113-119); a neighbour takes the current node's PARENT as its parent
    whenever that parent sees it (any-angle paths).  start / goal are rounded up to lattice points like the reference (math.ceil, :104-110).
    ``los(p, q)`` overrides the grid line-of-sight test (e.g. :func:`cylinder_line_of_sight`); the grid then only prices the safety zone.
    -> path [n_points, 3] float64 from start to goal, or None if the goal cannot be reached."""
    X, Y, Z = E_safe.shape
    s = tuple(int(min(max(math.ceil(v), 0), m)) for v, m in zip(start, (X, Y, Z)))
    g = tuple(int(min(max(math.ceil(v), 0), m)) for v, m in zip(goal, (X, Y, Z)))
    kg, kh, ke = K if K is not None else (1.0, 1.25, math.dist(s, g))

    def cell_cost(n):      # the safety value of the cells around a lattice point (max of the up-to-8 cells it touches)
        i, j, k = n
        sl = E_safe[max(i - 1, 0): min(i + 1, X), max(j - 1, 0): min(j + 1, Y), max(k - 1, 0): min(k + 1, Z)]
        return float(sl.max()) if sl.size else 1.0

    if los is None:
        los = lambda a, b: line_of_sight(E_safe, a, b, block)  # noqa: E731
        free = lambda n: cell_cost(n) < block or n == s or n == g  # noqa: E731
    else:
        free = lambda n: True  # noqa: E731   (a node inside an obstacle is unreachable: no leg into it has line of sight)

    G = {s: 0.0}
    parent = {s: s}
    openq = [(kh * math.dist(s, g), 0, s)]
    closed = set()
    tie = 1
    steps = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) != (0, 0, 0)]
    while openq:
        _, _, cur = heapq.heappop(openq)
        if cur in closed:
            continue
        if cur == g:
            break
        closed.add(cur)
        par = parent[cur]
        for d in steps:
            nb = (cur[0] + d[0], cur[1] + d[1], cur[2] + d[2])
            if not (0 <= nb[0] <= X and 0 <= nb[1] <= Y and 0 <= nb[2] <= Z) or nb in closed or not free(nb):
                continue
            if los(par, nb):          # path 2: skip the current node (theta_star_3D.py:72-85)
                cand, via = G[par] + math.dist(par, nb), par
            elif los(cur, nb):        # path 1: ordinary A* edge (:86-95)
                cand, via = G[cur] + math.dist(cur, nb), cur
            else:
                continue
            if cand < G.get(nb, math.inf):
                G[nb] = cand; parent[nb] = via
                heapq.heappush(openq, (kg * cand + kh * math.dist(nb, g) + ke * min(cell_cost(nb), 0.5), tie, nb)); tie += 1
    if g not in parent:
        return None
    path = [g]
    while path[-1] != s:
        path.append(parent[path[-1]])
    return np.asarray(path[::-1], np.float64)


# ---------------------------------------------------------------------------------------------------------------- drivers
def plan_routes(starts, ends, E_safe: np.ndarray, max_waypoints: int = 0, los=None):
    """One Theta* route per (start, end) pair.  The planner works on lattice points (the reference rounds start and end up, path_planning_main.py:104-110);
    the exact start / end are put in front of / behind the lattice route and the lattice end points dropped again where the exact ones see past them.
    -> (waypoints_list, n_points_list)"""
    wl, npl = [], []
    see = los if los is not None else (lambda a, b: line_of_sight(E_safe, a, b))
    for p0, p1 in zip(np.asarray(starts, np.float64).reshape(-1, 3), np.asarray(ends, np.float64).reshape(-1, 3)):
        lat = theta_star(E_safe, p0, p1, los=los)
        if lat is None:
            raise ValueError("no route from %s to %s" % (p0.tolist(), p1.tolist()))
        pts = [p0] + [q for q in lat] + [p1]
        while len(pts) > 2 and (np.array_equal(pts[0], pts[1]) or see(pts[0], pts[2])):
            del pts[1]
        while len(pts) > 2 and (np.array_equal(pts[-1], pts[-2]) or see(pts[-3], pts[-1])):
            del pts[-2]
        path = np.stack(pts)
        if max_waypoints and path.shape[0] > max_waypoints:
            raise ValueError("route needs %d waypoints (> %d)" % (path.shape[0], max_waypoints))
        wl.append(path.tolist()); npl.append(int(path.shape[0]))
    return wl, npl


def plan_scenario(starts, ends, map_size: Sequence[int], buildings=None, seed: Optional[int] = None, clearance: float = 0.35, save_dir: Optional[str] = None, **map_kw):
    """yaml-style start / end points (+ cylinders, or a random cylinder map when ``buildings`` is None) -> a ScenarioSet with obstacle-aware
    multi-waypoint routes: what path_planning_main.py:172-205 produces.  With ``save_dir`` the result is written as a directory the
    reference loads (data_1.json + E3d.npy + E3d_safe.npy, env_base.py:26-47).
    ``clearance``: distance every leg keeps from every cylinder (drone radius 0.2 + slack); with given cylinders the legs are checked against the
    exact geometry, with a random map against its occupancy grid."""
    if buildings is None:
        E, Es, buildings = random_cylinder_map(map_size, starts, ends, seed=seed, **map_kw)
        wl, npl = plan_routes(starts, ends, Es)
    else:
        E, Es = occupancy_from_buildings(buildings, map_size, clearance=clearance, fence=False)
        wl, npl = plan_routes(starts, ends, Es, los=cylinder_line_of_sight(buildings, clearance))
    d = {"drone_num": len(wl), "map_size": [float(v) for v in map_size], "waypoints_list": wl, "n_points_list": npl, "building_list": [list(map(float, b)) for b in buildings]}
    scn = ScenarioSet.from_dict(d)
    if save_dir is not None:
        import os
        scn.save_scenario_dir(save_dir)
        np.save(os.path.join(save_dir, "E3d.npy"), E); np.save(os.path.join(save_dir, "E3d_safe.npy"), Es)
    return scn

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference env
(imported from /root/reference through oracle/ref_shim.py) -- run in the build container:

    python oracle/gen_golden.py

Each fixture = scenario (data_1.json schema, as a json string) + the trace recorded by
oracle/lockstep.run_reference (pre-state, action, outputs, post-state, resets, post-reset obs).
The fixtures pin oracle/uaisa_oracle.c (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]

import lockstep  # noqa: E402
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3  # noqa: E402

REF_WORLDS = os.path.join(os.environ.get("UAISA_REFERENCE_ROOT", "/root/reference"), "uaisa_env", "world")


def swarm(n=12, rad=1.0, seed=0):
    """Dense crossing scenario: produces multi-row, sorted, nm-truncated observations."""
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(n) / n
    st = np.stack([5 + rad * np.cos(ang), 5 + rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    en = np.stack([5 - rad * np.cos(ang), 5 - rad * np.sin(ang), 2.5 + 0.3 * rng.uniform(-1, 1, n)], 1)
    return ScenarioSet.from_start_end(np.round(st, 2), np.round(en, 2), (10, 10, 5), buildings=[[5, 5, 1.5, 0.3]])


CASES = [
    # name, scenario dict, nm, steps, seed, kwargs
    ("world_3", json.load(open(os.path.join(REF_WORLDS, "world_3", "data_1.json"))), 10, 400, 0, {}),
    ("world_8", json.load(open(os.path.join(REF_WORLDS, "world_8", "data_1.json"))), 10, 300, 1, {}),
    ("world_8_nm2", json.load(open(os.path.join(REF_WORLDS, "world_8", "data_1.json"))), 2, 200, 2, {"block": 20}),
    ("c1_yaml4", scenario_c1().to_dict(), 10, 400, 3, {"full_reset_every": 150}),
    ("swarm12_nm3", swarm(12, 1.0, 0).to_dict(), 3, 250, 4, {"block": 20}),
    ("swarm12_nm10", swarm(12, 1.0, 1).to_dict(), 10, 250, 5, {"block": 20}),
    ("c3_n12_m3", scenario_c3(seed=1, n_drones=12, n_buildings=3).to_dict(), 10, 300, 6, {}),
    ("world_4_noreset", json.load(open(os.path.join(REF_WORLDS, "world_4", "data_1.json"))), 10, 200, 7, {"reset_on": ()}),
    # fp32-representable pre-states and actions (the reference is teacher-forced with float64 copies of fp32 values): the CUDA path has to match
    # these with no input-rounding slack
    ("f32_world_8", json.load(open(os.path.join(REF_WORLDS, "world_8", "data_1.json"))), 10, 200, 11, {"f32": True}),
    ("f32_c1_yaml4", scenario_c1().to_dict(), 10, 200, 12, {"f32": True, "full_reset_every": 90}),
    ("f32_swarm12_nm10", swarm(12, 1.0, 2).to_dict(), 10, 160, 13, {"f32": True, "block": 20}),
    ("f32_swarm12_nm3", swarm(12, 1.0, 3).to_dict(), 3, 120, 14, {"f32": True, "block": 15}),
    ("f32_c3_n12_m3", scenario_c3(seed=2, n_drones=12, n_buildings=3).to_dict(), 10, 160, 15, {"f32": True}),
]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for name, scn, nm, steps, seed, kw in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        tr = lockstep.run_reference(scn, nm, steps, seed, **kw)
        tr["scenario_json"] = np.array(json.dumps(scn))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **tr)
        print("%-16s N=%-3d steps=%-4d rows=%-5d maxcount=%d done=%d finish=%d  %.0f KB" % (
            name, tr["done"].shape[1], steps, int(tr["count"].sum()), int(tr["count"].max()), int(tr["done"].sum()),
            int(tr["finish"].sum()), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

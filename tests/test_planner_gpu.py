"""-m gpu: scenarios produced by the offline Theta* planner (uaisa_env.world) step through the CUDA path in lock-step with the oracle, and a
steering controller that follows the planned legs reaches goals without hitting the cylinders the routes were planned around."""
import numpy as np
import pytest
import torch

import _util  # noqa: F401
from gpu_harness import lockstep
from uaisa_env.scenario import scenario_c3
from uaisa_env.vec_env import VecMdin, unpack_state

pytestmark = pytest.mark.gpu


def test_planned_scenarios_in_lockstep_with_the_oracle():
    scn = scenario_c3(seed=4, n_scenarios=2, n_drones=8, n_buildings=5, planned=True)
    assert scn.n_points.max() >= 3
    st = lockstep(scn, 24, 10, 90, seed=3, epoch_len=50, block=15)
    st.assert_parity()
    assert st.finishes > 0 and st.resets > 0


def test_a_single_drone_follows_its_planned_route_around_the_cylinders():
    scn = scenario_c3(seed=9, n_scenarios=1, n_drones=8, n_buildings=6, planned=True)
    # one drone at a time (the others parked far outside their goals would collide at the centre): count building collisions only
    env = VecMdin(scn, 1, neighbors_num=10); env.reset()
    rng = np.random.default_rng(0)
    finished = np.zeros(env.A, bool); hit_building = 0
    bl = scn.buildings[0, : scn.n_buildings[0]]
    for t in range(120):
        st = unpack_state(env.get_state())
        des = env.cal_des().cpu().numpy().astype(np.float64)
        act = _util.steering_actions(rng, des, st["vel"], st["yaw"], st["pitch"], speed=0.35, noise=0.0).astype(np.float32)
        env.step(torch.from_numpy(act).cuda())
        post = unpack_state(env.get_state())
        d = env.done.cpu().numpy().astype(bool)
        for a in np.nonzero(d)[0]:
            p = post["pos"][a]
            hit_building += int(any(np.hypot(p[0] - b[0], p[1] - b[1]) <= 0.2 + b[3] and p[2] <= b[2] for b in bl))
        finished |= env.finish.cpu().numpy().astype(bool)
        env.reset((env.done | env.finish).to(torch.uint8))
    assert hit_building == 0 and finished.sum() >= 2, (hit_building, finished)
    env.close()

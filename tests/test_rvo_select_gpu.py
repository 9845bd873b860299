"""-m gpu: K7, classic RVO velocity selection (uaisa_rvo_select) against the corrected CPU restatement oracle/uaisa_oracle.c:uo_rvo_select.
PARITY UNPINNED against the reference (reciprocal_vel_obs.cal_vel is dead code that raises on its first call); the
restatement documents each fix (F1-F9).  Bar: the chosen candidate bit-exact, the derived env controls to 1e-6."""
import numpy as np
import pytest
import torch

import _util  # noqa: F401
import oracle
from uaisa_env.scenario import ScenarioSet, scenario_c1, scenario_c3, scenario_c4
from uaisa_env.vec_env import VecMdin, unpack_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,E,steps,acceler", [("c1", 64, 60, 0.5), ("c3", 32, 60, 0.5), ("c3", 16, 40, 1.0), ("c4", 2, 12, 0.5), ("swarm", 24, 50, 0.5)])
def test_rvo_select_matches_the_restatement(name, E, steps, acceler):
    import test_gpu_parity as tg
    scn = {"c1": scenario_c1, "c3": lambda: scenario_c3(seed=1, n_scenarios=3), "c4": lambda: scenario_c4(seed=2, n_scenarios=1), "swarm": lambda: tg.swarm(12, 1.0, 3)}[name]()
    env = VecMdin(scn, E, neighbors_num=10); ob = oracle.OracleBatch(scn, E, nm=10)
    env.reset(); ob.reset()
    rng = np.random.default_rng(0)
    n_bad = n_inside = n_total = 0
    for t in range(steps):
        st = unpack_state(env.get_state())
        ob.set_state(st)
        vel, ctrl = env.rvo_select(2.0, acceler)
        ov, oc = ob.rvo_select(2.0, acceler)
        g = vel.cpu().numpy()
        n_bad += int((g != ov.astype(np.float32)).any(axis=1).sum()); n_total += env.A
        assert np.allclose(ctrl.cpu().numpy(), oc, rtol=0, atol=2e-6), t
        # a candidate that is not the nearest-to-des grid point means some cone was in the way
        des = ob.cal_des()
        lo = np.clip(st["vel"] - acceler, -2, 2)
        near = np.clip(lo + 0.5 * np.round((des - lo) / 0.5), lo, np.clip(st["vel"] + acceler, -2, 2))
        n_inside += int((np.abs(ov - near) > 1e-9).any(axis=1).sum())
        # fly the selection for a step (every third step a random action keeps the swarm mixed)
        act = np.round(oc, 2).astype(np.float32) if t % 3 else _util.training_like_actions(rng, st["vel"]).astype(np.float32)
        env.step(torch.from_numpy(act).cuda()); env.autoreset(max_ep_len=40)
    assert n_bad == 0, (n_bad, n_total)
    if name != "c1":
        assert n_inside > 0.01 * n_total, (n_inside, n_total)       # the cones really constrain the choice in the dense scenarios
    env.close()


def test_rvo_baseline_controller_flies_the_sparse_scenario():
    """policy "rvo" in the batched evaluation loop: the scripted baseline reaches the goals of configs[1]'s scenario without collisions."""
    from uaisa_ppo.post_train import post_train
    env = VecMdin(scenario_c1(), 32, neighbors_num=10)
    res = post_train(env, num_episodes=32, max_ep_len=150).policy_test("rvo", policy_name="classic_rvo")
    assert len(res["episodes"]) == 32 and res["success_rate"] >= 0.9, res["line"]
    with pytest.raises(ValueError):
        post_train(env, num_episodes=1).policy_test("orca")
    env.close()

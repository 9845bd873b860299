"""Small end-to-end exercise of every kernel mode for compute-sanitizer (memcheck / racecheck), no oracle needed."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")]
from uaisa_env.scenario import scenario_c1, scenario_c3, scenario_world3
from uaisa_env.vec_env import VecMdin
rng = np.random.default_rng(0)
for scn, E in ((scenario_c1(), 40), (scenario_world3(), 45), (scenario_c3(seed=0, n_drones=12, n_buildings=3), 9)):
    env = VecMdin(scn, E, neighbors_num=10); env.reset()
    A = env.A
    for t in range(4):
        env.step(torch.from_numpy(np.round(rng.uniform(-1, 1, (A, 3)), 2).astype(np.float32)).cuda()); env.autoreset(max_ep_len=3)
    T = 4
    out = {"obs": torch.empty((T, A, env.obs_dim), device="cuda"), "count": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
           "nbr": torch.empty((T, A, 10), dtype=torch.int32, device="cuda"), "reward": torch.empty((T, A), device="cuda"),
           "done": torch.empty((T, A), dtype=torch.uint8, device="cuda"), "info": torch.empty((T, A), dtype=torch.uint8, device="cuda"),
           "finish": torch.empty((T, A), dtype=torch.uint8, device="cuda")}
    env.rollout(torch.from_numpy(np.round(rng.uniform(-1, 1, (T, A, 3)), 2).astype(np.float32)).cuda(), out, auto_reset=True, max_ep_len=3)
    m = torch.zeros(env.E, dtype=torch.uint8, device="cuda"); m[::3] = 1
    env.observe(m); env.cal_des(); env.indicators()
    r = env.step_host_sparse(np.round(rng.uniform(-1, 1, (A, 3)), 2).astype(np.float32), auto_reset=True)
    torch.cuda.synchronize()
    print("ok", scn.N, E, float(out["reward"].nan_to_num(0, 0, 0).sum()), len(r["rows"]))
    env.close()

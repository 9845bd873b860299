"""Debug aid: print the observation entries where the CUDA step and the oracle differ on the collinear-lattice states of tests/test_quirks_gpu.py."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import oracle
from uaisa_env.vec_env import VecMdin, pack_state, unpack_state
import test_quirks_gpu as tq

E = 4096
rng = np.random.default_rng(0)
scn = tq._pairs_scenario(E)
st = tq._base_state(E, 2)
base = rng.integers(-3, 4, (E, 3)).astype(np.float64)
base[np.all(base == 0, axis=1)] = [1, 2, 2]
rel = base * rng.choice([0.25, 0.5, 0.75, 1.0], (E, 1))
pa = 50.0 + rng.integers(0, 40, (E, 3)) * 0.25
st["pos"][0::2] = pa; st["pos"][1::2] = pa + rel
st["vel"][0::2] = rel * rng.choice([0.125, 0.25, 0.5], (E, 1)); st["vel"][1::2] = rel * rng.choice([-0.25, 0.125, 0.25], (E, 1))
act = np.zeros((2 * E, 3))
act[0::2] = rel * rng.choice([0.25, 0.5, 1.0], (E, 1)); act[1::2] = -rel * rng.choice([0.25, 0.5], (E, 1))
env = VecMdin(scn, E, neighbors_num=10); ob = oracle.OracleBatch(scn, E, nm=10)
env.reset(); ob.reset(); env.set_state(pack_state(st)); ob.set_state(unpack_state(env.get_state()))
env.step(torch.from_numpy(act.astype(np.float32)).cuda())
oobs, orew, *_ = ob.step(act.astype(np.float32).astype(np.float64))
g = env.obs.cpu().numpy(); o32 = oobs.astype(np.float32)
bad = np.argwhere(g != o32)
np.set_printoptions(precision=17, linewidth=250)
for a, c in bad[:20]:
    print("agent", a, "col", c, "gpu", repr(float(g[a, c])), "oracle", repr(float(oobs[a, c])), "count", int(ob.count[a]))
    e = a // 2
    print("  pre pos", st["pos"][2 * e], st["pos"][2 * e + 1], "vel", st["vel"][2 * e], st["vel"][2 * e + 1], "act", act[2 * e], act[2 * e + 1])
    post = ob.get_state()
    print("  post pos", post["pos"][2 * e], post["pos"][2 * e + 1], "vel", post["vel"][2 * e], post["vel"][2 * e + 1])
    print("  oracle row", oobs[a, 12:21], "\n  gpu row   ", g[a, 12:21].astype(np.float64))

"""CPU, world_size 2 over gloo: the env-sharding host logic.  Each rank steps ITS env range (with the
oracle standing in for the device), results are all-gathered and must equal the unsharded run bit for bit
(sharded result == single result for the same env ids)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _util  # noqa: F401
from uaisa_env.parallel import env_shard


def test_env_shard_partitions():
    for E, W in ((10, 3), (65536, 8), (7, 8), (1024, 1)):
        spans = [env_shard(E, r, W) for r in range(W)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == E
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        env_shard(4, 4, 4)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, E, T, out_q):
    import _util  # noqa: F401
    import oracle
    from uaisa_env.scenario import scenario_c3

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scn = scenario_c3(seed=5, n_scenarios=3, n_drones=6, n_buildings=2)
    e2s = (np.arange(E) % 3).astype(np.int32)
    start, cnt = env_shard(E, rank, world)
    oracle.set_threads(1)
    ob = oracle.OracleBatch(scn, cnt, env2scn=e2s[start:start + cnt], nm=10)
    ob.reset()
    acts = np.round(np.random.default_rng(7).uniform(-1, 1, (T, E, scn.N, 3)), 2)     # identical on every rank
    rew = np.zeros((T, cnt * scn.N)); done = np.zeros((T, cnt * scn.N), np.uint8)
    for t in range(T):
        _, r, d, _, f = ob.step(acts[t, start:start + cnt].reshape(-1, 3))
        rew[t], done[t] = r, d
        ob.reset((d | f).astype(np.uint8))
    # variable-size all-gather: pad to the largest shard
    mx = torch.tensor([cnt]); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    pad = torch.zeros((T, int(mx) * scn.N), dtype=torch.float64); pad[:, : cnt * scn.N] = torch.from_numpy(np.nan_to_num(rew, nan=7.0, posinf=8.0, neginf=9.0))
    got = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(got, pad)
    steps = torch.tensor([float(done.sum())]); dist.all_reduce(steps)
    if rank == 0:
        full = torch.cat([g[:, : env_shard(E, r, world)[1] * scn.N] for r, g in enumerate(got)], dim=1).numpy()
        ob2 = oracle.OracleBatch(scn, E, env2scn=e2s, nm=10); ob2.reset()
        ref = np.zeros((T, E * scn.N)); nd = 0
        for t in range(T):
            _, r, d, _, f = ob2.step(acts[t].reshape(-1, 3))
            ref[t] = np.nan_to_num(r, nan=7.0, posinf=8.0, neginf=9.0); nd += int(d.sum())
            ob2.reset((d | f).astype(np.uint8))
        out_q.put((bool(np.array_equal(full, ref)), float(steps.item()) == float(nd)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, 25, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, counted = q.get()
    assert same and counted

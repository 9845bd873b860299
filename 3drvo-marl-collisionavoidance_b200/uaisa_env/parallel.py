"""Multi-GPU sharding of environments.

Environments are independent -- all coupling is between the drones INSIDE one env
(rvo_inter.preprocess only walks that env's drone_state_list, rvo_inter.py:85-97) -- so ranks own
contiguous env ranges and the env step needs no data-path collective.  One process per GPU
(torch.distributed is only used for rendezvous, barriers and reducing timings / statistics).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def env_shard(n_envs_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition: (first env id, number of envs) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_envs_total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def dist_env():
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def make_sharded_env(scenarios, n_envs_total: int, env2scenario: Optional[np.ndarray] = None, rank: int = 0, world: int = 1,
                     device=None, **kw):
    """VecMdin over this rank's env range.  `env2scenario` is the GLOBAL [n_envs_total] map."""
    from .vec_env import VecMdin

    start, count = env_shard(n_envs_total, rank, world)
    if env2scenario is None:
        env2scenario = np.arange(n_envs_total, dtype=np.int32) % scenarios.S
    local = np.ascontiguousarray(env2scenario[start:start + count], np.int32)
    env = VecMdin(scenarios, count, env2scenario=local, device=device or "cuda", **kw)
    env.env_offset = start
    return env

"""-m gpu, world_size 2: the PRODUCT's multi-GPU paths.  Two processes -- one per GPU over NCCL when the box has two
GPUs, otherwise both on cuda:0 with gloo carrying the collectives -- run (1) ``make_sharded_env``: each rank steps its env range on the device,
the gathered outputs must equal the unsharded VecMdin bit for bit; (2) ``PPOTrainer``'s data-parallel update: after an update on the two halves of
a fixed buffer every rank holds IDENTICAL parameters, equal (to summation order) to the single-process update on the whole buffer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import _util  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _buffers(T, A, seed=3):
    rng = np.random.default_rng(seed)
    n = T * A
    obs = np.zeros((n, 102), np.float32)
    obs[:, :12] = np.round(rng.normal(0, 3, (n, 12)), 2)
    cnt = (rng.uniform(0, 1, n) < 0.4).astype(np.uint8)
    obs[:, 12:21] = np.round(rng.normal(0, 2, (n, 9)), 2) * cnt[:, None]
    act = np.round(rng.uniform(-1, 1, (n, 3)), 2).astype(np.float32)
    return (obs.reshape(T, A, 102), cnt.reshape(T, A), act.reshape(T, A, 3), rng.normal(-4, 0.5, (T, A)).astype(np.float32),
            rng.normal(0, 2, (T, A)).astype(np.float32), rng.normal(-3, 5, (T, A)).astype(np.float32))


def _run_update(tr, cols, bufs):
    obs, cnt, act, logp, adv, ret = bufs
    dev = tr.dev
    tr.buf["obs"].copy_(torch.from_numpy(obs[:, cols]).to(dev)); tr.buf["count"].copy_(torch.from_numpy(cnt[:, cols]).to(dev))
    tr.buf["act"].copy_(torch.from_numpy(act[:, cols]).to(dev)); tr.buf["logp"].copy_(torch.from_numpy(logp[:, cols]).to(dev))
    st = tr.update(torch.from_numpy(adv[:, cols].copy()).to(dev), torch.from_numpy(ret[:, cols].copy()).to(dev))
    return st


def _worker(rank, world, port, backend, q):
    import torch.distributed as dist
    import _util  # noqa: F401
    from uaisa_env.parallel import env_shard, make_sharded_env
    from uaisa_env.scenario import scenario_c1, scenario_c3
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.policy import rnn_ac
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": dev} if backend == "nccl" else {}))
    ok = {}

    def gather(t):      # gloo gathers host tensors only; NCCL device tensors only
        t = t.cpu() if backend == "gloo" else t
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        return [x.to(dev) for x in lst]
    # ---- (1) env sharding through the product API
    scn = scenario_c3(seed=5, n_scenarios=3, n_drones=6, n_buildings=2)
    E, T = 22, 20
    e2s = (np.arange(E) % 3).astype(np.int32)
    acts = np.round(np.random.default_rng(7).uniform(-1, 1, (T, E, scn.N, 3)), 2).astype(np.float32)
    env = make_sharded_env(scn, E, env2scenario=e2s, rank=rank, world=world, device=dev, neighbors_num=10)
    start, cnt = env_shard(E, rank, world)
    env.reset()
    mx = max(env_shard(E, r, world)[1] for r in range(world))
    rew = torch.full((T, mx * scn.N), 5.0, device=dev); obs = torch.zeros((T, mx * scn.N, env.obs_dim), device=dev)
    for t in range(T):
        env.step(torch.from_numpy(acts[t, start:start + cnt].reshape(-1, 3).copy()).to(dev)); env.autoreset(max_ep_len=9)
        rew[t, : cnt * scn.N] = torch.nan_to_num(env.reward, 7.0, 8.0, 9.0); obs[t, : cnt * scn.N] = env.obs
    g_r, g_o = gather(rew), gather(obs)
    if rank == 0:
        full = VecMdin(scn, E, env2scenario=e2s, neighbors_num=10, device=dev); full.reset()
        same = True
        for t in range(T):
            full.step(torch.from_numpy(acts[t].reshape(-1, 3)).to(dev)); full.autoreset(max_ep_len=9)
            cr = torch.cat([g[t, : env_shard(E, r, world)[1] * scn.N] for r, g in enumerate(g_r)]); co = torch.cat([g[t, : env_shard(E, r, world)[1] * scn.N] for r, g in enumerate(g_o)])
            same = same and torch.equal(cr, torch.nan_to_num(full.reward, 7.0, 8.0, 9.0)) and torch.equal(co, full.obs)
        ok["sharding"] = bool(same)
        full.close()
    env.close()
    # ---- (2) data-parallel PPO update: ranks own the two halves of the agent columns of one fixed buffer
    Tb, Ab = 25, 8
    bufs = _buffers(Tb, Ab)
    torch.manual_seed(1234); ac = rnn_ac()
    e1 = VecMdin(scenario_c1(), 1, device=dev)
    cfg = PPOConfig(steps_per_epoch=Tb, pi_lr=1e-3, vf_lr=1e-3, train_pi_iters=3, train_v_iters=3, target_kl=0.5, minibatch=10 ** 6, use_fused_policy=False,
                    use_cuda_graph=False, graph_update=(backend == "nccl"))
    tr = PPOTrainer(e1, ac=ac, cfg=cfg)
    assert tr.world == world
    st = _run_update(tr, slice(4 * rank, 4 * rank + 4), bufs)
    flat = tr.flat.p.detach().clone()
    got = gather(flat)
    ok["ranks_identical"] = bool(all(torch.equal(got[0], g) for g in got)) if rank == 0 else True
    if rank == 0:
        ok["pi_iters"] = st["pi_iters"]; ok["update_path"] = tr.update_path
        q.put((ok, flat.cpu().numpy()))
    dist.barrier()
    tr.close()                  # captured NCCL graphs go before the process group
    del tr
    e1.close()
    dist.destroy_process_group()


def test_two_rank_product_sharding_and_ppo_update():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        ok, flat2 = q.get(timeout=240)          # (read BEFORE joining: a 2.7 MB payload does not fit the pipe buffer, the writer would block)
    finally:
        for p in procs:
            p.join(60)
            if p.exitcode is None:
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ok["sharding"] and ok["ranks_identical"] and ok["pi_iters"] == 3, ok
    # the same update in ONE process on the whole buffer (full batch: the mean gradient is the same whatever the split)
    from uaisa_env.scenario import scenario_c1
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.policy import rnn_ac
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer
    torch.manual_seed(1234); ac = rnn_ac()
    e2 = VecMdin(scenario_c1(), 2)
    tr = PPOTrainer(e2, ac=ac, cfg=PPOConfig(steps_per_epoch=25, pi_lr=1e-3, vf_lr=1e-3, train_pi_iters=3, train_v_iters=3, target_kl=0.5, minibatch=10 ** 6,
                                             use_fused_policy=False, use_cuda_graph=False, graph_update=False))
    _run_update(tr, slice(0, 8), _buffers(25, 8))
    one = tr.flat.p.detach().cpu().numpy()
    d = np.abs(one - flat2)
    lr = 1e-3
    assert np.mean(d <= 0.02 * lr) >= 0.98 and d.max() <= 6.5 * lr, (np.mean(d <= 0.02 * lr), d.max(), ok)
    e2.close()

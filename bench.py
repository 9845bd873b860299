#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched 3-D RVO env step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5|sat]

A "step" is one env step (PASS A -> move -> PASS B -> trainer auto-reset + re-observe) of every env of
the workload; `value` = agent-steps/s over all ranks with actions and state resident in HBM (T-step rollout
launches, device-timed with CUDA events, L2 flushed before the timed region); `e2e` = the same metric through
the reference-facing host call uaisa_step_host_sparse: HOST action in, HOST results out (reward, flags, count,
12 self values per agent and every existing neighbour row) every step, wall-clock timed.
Default workload = BASELINE.json configs[1]: 1024 envs x 4 drones (train/drone_paths.yaml scenario).
Envs are independent, so ranks shard them with no data-path collective ("weak": E per GPU is fixed).

Extra objects in the JSON line: `roofline` (HBM: 556 algorithmic bytes per agent-step over the kernel's launch time, DRAM traffic
from the committed ncu capture), `roofline_flops` (SURVEY 8d's flop formula fed by the kernel's own pair counters, against the
nominal fp64 pipe), `saturating` (the same kernel at 262144 x 4 agents: the default workload is latency-bound by size),
`single_step_launch`, `cpu_baseline` (the C oracle with OpenMP on the host cores, bounded sample), `clocks`, `gpu_launches`.
`--impl reference` times that CPU port alone (rank 0 only); `--workload ppo` reports PPO samples/s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "3drvo-marl-collisionavoidance_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from uaisa_env.scenario import scenario_c1, scenario_c3, scenario_c4  # noqa: E402

NM = 10
MAX_EP_LEN = 300             # train_process.py default max_ep_len
BYTES_PER_AGENT_STEP = 156 + 40 * NM   # 556 B at nm=10


def workload(name):
    """-> (description, ScenarioSet, envs per GPU)"""
    if name == "c2":
        return "configs[1]: 1024 envs x 4 drones (train/drone_paths.yaml, 50x50x10, straight routes), nm=10", scenario_c1(), 1024
    if name == "c3":
        return "configs[2]: 4096 envs x 16 drones, 8 cylinders, 3-4 waypoint routes, nm=10", scenario_c3(seed=0, n_scenarios=8), 4096
    if name == "c4":
        return "configs[3]: 64 envs x 256 drones dense airspace 50x50x10, nm=10", scenario_c4(seed=0, n_scenarios=2), 64
    if name == "c5":
        return "configs[4] env part: 65536 envs x 4 drones per GPU, nm=10", scenario_c1(), 65536
    if name == "sat":
        return "saturating size: 262144 envs x 4 drones per GPU (configs[1] scenario), nm=10", scenario_c1(), 262144
    raise SystemExit("unknown workload %r" % name)


def synth_actions(rng, T, A):
    """i.i.d. U(-1,1)^3 rounded to 2 dp (the trainer rounds its actions to 2 dp, multi_ppo.py:197-205)."""
    return np.round(rng.uniform(-1.0, 1.0, size=(T, A, 3)), 2).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(scn, E, seconds=12.0, threads=0, seed=0):
    """The oracle port (oracle/uaisa_oracle.c, OpenMP over envs) timed on the host cores on a bounded
    sample of the same workload: same scenario, E envs, same action distribution, same reset policy."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    ob = oracle.OracleBatch(scn, E, nm=NM)
    try:
        avail = len(os.sched_getaffinity(0))      # torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = threads or avail
    oracle.set_threads(cores)
    rng = np.random.default_rng(seed)
    A = ob.A
    ep_len = np.zeros(A, np.int32)
    ob.reset(); ob.observe()
    chunk = 8
    acts = synth_actions(rng, chunk, A).astype(np.float64)

    def one(t):
        _, _, done, _, fin = ob.step(acts[t % chunk])
        _, eobs, _ = ob.autoreset(done, fin, ep_len, MAX_EP_LEN, False)
        if eobs.any():
            ob.observe(eobs)

    for t in range(2):
        one(t)
    n = 0
    t0 = time.perf_counter()
    while True:
        one(n); n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 100000:
            break
    return {"value": A * n / el, "unit": "agent-steps/s", "cores": int(cores), "kind": "port",
            "sample": "%d steps of %d envs x %d drones in %.1f s (oracle/uaisa_oracle.c, fp64, OpenMP over envs)" % (n, E, scn.N, el)}, n, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference is
    pure Python and /root/reference does not travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, scn, E = workload(args.workload)
    E_total = E * args.gpus
    # bounded sample: each "step" is one env step of the full E_total-env workload on the host
    E_run = min(E_total, 16384)
    budget = 20.0
    cb, n, el = cpu_baseline(scn, E_run, seconds=budget)
    K = max(args.steps, 1)
    line = {
        "impl": "reference", "metric": "agent-steps/sec (batched 3D RVO env step)", "value": cb["value"], "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * E_total * scn.N / cb["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "envs_total": E_total, "drones_per_env": scn.N, "nm": NM,
                   "sampled_envs": E_run, "sampled_steps": n, "note": "steps/warmup flags bound nothing here: the sample is time-bounded (%.0f s)" % budget},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    del K
    print(json.dumps(line))


def run_ppo(args):
    """--workload ppo: PPO samples/s = E*N*steps_per_epoch / (rollout + GAE + update [+ NCCL all-reduce]) per epoch
    (SURVEY.md section 8d).  configs[4]: 65536 envs x 4 drones over 8 GPUs = 8192 envs per GPU (weak scaling)."""
    import torch
    import torch.distributed as dist

    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
    scn = scenario_c1()
    E = args.ppo_envs
    env = VecMdin(scn, E, neighbors_num=NM, device=dev)
    cfg = PPOConfig(steps_per_epoch=args.ppo_steps)
    tr = PPOTrainer(env, cfg=cfg)
    epochs, warm = max(1, min(args.steps, 50)), max(1, min(args.warmup, 3))
    for _ in range(warm):
        tr.train_epoch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro = up = 0.0
    for _ in range(epochs):
        st = tr.train_epoch(); ro += st["rollout_s"]; up += st["update_s"]
    e1.record(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if clocks else None
    samples = world * env.A * cfg.steps_per_epoch * epochs
    if rank == 0:
        line = {"metric": "PPO samples/sec (rollout + GAE + update)", "value": samples / (float(ms.item()) * 1e-3), "unit": "samples/s", "n_gpus": world,
                "steps": epochs, "warmup": warm, "ms_per_step": float(ms.item()) / epochs, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if args.tf32 else "f32", "data": "synthetic",
                "config": {"workload": "configs[4]: %d envs x 4 drones per GPU, %d steps/epoch, rnn_ac 680991 params, reference PPO hyper-parameters" % (E, cfg.steps_per_epoch),
                           "minibatch": cfg.minibatch, "pi_iters": cfg.train_pi_iters, "v_iters": cfg.train_v_iters,
                           "policy": "torch (cuBLAS GEMMs) -- library path; fused forward is next", "rollout_frac": ro / max(ro + up, 1e-9)},
                "gpu_launches": int(env.launch_count - l0), "clocks": clk, "stats": {k: (float(v) if not isinstance(v, (int, float)) else v) for k, v in tr.stats.items()}}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--chunk", type=int, default=0, help="env steps per rollout launch (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sat", action="store_true")
    ap.add_argument("--no-autoreset", action="store_true", help="diagnostic: plain steps without the trainer reset policy")
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--slices", type=int, default=0, help="neighbour slices per drone (0 = auto)")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--ppo-envs", type=int, default=8192)
    ap.add_argument("--ppo-steps", type=int, default=300)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "ppo":
        return run_ppo(args)

    import torch
    import torch.distributed as dist

    from uaisa_env.vec_env import VecMdin

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc, scn, E = workload(args.workload)
    N = scn.N
    A = E * N
    K, W = max(args.steps, 1), max(args.warmup, 3)
    env = VecMdin(scn, E, neighbors_num=NM, device=dev, threads_per_block=args.threads_per_block, neighbor_slices=args.slices)
    env.reset()
    # ---- inputs resident in HBM: K steps of actions, outputs for one chunk of steps
    per_step_out = A * (env.obs_dim * 4 + 1 + NM * 4 + 4 + 3)
    chunk = args.chunk or max(1, min(K, int(6e9 // per_step_out), 512))
    rng = np.random.default_rng(1234 + rank)
    n_act = min(K, max(chunk, 64))
    acts = torch.from_numpy(synth_actions(rng, n_act, A)).to(dev)           # reused cyclically when K > n_act
    out = {"obs": torch.empty((chunk, A, env.obs_dim), dtype=torch.float32, device=dev), "count": torch.empty((chunk, A), dtype=torch.uint8, device=dev),
           "nbr": torch.empty((chunk, A, NM), dtype=torch.int32, device=dev), "reward": torch.empty((chunk, A), dtype=torch.float32, device=dev),
           "done": torch.empty((chunk, A), dtype=torch.uint8, device=dev), "info": torch.empty((chunk, A), dtype=torch.uint8, device=dev),
           "finish": torch.empty((chunk, A), dtype=torch.uint8, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def run_steps(nsteps, events=None):
        done_steps = 0
        while done_steps < nsteps:
            T = min(chunk, nsteps - done_steps, n_act)
            a0 = done_steps % n_act
            if a0 + T > n_act:
                T = n_act - a0
            o = {k: v[:T] for k, v in out.items()}
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            env.rollout(acts[a0:a0 + T], o, auto_reset=not args.no_autoreset, max_ep_len=MAX_EP_LEN)
            if events is not None:
                e1.record(); events.append((e0, e1, T))
            done_steps += T

    run_steps(W)
    torch.cuda.synchronize()
    flush.fill_(1)
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = env.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    run_steps(K, evs)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t_start.elapsed_time(t_end)
    launches = env.launch_count - l0
    tpc = env.cfg.threads_per_block
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in evs)
    clk = clocks.stop() if clocks else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * A * K / (ms_max * 1e-3)

    # ---- single-launch-per-step variant (what a closed-loop policy would see), device timed
    Ks = min(K, 200)
    o1 = {k: v[:1] for k, v in out.items()}
    for t in range(3):
        env.rollout(acts[t % n_act:t % n_act + 1], o1, auto_reset=True, max_ep_len=MAX_EP_LEN)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(Ks):
        env.rollout(acts[t % n_act:t % n_act + 1], o1, auto_reset=True, max_ep_len=MAX_EP_LEN)
    s1.record(); torch.cuda.synchronize()
    single_ms = s0.elapsed_time(s1) / Ks

    # ---- end to end through the host C-ABI call (uaisa_step_host_sparse): every step the action is produced on the
    # host and copied H2D, and the step's complete result (rewards, flags, the 12 self values of every observation and
    # every neighbour row that exists) is copied D2H and read on the host
    Ke = args.e2e_steps or min(K, 400 if A <= 65536 else 40)
    h_acts = synth_actions(rng, 8, A)
    for t in range(3):
        env.step_host_sparse(h_acts[t % 8], auto_reset=True, max_ep_len=MAX_EP_LEN)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rows_total = 0
    te0 = time.perf_counter()
    for t in range(Ke):
        r = env.step_host_sparse(h_acts[t % 8], auto_reset=True, max_ep_len=MAX_EP_LEN)
        _ = float(r["reward"][0])                   # the result is read on the host
        rows_total += len(r["rows"])
    e2e_s = time.perf_counter() - te0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * A * Ke / float(te.item())
    h2d = A * 3 * 4
    # result block written by the kernel: header (ready word + 2 timestamps), reward, self obs, flags + count, and the rows that exist
    d2h = int(24 + A * (4 + 48 + 4) + rows_total / max(Ke, 1) * 48)

    # ---- arithmetic roofline: algorithmic flops per agent-step from the kernel's own pair counters, measured on a
    # small copy of the workload (counting costs atomics, so it is not done in the timed region)
    flops = None
    try:
        cE = min(E, 1024)
        cenv = VecMdin(scn, cE, neighbors_num=NM, device=dev, count_pairs=True)
        cenv.reset()
        cT = 50
        cacts = torch.from_numpy(synth_actions(np.random.default_rng(123), cT, cenv.A)).to(dev)
        cenv.rollout(cacts, {}, auto_reset=True, max_ep_len=MAX_EP_LEN); cenv.counters()
        cenv.rollout(cacts, {"count": torch.empty((cT, cenv.A), dtype=torch.uint8, device=dev)}, auto_reset=True, max_ep_len=MAX_EP_LEN)
        cc = cenv.counters()
        rng_pp = cc["in_range_pairs"] / (cenv.A * cT); cone_pp = cc["cone_evals"] / (cenv.A * cT)
        M_ = int(scn.n_buildings.max()) if hasattr(scn, "n_buildings") else 0
        F = 150.0 + 2.0 * (N - 1) * 10.0 + cone_pp * 112.0 + M_ * 12.0      # F_self + range tests of both passes + full cone evaluations + cylinders
        fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12
        flops = {"bound": "fp64 pipe", "flops_per_agent_step": F, "in_range_pairs_per_agent_step": rng_pp, "cone_evaluations_per_agent_step": cone_pp,
                 "achieved": None, "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "nominal 148 SMs x 64 FMA/clk x 2 x 1.965 GHz (the step computes in fp64)",
                 "formula": "150 + 2 (N-1) 10 + cone_evals 112 + M 12 (SURVEY.md 8d; every add/mul/cmp/div/sqrt/transcendental counts 1)"}
        cenv.close()
    except Exception as ex:      # counters are diagnostics: never fail the bench line on them
        flops = {"error": str(ex)}

    # ---- side measurement at the saturating size (the default workload is 4096 agents = 128 warps on 148 SMs, i.e.
    # latency-bound by construction; SURVEY.md section 8d asks for the exact BASELINE config AND the saturating size)
    sat = None
    if args.workload == "c2" and not args.no_sat:
        env.close()
        del out, acts
        torch.cuda.empty_cache()
        sdesc, sscn, sE = workload("sat")
        senv = VecMdin(sscn, sE, neighbors_num=NM, device=dev)
        senv.reset()
        sA, sT, sK = senv.A, 10, 40
        sacts = torch.from_numpy(synth_actions(rng, sT, sA)).to(dev)
        sout = {"obs": torch.empty((sT, sA, senv.obs_dim), dtype=torch.float32, device=dev), "count": torch.empty((sT, sA), dtype=torch.uint8, device=dev),
                "nbr": torch.empty((sT, sA, NM), dtype=torch.int32, device=dev), "reward": torch.empty((sT, sA), dtype=torch.float32, device=dev),
                "done": torch.empty((sT, sA), dtype=torch.uint8, device=dev), "info": torch.empty((sT, sA), dtype=torch.uint8, device=dev),
                "finish": torch.empty((sT, sA), dtype=torch.uint8, device=dev)}
        senv.rollout(sacts, sout, auto_reset=True, max_ep_len=MAX_EP_LEN)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(sK // sT):
            senv.rollout(sacts, sout, auto_reset=True, max_ep_len=MAX_EP_LEN)
        q1.record(); torch.cuda.synchronize()
        sms = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        sval = world * sA * sK / (float(sms.item()) * 1e-3)
        sat = {"workload": sdesc, "value": sval, "unit": "agent-steps/s", "ms_per_step": float(sms.item()) / sK, "steps": sK, "steps_per_launch": sT,
               "per_gpu_hbm_gbs": sval / world * BYTES_PER_AGENT_STEP / 1e9}
        senv.close()
        env = None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        avg_launch_ms = kern_ms / max(len(evs), 1)
        steps_per_launch = K / max(len(evs), 1)
        achieved = BYTES_PER_AGENT_STEP * A * steps_per_launch / (avg_launch_ms * 1e-3) / 1e9
        traffic = None
        try:   # dram__bytes_read+write of one captured launch (profiles/traffic.json, from ncu --set full), scaled per agent-step
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["dram_bytes_per_agent_step"] * A * steps_per_launch
        except Exception:
            pass
        if sat is not None:
            sat["roofline_frac"] = sat["per_gpu_hbm_gbs"] / peak
        cb = None
        if not args.no_cpu:
            cb, _, _ = cpu_baseline(scn, min(E, 16384), seconds=12.0)
        line = {
            "metric": "agent-steps/sec (batched 3D RVO env step)", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "envs_per_gpu": E, "drones_per_env": N, "nm": NM, "auto_reset": "trainer policy (multi_ppo.py:221-281), max_ep_len=300",
                       "actions": "U(-1,1)^3 rounded to 2 dp, resident in HBM", "steps_per_launch": steps_per_launch, "threads_per_cta": tpc or "auto",
                       "l2": "flushed (256 MiB write) before the timed region; per-launch outputs %.0f MB" % (per_step_out * steps_per_launch / 1e6),
                       "state": "fp32 SoA in HBM, fp64 arithmetic"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "uaisa_env_kernel", "bytes_per_agent_step": BYTES_PER_AGENT_STEP, "avg_launch_ms": avg_launch_ms,
                         "peak_source": peak_src},
            "single_step_launch": {"value": world * A / (single_ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": single_ms,
                                   "note": "one launch per env step (closed-loop use), device-timed over %d steps" % Ks},
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "api": "uaisa_step_host_sparse: one CUDA graph per step with a single env-kernel node that writes the result block to mapped pinned host memory "
                           "(PCIe traffic inside the kernel, host polls a ready word); the action is read from mapped pinned memory (<= 1 MiB) or copied H2D from the caller's buffer; "
                           "result = reward/done/info/finish/count + 12 self values per agent + every existing neighbour row",
                    "neighbour_rows_per_step": rows_total / max(Ke, 1)},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if flops and "flops_per_agent_step" in flops:
            flops["achieved"] = value / world * flops["flops_per_agent_step"] / 1e12
            flops["frac"] = flops["achieved"] / flops["peak"]
        line["roofline_flops"] = flops
        if cb:
            line["cpu_baseline"] = cb
        if sat is not None:
            line["saturating"] = sat
        print(json.dumps(line))
    if env is not None:
        env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

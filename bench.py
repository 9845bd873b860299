#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched 3-D RVO env step and PPO samples/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5|sat|ppo]

A "step" is one env step (PASS A -> move -> PASS B -> trainer auto-reset + re-observe) of every env of the workload.

`value` = agent-steps/s over all ranks with actions and state resident in HBM.  One *repetition* is the K steps the caller
asked for (`--steps`), issued as uaisa_rollout launches of `steps_per_launch` steps; the timed region holds `config.repeats`
= R repetitions, R chosen so that the region lasts >= 100 ms whatever K is.  The repetitions are captured in a CUDA graph
(kernel nodes back to back) and the graph replays are enqueued behind a device-side gate (a spin kernel queued BEFORE the
start event), so no host launch latency sits between the two CUDA events.  Outputs rotate through buffers whose footprint
exceeds the 126 MB L2, and L2 is flushed before the region.  Time = max over ranks of the device-timed region.

`e2e` = the same metric through the reference-facing host call uaisa_step_host_sparse: HOST action in, HOST results out
(reward, flags, count, 12 self values per agent and every existing neighbour row) every step, wall-clock timed.

Default workload = BASELINE.json configs[1] (1024 envs x 4 drones per GPU, train/drone_paths.yaml scenario).  Envs are
independent, so ranks shard them with no data-path collective ("weak": E per GPU is fixed).  The default line also carries
time-bounded sub-measurements under `workloads`: configs[2] (4096x16 + cylinders), configs[3] (64x256), configs[4]'s env part
(65536x4 per GPU), the saturating size (262144x4) and PPO samples/s on configs[4] (65536 envs TOTAL over the run's ranks:
rollout + GAE + update with the NCCL gradient all-reduce) -- so the driver's BENCH/SCALE records hold them at every N.

Other objects in the JSON line: `roofline` (HBM: 556 algorithmic bytes per agent-step over the kernel's launch time; DRAM traffic
from the committed ncu capture), `roofline_flops` (SURVEY 8d's flop formula fed by the kernel's own pair counters, against the
fp64 pipe peak measured by profiles/fma_peak), `cpu_baseline` (the C oracle with OpenMP on the host cores, bounded sample, plus the
unmodified Python reference as measured in the build container), `clocks`, `gpu_launches`.
`--impl reference` times the CPU port alone (rank 0 only); `--workload ppo` prints the PPO line alone.
"""
from __future__ import annotations

import argparse
import json
import math
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
PPO_ENVS_TOTAL = 65536       # BASELINE.json configs[4]
MIN_REGION_MS = 120.0


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


def base_config(desc, E, N):
    """The part of `config` both arms (this one and --impl reference) print identically."""
    return {"workload": desc, "envs_per_gpu": E, "drones_per_env": N, "nm": NM,
            "auto_reset": "trainer policy (multi_ppo.py:221-281), max_ep_len=300",
            "actions": "U(-1,1)^3 rounded to 2 dp"}


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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
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
        time.sleep(0.12)
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


def container_reference():
    """The UNMODIFIED Python reference, timed in the build container by oracle/time_reference.py (it cannot travel to the GPU box)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2", "container_reference.json")))
    except Exception:
        return None


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
    cb = {"value": A * n / el, "unit": "agent-steps/s", "cores": int(cores), "kind": "port",
          "sample": "%d steps of %d envs x %d drones in %.1f s (oracle/uaisa_oracle.c, fp64, OpenMP over envs)" % (n, E, scn.N, el)}
    cr = container_reference()
    if cr and "env" in cr:
        cb["python_reference_build_container"] = {
            "note": "the UNMODIFIED reference env (uaisa_env/drone_envs/mdin.py:19-30) through oracle/ref_shim.py, P processes x 1 env, measured in the "
                    "build container (it cannot travel to the GPU box): oracle/time_reference.py", "cpu": cr.get("cpu"), "cores": cr.get("cores"), "env": cr["env"]}
    return cb, n, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference is
    pure Python and /root/reference does not travel to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload if args.workload != "ppo" else "c2"
    desc, scn, E = workload(wl)
    E_total = E * args.gpus
    # bounded sample: each "step" is one env step of (a sample of) the full E_total-env workload on the host
    E_run = min(E_total, 16384)
    budget = 20.0
    cb, n, el = cpu_baseline(scn, E_run, seconds=budget)
    cfg = base_config(desc, E, scn.N)
    cfg.update({"repeats": 1, "l2": "n/a (CPU arm)", "state": "fp64 (reference arithmetic)", "steps_per_launch": None})
    line = {
        "impl": "reference", "metric": "agent-steps/sec (batched 3D RVO env step)", "value": cb["value"], "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * E_total * scn.N / cb["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "sample": {"envs_total": E_total, "sampled_envs": E_run, "sampled_steps": n, "note": "steps/warmup flags bound nothing here: the sample is time-bounded (%.0f s)" % budget},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    cr = container_reference()
    if cr and "ppo" in cr:
        line["ppo_python_reference_build_container"] = cr["ppo"]
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1")); self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        else:
            self.dist = None
        self.sm_hz = float(torch.cuda.get_device_properties(self.dev).clock_rate) * 1e3     # kHz -> Hz (max SM clock)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_min(self, x):
        """max and min over ranks of a python float"""
        t = self.torch.tensor([x, -x], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0].item()), -float(t[1].item())

    def gate(self, ms):
        """device-side gate: a spin kernel that keeps the stream busy for ~ms while the host enqueues what follows"""
        self.torch.cuda._sleep(int(ms * 1e-3 * self.sm_hz))

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure_env(D, name, K, W, args, min_ms=MIN_REGION_MS, e2e=True, counters=True, single=True, clocks=False):
    """Device-resident throughput (+ optional e2e, flop counters) of one env workload at this run's world size."""
    torch = D.torch
    from uaisa_env.vec_env import VecMdin

    desc, scn, E = workload(name)
    N = scn.N
    A = E * N
    dev = D.dev
    env = VecMdin(scn, E, neighbors_num=NM, device=dev, threads_per_block=args.threads_per_block, neighbor_slices=args.slices)
    env.reset()
    auto = not args.no_autoreset
    per_step_out = A * (env.obs_dim * 4 + 1 + NM * 4 + 4 + 3)
    T = args.chunk or max(1, min(K, int(12e9 // per_step_out)))          # env steps per launch (one launch's outputs <= 12 GB of the 180 GB)
    lpr = (K + T - 1) // T                                               # launches per repetition (K steps)
    nbuf = int(min(64, max(2 if per_step_out * T < 2e9 else 1, math.ceil(300e6 / (per_step_out * T)))))     # output buffers rotate: footprint > L2 (one buffer when a single launch already writes GBs)
    rng = np.random.default_rng(1234 + D.rank)
    acts = torch.from_numpy(synth_actions(rng, K, A)).to(dev)
    outs = [{"obs": torch.empty((T, A, env.obs_dim), dtype=torch.float32, device=dev), "count": torch.empty((T, A), dtype=torch.uint8, device=dev),
             "nbr": torch.empty((T, A, NM), dtype=torch.int32, device=dev), "reward": torch.empty((T, A), dtype=torch.float32, device=dev),
             "done": torch.empty((T, A), dtype=torch.uint8, device=dev), "info": torch.empty((T, A), dtype=torch.uint8, device=dev),
             "finish": torch.empty((T, A), dtype=torch.uint8, device=dev)} for _ in range(nbuf)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    ctr = [0]

    def repetition():
        """K env steps = lpr launches, each into the next output buffer"""
        s = 0
        while s < K:
            t = min(T, K - s)
            o = {k: v[:t] for k, v in outs[ctr[0] % nbuf].items()}
            env.rollout(acts[s:s + t], o, auto_reset=auto, max_ep_len=MAX_EP_LEN)
            ctr[0] += 1
            s += t

    # ---- warm-up (W untimed steps at least), then calibrate R on a short eager run
    wsteps = 0
    while wsteps < max(W, 3):
        repetition(); wsteps += K
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.gate(2.0)
    c0.record()
    for _ in range(3):
        repetition()
    c1.record(); torch.cuda.synchronize()
    rep_ms = max(c0.elapsed_time(c1) / 3.0, 1e-3)
    rep_ms, _ = D.max_min(rep_ms)
    R = int(min(20000, max(1, math.ceil(min_ms / rep_ms))))
    G = int(max(1, min(R, max(1, 192 // lpr))))                           # repetitions per captured graph (<= ~192 kernel nodes)
    replays = (R + G - 1) // G
    R = replays * G
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(G):
                repetition()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph.replay()                                                        # one untimed replay (graph upload)
    torch.cuda.synchronize()
    flush.fill_(1)
    sampler = ClockSampler(D.local) if (clocks and D.rank == 0) else None
    D.barrier()
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.gate(max(3.0, 0.03 * replays))                                      # the host enqueues every replay while the stream spins
    t_start.record()
    for _ in range(replays):
        graph.replay()
    t_end.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = t_start.elapsed_time(t_end)
    ms_max, ms_min = D.max_min(ms)
    clk = sampler.stop() if sampler else None
    n_launch = replays * G * lpr
    value = D.world * A * K * R / (ms_max * 1e-3)
    avg_launch_ms = ms / n_launch
    spl = K / lpr
    res = {"desc": desc, "scn": scn, "E": E, "N": N, "A": A, "value": value, "ms_per_step": ms_max / (K * R), "repeats": R, "steps_per_launch": spl,
           "launches": n_launch, "region_ms": ms_max, "region_ms_min_rank": ms_min, "avg_launch_ms": avg_launch_ms, "clocks": clk,
           "per_gpu_hbm_gbs": BYTES_PER_AGENT_STEP * A * spl / (avg_launch_ms * 1e-3) / 1e9, "threads_per_cta": env.cfg.threads_per_block or "auto",
           "l2": "flushed (256 MiB write) before the region; outputs rotate through %d buffers = %.0f MB (> 126 MB L2)" % (nbuf, nbuf * per_step_out * T / 1e6)}

    # ---- single launch per env step (what a closed-loop caller sees), device timed, launch latency included on purpose
    if single:
        Ks = 200
        o1 = {k: v[:1] for k, v in outs[0].items()}
        for t in range(3):
            env.rollout(acts[t % K:t % K + 1], o1, auto_reset=True, max_ep_len=MAX_EP_LEN)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for t in range(Ks):
            env.rollout(acts[t % K:t % K + 1], o1, auto_reset=True, max_ep_len=MAX_EP_LEN)
        s1.record(); torch.cuda.synchronize()
        single_ms = s0.elapsed_time(s1) / Ks
        res["single_step_launch"] = {"value": D.world * A / (single_ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": single_ms,
                                     "note": "one launch per env step (closed-loop use), device-timed over %d steps, host launch latency included" % Ks}

    # ---- end to end through the host C-ABI call (uaisa_step_host_sparse): every step the action comes from host memory and the step's
    # complete result (rewards, flags, the 12 self values of every observation and every neighbour row that exists) is read on the host
    if e2e:
        h_acts = synth_actions(rng, 8, A)
        tc0 = time.perf_counter()
        for t in range(4):
            env.step_host_sparse(h_acts[t % 8], auto_reset=True, max_ep_len=MAX_EP_LEN)
        per_call = (time.perf_counter() - tc0) / 4
        Ke = args.e2e_steps or int(min(2000, max(20, math.ceil(0.25 / max(per_call, 1e-6)))))
        D.barrier()
        torch.cuda.synchronize()
        rows_total = 0
        te0 = time.perf_counter()
        for t in range(Ke):
            r = env.step_host_sparse(h_acts[t % 8], auto_reset=True, max_ep_len=MAX_EP_LEN)
            _ = float(r["reward"][0])                   # the result is read on the host
            rows_total += len(r["rows"])
        e2e_s = time.perf_counter() - te0
        e2e_max, _ = D.max_min(e2e_s)
        # result block written by the kernel: header (ready word + 2 timestamps), reward, self obs, flags + count, and the rows that exist
        res["e2e"] = {"value": D.world * A * Ke / e2e_max, "unit": "agent-steps/s", "h2d_bytes_per_step": A * 3 * 4,
                      "d2h_bytes_per_step": int(24 + A * (4 + 48 + 4) + rows_total / max(Ke, 1) * 48), "steps": Ke,
                      "neighbour_rows_per_step": rows_total / max(Ke, 1)}

    # ---- arithmetic roofline: algorithmic flops per agent-step from the kernel's own pair counters, measured on a
    # small copy of the workload (counting costs atomics, so it is not done in the timed region)
    if counters:
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
            peak, src = fp64_peak()
            ach = value / D.world * F / 1e12
            res["roofline_flops"] = {"bound": "fp64 pipe", "flops_per_agent_step": F, "in_range_pairs_per_agent_step": rng_pp, "cone_evaluations_per_agent_step": cone_pp,
                                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": src,
                                     "formula": "150 + 2 (N-1) 10 + cone_evals 112 + M 12 (SURVEY.md 8d; every add/mul/cmp/div/sqrt/transcendental counts 1)"}
            cenv.close()
        except Exception as ex:      # counters are diagnostics: never fail the bench line on them
            res["roofline_flops"] = {"error": str(ex)}
    env.close()
    del outs, acts, flush
    torch.cuda.empty_cache()
    return res


def fp64_peak():
    """fp64 FMA peak of the B200 as measured by profiles/fma_peak (micro-kernel of independent DFMA chains); nominal otherwise."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r2", "fma_peak.json")))
        return float(j["fp64_tflops"]), "measured: profiles/fma_peak.py DFMA micro-kernel on B200 (profiles/r2/fma_peak.json)"
    except Exception:
        return 148 * 64 * 2 * 1.965e9 / 1e12, "nominal 148 SMs x 64 FMA/clk x 2 x 1.965 GHz (the step computes in fp64)"


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def roofline_obj(res):
    peak, src = hbm_peak()
    traffic = None
    try:   # dram__bytes_read+write of one captured launch (profiles/traffic.json, from ncu --set full), scaled per agent-step
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["dram_bytes_per_agent_step"] * res["A"] * res["steps_per_launch"]
    except Exception:
        pass
    return {"bound": "hbm", "achieved": res["per_gpu_hbm_gbs"], "peak": peak, "unit": "GB/s", "frac": res["per_gpu_hbm_gbs"] / peak, "traffic": traffic,
            "kernel": "uaisa_env_kernel", "bytes_per_agent_step": BYTES_PER_AGENT_STEP, "avg_launch_ms": res["avg_launch_ms"],
            "avg_launch_ms_from": "timed region / launches (graph kernel nodes run back to back behind a device-side gate)", "peak_source": src}


def sub_object(res):
    o = {"workload": res["desc"], "value": res["value"], "unit": "agent-steps/s", "ms_per_step": res["ms_per_step"], "repeats": res["repeats"],
         "steps_per_launch": res["steps_per_launch"], "region_ms": res["region_ms"], "roofline": roofline_obj(res)}
    for k in ("e2e", "roofline_flops"):
        if k in res:
            o[k] = res[k]
    return o


def measure_ppo(D, args, envs_per_gpu, epochs=1, warm=1):
    """PPO samples/s = E*N*steps_per_epoch / (rollout + GAE + update [+ NCCL all-reduce]) per epoch."""
    torch = D.torch
    from uaisa_env.vec_env import VecMdin
    from uaisa_ppo.trainer import PPOConfig, PPOTrainer

    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
    scn = scenario_c1()
    env = VecMdin(scn, envs_per_gpu, neighbors_num=NM, device=D.dev)
    cfg = PPOConfig(steps_per_epoch=args.ppo_steps, tc_mlp=not args.no_tc_mlp)
    tr = PPOTrainer(env, cfg=cfg)
    for _ in range(warm):
        tr.train_epoch()
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro = up = 0.0
    per_epoch = []
    for _ in range(epochs):
        st = tr.train_epoch(); ro += st["rollout_s"]; up += st["update_s"]
        per_epoch.append({k: round(float(st[k]), 4) for k in ("rollout_s", "update_s", "update_prepare_ms", "update_pi_ms", "update_v_ms") if k in st})
    e1.record(); torch.cuda.synchronize()
    D.barrier()
    ms_max, _ = D.max_min(e0.elapsed_time(e1))
    samples = D.world * env.A * cfg.steps_per_epoch * epochs
    n = env.A * cfg.steps_per_epoch
    out = {"metric": "PPO samples/sec (rollout + GAE + update)", "value": samples / (ms_max * 1e-3), "unit": "samples/s", "epochs_timed": epochs, "epochs_warmup": warm,
           "ms_per_epoch": ms_max / epochs, "rollout_s_per_epoch": ro / epochs, "update_s_per_epoch": up / epochs, "per_epoch": per_epoch, "dtype": "tf32" if (args.tf32 or cfg.tc_mlp) else "f32",
           "workload": "configs[4]: %d envs x 4 drones per GPU x %d GPUs, %d steps/epoch, rnn_ac 680991 params, reference PPO hyper-parameters" % (envs_per_gpu, D.world, cfg.steps_per_epoch),
           "minibatch": cfg.minibatch, "pi_iters": cfg.train_pi_iters, "v_iters": cfg.train_v_iters,
           "minibatch_per_rank": int(tr._u["mb"]),
           "passes_per_sample": {"pi": tr.stats.get("pi_iters", 0) * int(tr._u["mb"]) / n, "v": cfg.train_v_iters * int(tr._u["mb"]) / n,
                                 "note": "the reference makes 50+50 FULL-BATCH passes over each robot's 300 samples (multi_ppo.py:355-376); here each iteration is one "
                                         "minibatch drawn without replacement from the whole [T x A] buffer -- samples/s is NOT a like-for-like ratio to the reference's"},
           "policy": tr.policy_path, "update": tr.update_path,
           "gpu_launches_per_epoch": {"env (action transform, step + reset policy in one launch) + fused policy forward + sampling (uaisa_policy_sample)": 4 * cfg.steps_per_epoch, "gae": 1,
                                      "clip/Adam (prepare + apply)": 2 * (cfg.train_pi_iters + cfg.train_v_iters), "policy repack": 1,
                                      "update forward/backward (reader fwd + bwd, 6 x uaisa_mlp_gemm, out_bwd, colsum)": 10 * (cfg.train_pi_iters + cfg.train_v_iters) if cfg.tc_mlp else 2 * (cfg.train_pi_iters + cfg.train_v_iters),
                                      "note": "own kernels per epoch (graph replays included)" + ("" if cfg.tc_mlp else "; the update's MLP GEMMs are torch / cuBLAS")},
           "stats": {k: (float(v) if not isinstance(v, (int, float, str)) else v) for k, v in tr.stats.items()}}
    tr.close()
    env.close()
    del tr
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--chunk", type=int, default=0, help="env steps per rollout launch (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-subs", action="store_true", help="skip the sub-measurements (configs[2..4], saturating size, PPO) of the default line")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--no-autoreset", action="store_true", help="diagnostic: plain steps without the trainer reset policy")
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--slices", type=int, default=0, help="neighbour slices per drone (0 = auto)")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--no-tc-mlp", action="store_true", help="PPO update MLPs through torch autograd / cuBLAS fp32 instead of the tcgen05 kernels (A/B)")
    ap.add_argument("--ppo-envs", type=int, default=0, help="envs per GPU for PPO (0 = 65536 / world size)")
    ap.add_argument("--ppo-steps", type=int, default=300)
    ap.add_argument("--ppo-epochs", type=int, default=1)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    D = Dist()
    K, W = max(args.steps, 1), max(args.warmup, 3)
    ppo_envs = args.ppo_envs or max(1, PPO_ENVS_TOTAL // D.world)
    if args.workload == "ppo":
        p = measure_ppo(D, args, ppo_envs, epochs=max(1, args.ppo_epochs), warm=2)      # (the second epoch still captures the update graph)
        if D.rank == 0:
            line = {"metric": p["metric"], "value": p["value"], "unit": p["unit"], "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": p["ms_per_epoch"], "higher_is_better": True, "scaling": "strong" if not args.ppo_envs else "weak", "vs_baseline": None,
                    "dtype": p["dtype"], "data": "synthetic", "config": {"workload": p["workload"], "a_step_is": "one PPO epoch"}, "ppo": p}
            print(json.dumps(line))
        D.close()
        return

    main_res = measure_env(D, args.workload, K, W, args, clocks=True)
    subs = {}
    if args.workload == "c2" and not args.no_subs:
        for key, name in (("configs[2]", "c3"), ("configs[3]", "c4"), ("configs[4]_env", "c5"), ("saturating", "sat")):
            r = measure_env(D, name, K, W, args, min_ms=100.0, e2e=name in ("c3", "c4"), counters=name in ("c3", "c4"), single=False)
            subs[key] = sub_object(r)
        if not args.no_ppo:
            try:
                p = measure_ppo(D, args, ppo_envs, epochs=1, warm=2)
                p["scaling"] = "strong (65536 envs in total, split over the ranks)"
                subs["ppo"] = p
            except Exception as ex:      # keep the env line even if the PPO leg fails (reported, not hidden)
                subs["ppo"] = {"error": repr(ex)}

    if D.rank == 0:
        desc, scn, E, N = main_res["desc"], main_res["scn"], main_res["E"], main_res["N"]
        cb = None
        if not args.no_cpu:
            cb, _, _ = cpu_baseline(scn, min(E, 16384), seconds=12.0)
        cfg = base_config(desc, E, N)
        cfg.update({"repeats": main_res["repeats"], "l2": main_res["l2"], "state": "fp32 SoA in HBM, fp64 arithmetic", "steps_per_launch": main_res["steps_per_launch"],
                    "threads_per_cta": main_res["threads_per_cta"], "timing": "R repetitions of the K steps in CUDA graphs behind a device-side gate; region %.1f ms (min over ranks %.1f ms)" % (main_res["region_ms"], main_res["region_ms_min_rank"])})
        line = {
            "metric": "agent-steps/sec (batched 3D RVO env step)", "value": main_res["value"], "unit": "agent-steps/s", "n_gpus": D.world,
            "steps": K, "warmup": W, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline_obj(main_res),
            "e2e": dict(main_res.get("e2e", {}), api="uaisa_step_host_sparse: one CUDA graph per step with a single env-kernel node that writes the result block to mapped pinned host memory "
                        "(PCIe traffic inside the kernel, host polls a ready word); the action is read from mapped pinned memory (<= 1 MiB) or copied H2D from the caller's buffer; "
                        "result = reward/done/info/finish/count + 12 self values per agent + every existing neighbour row"),
            "gpu_launches": int(main_res["launches"]), "clocks": main_res["clocks"],
        }
        for k in ("single_step_launch", "roofline_flops"):
            if k in main_res:
                line[k] = main_res[k]
        if cb:
            line["cpu_baseline"] = cb
        if subs:
            line["workloads"] = subs
            if "saturating" in subs:       # kept under its round-1 name too
                s = subs["saturating"]
                line["saturating"] = {"workload": s["workload"], "value": s["value"], "unit": "agent-steps/s", "ms_per_step": s["ms_per_step"],
                                      "per_gpu_hbm_gbs": s["roofline"]["achieved"], "roofline_frac": s["roofline"]["frac"]}
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()

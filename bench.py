#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused quadcopter env step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c4|c2|k1|c3|c5|c1] [--impl reference]

One bench "step" = ONE launch of the fused kernel: `fuse` consecutive env steps over the rank's
shard of envs, state in registers, the full rollout record (next obs, action, reward, done)
written to HBM.  `value` = env-steps/s over all ranks with inputs resident in HBM; `e2e` = the
same metric through the reference-facing API (DroneVecEnv.step with HOST numpy buffers: H2D of
the actions and D2H of obs / reward / done inside the timed region).

Workloads (BASELINE.json configs):
  c4 (default)  configs[3] per-GPU shard: 8,388,608 envs per GPU (= 64M over 8 GPUs, weak
                scaling), DroneGymEnv spec (15-dim obs, curriculum target, auto-reset),
                in-kernel Philox random actions, fuse=32.
  c2            configs[1]: 4096 envs x 1000 steps, random actions streamed from HBM, one launch.
  k1            the SB3-style boundary: one env step per launch (dronecu_step), 8M envs.
  c3            configs[2]: 1,048,576 envs, fused K-step rollout with the PPO MLP policy forward in-kernel (tcgen05).
  c5            configs[4]: full PPO iteration (rollout + GAE + 10 epochs x 4 minibatches), NCCL gradient all-reduce for N > 1.
  c1            configs[0]: the reference's own shape (1 env, n_steps 2048, batch 64, 10 epochs) -- latency-bound.

--impl reference times the reference's CPU implementation of the same path (the numpy port in
oracle/ -- /root/reference is Python and cannot travel to the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS_DIM = 15
REC_BYTES = OBS_DIM * 4 + 16 + 4 + 1          # next_obs + action + reward + done per env-step
STATE_BYTES = 80                              # 5 quads per env, read once + written once per launch


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (pynvml) -- runs during the timed region
# ------------------------------------------------------------------------------------------------
class stdout_to_stderr:
    """NCCL (and torch's process group) print a version banner on STDOUT when the communicator is created; the contract is ONE
    JSON line on stdout, so file descriptor 1 points at stderr while the process group initialises and runs its first collective."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the numpy port of the reference env (oracle/) on the host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(n_envs, seed):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import drone_oracle as do
    np.seterr(all="ignore")
    _W["env"] = do.BatchedDroneOracle(n_envs, do.SINGLE, seed=seed, env_offset=seed * n_envs)
    _W["rng"] = np.random.default_rng(seed)
    _W["env"].reset()


def _cpu_worker_steps(k_inner):
    env, rng = _W["env"], _W["rng"]
    for _ in range(k_inner):
        a = rng.uniform(0, 7.3575, (env.n, 4)).astype(np.float32)
        env.step(a)
    return env.n * k_inner


class CpuBaseline:
    """`cores` worker processes, each stepping its own shard of the reference env port."""

    def __init__(self, cores, envs_per_core=16384):
        import multiprocessing as mp
        self.cores, self.envs_per_core = cores, envs_per_core
        ctx = mp.get_context("spawn")
        self.pools = [ctx.Pool(1, initializer=_cpu_worker_init, initargs=(envs_per_core, i)) for i in range(cores)]

    def step(self, k_inner):
        res = [p.apply_async(_cpu_worker_steps, (k_inner,)) for p in self.pools]
        return sum(r.get() for r in res)

    def close(self):
        for p in self.pools:
            p.terminate()


def time_cpu_baseline(cores, budget_s=12.0, k_inner=8):
    cb = CpuBaseline(cores)
    cb.step(2)                                   # warm-up (imports, first-touch)
    t0, units = time.perf_counter(), 0
    while True:
        units += cb.step(k_inner)
        dt = time.perf_counter() - t0
        if dt > budget_s:
            break
    cb.close()
    return units / dt, f"{cores} procs x {cb.envs_per_core} envs, DroneGymEnv spec + auto-reset, random actions, {units} env-steps in {dt:.1f}s"


def run_reference(args):
    """Reference arm: the reference's CPU path (numpy port, all host cores) on the same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload in ("c3", "c5", "c1"):
        return run_reference_ppo(args, cores)
    k_inner = 8
    cb = CpuBaseline(cores)
    for _ in range(max(1, args.warmup)):
        cb.step(k_inner)
    t0, units = time.perf_counter(), 0
    for _ in range(args.steps):
        units += cb.step(k_inner)
    dt = time.perf_counter() - t0
    cb.close()
    v = units / dt
    sample = f"{cores} procs x {cb.envs_per_core} envs x {k_inner} env-steps per bench step"
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.workload], "sample": sample},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


PPO_CPU_SAMPLE = dict(n_envs=4096, n_steps=32, n_epochs=10, minibatches=4)


def time_cpu_ppo(update, budget_s=10.0, reference_shape=False):
    """The restated-SB3 CPU loop (oracle/ppo_loop.py, torch CPU with all host threads + the numpy env port):
    (a) batched the way this repo's workload is, (b) exactly as the reference runs it (train.py:33-43:
    one env, n_steps 2048, batch 64, 10 epochs)."""
    import torch
    from oracle import ppo_loop
    if reference_shape:
        v1, its1, dt1 = ppo_loop.time_loop(1, 2048, 64, 10, budget_s=2 * budget_s, update=True)
        return v1, (f"restated SB3 loop on CPU (torch {torch.get_num_threads()} threads + numpy env port), the reference's own "
                    f"config: 1 env, n_steps 2048, batch 64, 10 epochs; {its1} iterations in {dt1:.1f}s")
    c = PPO_CPU_SAMPLE
    v, its, dt = ppo_loop.time_loop(c["n_envs"], c["n_steps"], c["n_envs"] * c["n_steps"] // c["minibatches"],
                                    c["n_epochs"], budget_s=budget_s, update=update)
    sample = (f"restated SB3 loop on CPU (torch {torch.get_num_threads()} threads + numpy env port): {c['n_envs']} envs x "
              f"{c['n_steps']} steps" + (f", {c['n_epochs']} epochs x {c['minibatches']} minibatches" if update else ", rollout only") +
              f"; {its} iterations in {dt:.1f}s")
    if update:
        v1, its1, dt1 = ppo_loop.time_loop(1, 2048, 64, 10, budget_s=budget_s, update=True)
        sample += f" | reference's own config (1 env, n_steps 2048, batch 64, 10 epochs): {v1:.0f} env-steps/s"
    return v, sample


def run_reference_ppo(args, cores):
    from oracle import ppo_loop
    c = PPO_CPU_SAMPLE if args.workload != "c1" else dict(n_envs=1, n_steps=2048, n_epochs=10, minibatches=32)
    update = args.workload != "c3"
    loop = ppo_loop.PPOLoopOracle(c["n_envs"], c["n_steps"], c["n_envs"] * c["n_steps"] // c["minibatches"], c["n_epochs"])

    def one():
        loop.collect_rollouts()
        if update:
            loop.train()
    for _ in range(max(1, args.warmup)):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    v = args.steps * c["n_envs"] * c["n_steps"] / dt
    sample = (f"restated SB3 loop on CPU (torch, all host threads + numpy env port): {c['n_envs']} envs x {c['n_steps']} steps per "
              f"bench step" + (f", {c['n_epochs']} epochs x {c['minibatches']} minibatches" if update else ", rollout only"))
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.workload], "sample": sample},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


WORKLOAD_NAMES = {
    "c4": "configs[3] shard: step-only fused rollout, 8388608 envs/GPU (64M over 8), DroneGymEnv spec, Philox actions",
    "c2": "configs[1]: vectorized step-only, 4096 envs x 1000 steps, random actions streamed from HBM",
    "k1": "SB3 boundary: one env step per launch (dronecu_step), 8388608 envs/GPU, streamed actions",
    "c3": "configs[2]: 1048576 envs fused K-step rollout with the PPO MLP policy/value forward in-kernel",
    "c5": "configs[4]: full PPO loop (in-kernel-policy rollout + GAE + 10-epoch minibatch update, NCCL grad all-reduce)",
    "c1": "configs[0]: the reference's own shape -- ONE env, n_steps 2048, batch 64, 10 epochs (train.py defaults), full PPO loop",
}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                           # creates the communicator (and prints NCCL's banner) now
            torch.cuda.synchronize()
    import drone_rl_b200 as drl

    wl = args.workload
    if wl == "c2":
        n, fuse = 4096, 1000
        cfg = drl.EnvConfig.vector()
    else:
        n, fuse = args.envs_per_gpu, (1 if wl == "k1" else args.fuse)
        cfg = drl.EnvConfig.single()
    D = cfg.obs_dim
    batch = drl.DroneBatch(n, cfg, device=local, seed=args.seed, env_offset=rank * n)

    # ---- device-resident buffers ----------------------------------------------------------------
    nxt = torch.empty(fuse, n, D, device=dev)
    rew = torch.empty(fuse, n, device=dev)
    done = torch.empty(fuse, n, dtype=torch.uint8, device=dev)
    if wl == "c4":
        acts_in, acts_out = None, torch.empty(fuse, n, 4, device=dev)
    else:
        acts_in = torch.rand(fuse, n, 4, device=dev) * 7.3575
        acts_out = None

    def one_step():
        if wl == "k1":
            batch.step(acts_in[0], out={"obs": nxt[0], "reward": rew[0], "done": done[0]})
        else:
            batch.rollout(fuse, acts_in, next_obs=nxt, out_actions=acts_out, reward=rew, done=done)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = batch.launch_count
    with ClockSampler(local) as clocks:
        ev[0].record()
        for i in range(args.steps):
            one_step()
            ev[i + 1].record()
        barrier()
    launches = batch.launch_count - l0
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]        # ms, per launch, on the launching stream
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    units = n * fuse * args.steps * world
    value = units / (total_ms * 1e-3)

    # ---- roofline of the dominant (only) kernel ----------------------------------------------------
    in_bytes = 16 if acts_in is not None else 0
    out_bytes = D * 4 + 4 + 1 + (16 if acts_out is not None else 0)
    alg_bytes_per_launch = n * (fuse * (in_bytes + out_bytes) + 2 * STATE_BYTES)
    avg_ms = float(np.mean(per))
    peak, peak_src = peaks()
    achieved = alg_bytes_per_launch / (avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "rollout_kernel",
                "algorithmic_bytes_per_env_step": in_bytes + out_bytes + 2 * STATE_BYTES / fuse,
                "avg_launch_ms": avg_ms, "min_launch_ms": float(np.min(per))}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(wl)
        except Exception:
            pass

    # ---- e2e through the reference-facing API with HOST buffers ---------------------------------------
    del nxt, rew, done, acts_in, acts_out
    torch.cuda.empty_cache()
    e2e = None if args.no_e2e else measure_e2e(drl, args, wl, n, rank, local, world, dist, barrier)
    batch.close()

    if rank == 0:
        line = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES[wl], "envs_per_gpu": n, "env_steps_per_launch": fuse,
                           "obs_dim": D, "outputs": "next_obs+action+reward+done" if wl == "c4" else "next_obs+reward+done",
                           "l2": "inputs larger than L2 (state %.0f MB, outputs %.1f GB per launch)" % (
                               n * STATE_BYTES / 1e6, n * fuse * out_bytes / 1e9) if wl != "c2" else
                           "L2 flushed by the 196 MB obs output of every launch"},
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary()}
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            v, sample = time_cpu_baseline(cores)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(drl, args, wl, n, rank, local, world, dist, barrier):
    """DroneVecEnv.step (SB3 VecEnv surface) with pinned host numpy buffers; every call copies the
    actions H2D and obs / reward / done D2H."""
    import torch
    cfg = drl.EnvConfig.vector() if wl == "c2" else drl.EnvConfig.single()
    env = drl.DroneVecEnv(n, seed=args.seed, device=local, env_offset=rank * n, info_mode="arrays", copy=False, config=cfg)
    D = cfg.obs_dim
    acts = torch.empty(n, 4, pin_memory=True).uniform_(0, 7.3575).numpy()
    env.reset()
    steps = max(3, min(args.steps, 10)) if wl != "c2" else 200
    for _ in range(3):
        env.step(acts)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, rew, done, _ = env.step(acts)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=torch.device("cuda", local), dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    checksum = float(rew[: min(n, 1024)].sum())
    env.close()
    return {"value": n * steps * world / dt, "unit": "env-steps/s", "h2d_bytes_per_step": n * 16,
            "d2h_bytes_per_step": n * (D * 4 + 4 + 1 + 1), "api": "DroneVecEnv.step (numpy, pinned)", "calls": steps,
            "ms_per_call": 1e3 * dt / steps, "reward_checksum": checksum}


def run_ppo(args):
    """c3 (rollout with the in-kernel MLP) and c5 (full PPO iteration) -- env-steps/s."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
    import drone_rl_b200 as drl
    from drone_rl_b200.ppo import PPO
    wl = args.workload
    n = args.ppo_envs
    K = args.fuse
    env = drl.DroneBatch(n, drl.EnvConfig.single(), device=local, seed=args.seed, env_offset=rank * n)
    model = PPO(env, n_steps=K, batch_size=n * K // args.ppo_minibatches, n_epochs=args.ppo_epochs, seed=args.seed,
                rollout_precision=args.precision, update_precision=args.update_precision or args.precision)

    def one_step():
        model.collect_rollouts()
        if wl in ("c5", "c1"):
            model.train()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    l0 = env.launch_count + model.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model.grad_events = [] if wl == "c5" else None        # c1 keeps the CUDA-graph path (no per-launch events inside a graph)
    roll_events = []
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            if wl == "c3":
                e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                e[0].record()
            one_step()
            if wl == "c3":
                e[1].record()
                roll_events.append(e)
        ev1.record()
        barrier()
    # the dominant kernel of the workload, timed live by CUDA events on the launching stream
    if wl == "c1":
        ms, units, kflop, kname = [ev0.elapsed_time(ev1) / args.steps], float(n * K), 20864.0 * (1 + 3 * args.ppo_epochs), "whole iteration (latency-bound: 2048 sequential env steps + one CUDA graph per epoch)"
    elif wl == "c5":
        ms = [a.elapsed_time(b) for a, b, _ in model.grad_events]
        units = float(np.mean([m for _, _, m in model.grad_events]))
        up = args.update_precision or args.precision
        kflop, kname = 3 * 20864.0, {"tf32": "ppo_grad_tc_kernel", "bf16": "ppo_grad_bf16_kernel", "fp32": "ppo_grad_kernel"}[up] + " (+ its fixed-order reduce)"
        model.grad_events = None
    else:
        ms = [a.elapsed_time(b) for a, b in roll_events]
        units = float(n * K)
        kflop, kname = 20864.0, ("policy_rollout_tc_kernel" if args.precision == "tf32" else "policy_rollout_kernel") + " (+ gae_kernel)"
    k_ms = float(np.mean(ms))
    k_tflops = units * kflop / (k_ms * 1e-3) / 1e12
    launches = env.launch_count + model.launches - l0
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n * K * args.steps * world / (total_ms * 1e-3)
    flop_per_step = 20864.0 * (1 if wl == "c3" else 1 + 3 * args.ppo_epochs)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(pk.get("bf16_tflops_sustained", 1400.0))
    achieved = value / world * flop_per_step / 1e12
    tensor = (args.precision == "tf32") and (wl == "c3" or (args.update_precision or args.precision) in ("tf32", "bf16"))

    # ---- e2e: the public API (PPO.collect_rollouts / PPO.learn), wall clock, every iteration reads its
    # result (episode statistics + train/* scalars) back to the host.  The simulator lives on the GPU, so an
    # iteration has no host input: h2d is 0 by construction.
    e2e = None
    if not args.no_e2e:
        its = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        if wl in ("c5", "c1"):
            model.num_timesteps = 0
            model.learn(total_timesteps=its * n * K * world)
            d2h = 9 * 4 + 128 * 64
        else:
            for _ in range(its):
                model.collect_rollouts()
                st = env.episode_stats(reset=True)
            d2h = 128 * 64
        barrier()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": its * n * K * world / float(td.item()), "unit": "env-steps/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": d2h, "api": "PPO.learn" if wl in ("c5", "c1") else "PPO.collect_rollouts + episode_stats",
               "iterations": its, "note": "GPU-resident simulator: no host inputs per iteration; the result read back is the "
               "episode statistics (128 slots x 64 B)" + (" and the train/* scalars" if wl == "c5" else "")}
    if rank == 0:
        line = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": ("f32" if args.precision == "fp32" else "tf32 (fp32 accumulate)") if wl == "c3" else
                         "rollout MLP %s, update %s (fp32 accumulate; env step, loss, Adam in f32)" % (args.precision, args.update_precision or args.precision),
                "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES[wl], "envs_per_gpu": n, "n_steps": K, "n_epochs": args.ppo_epochs,
                           "minibatches_per_epoch": args.ppo_minibatches, "policy": "MlpPolicy 15-64-64-{4,1} tanh, random init", "rollout_precision": args.precision,
                           "update_precision": args.update_precision or args.precision,
                           "l2": ("rollout buffers %.1f GB per iteration, larger than L2" % (n * K * 93 / 1e9)) if wl != "c1" else
                           "latency-bound workload (190 KB of buffers): no L2 effect to flush"},
                "roofline": {"bound": "tensor", "achieved": k_tflops, "peak": peak, "unit": "TFLOP/s", "frac": k_tflops / peak,
                             "traffic": None, "kernel": kname, "avg_launch_ms": k_ms, "units_per_launch": units,
                             "algorithmic_flop_per_unit": kflop, "whole_step_tflops": achieved, "note": ("tcgen05 kind::tf32 MMAs (tf32 dense peak = half the bf16 peak used as the "
                             "denominator)" if tensor else "fp32 CUDA-core parity path (nominal 74.4 TFLOP/s FMA peak: frac_of_fp32 = "
                             "%.3f); the denominator is the measured bf16 tensor peak" % (achieved / 74.4)),
                             "flop_per_env_step": flop_per_step},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
                "train": {k: v for k, v in model.logger_values.items() if k.startswith("train/")}}
        if world == 1 and not args.no_cpu:
            v, sample = time_cpu_ppo(update=(wl != "c3"), reference_shape=(wl == "c1"))
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # NCCL prints its version banner on STDOUT when NCCL_DEBUG=VERSION (some images export it): keep stdout = the JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(WORKLOAD_NAMES))
    ap.add_argument("--envs-per-gpu", type=int, default=8_388_608)
    ap.add_argument("--fuse", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (A/B kernel timing only)")
    ap.add_argument("--ppo-envs", type=int, default=1_048_576)
    ap.add_argument("--ppo-epochs", type=int, default=10)
    ap.add_argument("--ppo-minibatches", type=int, default=4)
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"], help="policy MLP in the rollout: CUDA cores or tcgen05")
    ap.add_argument("--update-precision", default="bf16", choices=["fp32", "tf32", "bf16"], help="PPO minibatch gradient: fp32 = CUDA cores (parity path), tf32 = tcgen05 all-tf32, bf16 = tcgen05 with bf16 weight-gradient operands, three tiles per SM (default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c3", "c5", "c1"):
        if args.workload == "c1":          # reference train.py:12-16, :36-43
            args.ppo_envs, args.fuse, args.ppo_minibatches, args.ppo_epochs = 1, 2048, 32, 10
        run_ppo(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched F-16 longitudinal simulator on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 2^20
independent aircraft PER GPU, fp32, random initial states (plane.random_state_bound) drawn
on the device, random `combo` pitch references, uniform random stick, gymnasium autoreset
semantics.  One bench "step" = one F16VecEnv.step() = ONE kernel launch that advances every
env by one env-step and emits obs / reward / done (K = 1, the closed-loop gym step).  Envs
are sharded over GPUs with no data-path collective (weak scaling).

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB write)
before each timed step and excluded from the timing, summed over the K steps, max over
ranks.  `e2e` repeats the measurement through the host-buffer entry point
(f16_env_step_host: pinned host actions in, host obs/reward/done out, copies inside the
timed region).  `cpu_baseline` / `--impl reference` time the CPU oracle (oracle/f16_oracle.c,
a C port of the reference's NumPy path -- the reference itself is Python and cannot travel
to the GPU box) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (F-16 long., batched)"
UNIT = "env-steps/s"
ENVS_PER_GPU = 1 << 20
BYTES_PER_ENV_STEP_F32 = 113      # 48 B read + 65 B written per env-step, fp32, K = 1
# (the kernel really moves 117 B: + total_return r/w 8, - the V row write 4)
CFG = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"      # B200_PROFILING.md fallback


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_run(n_envs, n_steps, threads=0, seed=0):
    """Time the CPU oracle on a bounded sample of the same workload: n_envs aircraft from the
    same initial-state distribution, uniform random stick, n_steps env-steps each."""
    import numpy as np

    from oracle import f16_oracle as orc
    import f16_model_b200 as f16

    rng = np.random.default_rng(seed)
    d5 = np.radians(5)
    al = rng.uniform(-d5, d5, n_envs)
    x0 = np.zeros((n_envs, 9))
    x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n_envs), rng.uniform(-d5, d5, n_envs), al
    x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n_envs), al
    acts = rng.uniform(-1, 1, (n_steps, n_envs))
    bank = f16.build_reference_bank(0.01, 10, "combo", False, bank_size=64, seed=seed)
    idx = rng.integers(0, bank.shape[0], n_envs).astype(np.int32)
    t0 = time.perf_counter()
    out = orc.rollout(x0, acts, bank, idx, want_states=False, want_obs=True, nthreads=threads)
    dt = time.perf_counter() - t0
    return out["env_steps"], dt, orc.num_threads() if threads == 0 else threads


def reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path -- here the C oracle port,
    all host threads.  Rank 0 only."""
    if rank != 0:
        return
    n_envs, n_steps = 1 << 16, 16
    for _ in range(args.warmup):
        cpu_oracle_run(n_envs, n_steps)
    total, tsum, cores = 0, 0.0, 1
    for s in range(args.steps):
        steps, dt, cores = cpu_oracle_run(n_envs, n_steps, seed=s)
        total += steps
        tsum += dt
    value = total / tsum
    sample = f"{n_envs} envs x {n_steps} env-steps per bench step, config-3 distributions, C oracle port (OpenMP)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tsum / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "F-16 longitudinal env-step, 2^20 envs/GPU, fp32, K=1 (reference arm: bounded CPU sample)",
                   "envs_per_gpu": ENVS_PER_GPU, "substeps_per_launch": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (SciPy interpolators rebuilt every step, ~156 env-steps/s/core measured in the "
                "build container, BASELINE.md section 2) and cannot travel to the GPU box; this arm times its C restatement",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--no-sweep", action="store_true", help="skip the K-substeps sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fast-math", type=int, default=1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import f16_model_b200 as f16
    from f16_model_b200.sharding import make_sharded_env

    f16.load_library()                       # fails loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = args.envs
    fast = bool(args.fast_math)
    env = make_sharded_env(CFG, N * world, rank, world, device, dtype=torch.float32, autoreset=True, seed=1,
                           fast_math=fast, bank_size=256)
    g = torch.Generator(device=device).manual_seed(1000 + rank)
    pool = torch.rand((16, N), device=device, generator=g) * 2 - 1          # 16 different stick tensors, cycled
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)          # > 126 MB L2

    static_actions = pool[0].clone()
    graph, _ = env.make_graphed_step(static_actions)                          # same launch, replayed from a CUDA graph

    def timed_steps(n_steps, do_flush=True, graphed=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for s in range(n_steps):
            if graphed:
                static_actions.copy_(pool[s % 16])
            if do_flush:
                flush.fill_(s & 0xFF)
            ev[s][0].record()
            if graphed:
                graph.replay()
            else:
                env.step(pool[s % 16])
            ev[s][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev)                          # ms on the device

    for _ in range(max(args.warmup, 3)):
        flush.fill_(0)
        env.step(pool[0])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = timed_steps(args.steps, do_flush=True)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    warm_ms = timed_steps(args.steps, do_flush=False)                         # L2-warm variant, reported separately
    barrier()
    graph_ms = max_over_ranks_later = timed_steps(args.steps, do_flush=True, graphed=True)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms = max_over_ranks(dev_ms)
    warm_ms = max_over_ranks(warm_ms)
    graph_ms = max_over_ranks(graph_ms)
    total_env_steps = N * world * args.steps
    value = total_env_steps / (dev_ms * 1e-3)

    # ---- end to end through the host-buffer entry point -------------------------------------
    e2e = None
    if not args.no_e2e:
        a_host = (torch.rand((4, N)) * 2 - 1).pin_memory()
        out = (torch.empty((N, 5)).pin_memory(), torch.empty((1, N)).pin_memory(),
               torch.empty((1, N), dtype=torch.uint8).pin_memory())
        for s in range(3):
            env.step_host(a_host[s % 4], out=out)
        barrier()
        n_e2e = max(10, min(args.steps, 50))
        t0 = time.perf_counter()
        for s in range(n_e2e):
            env.step_host(a_host[s % 4], out=out)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": N * world * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": 25 * N,
               "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e,
               "path": "f16_env_step_host: pinned host actions -> H2D -> kernel -> D2H obs/reward/done, 8 chunks on 3 streams"}

    # ---- K-substeps sweep (state in registers across K env-steps per launch) -----------------
    sweep = None
    if not args.no_sweep:
        sweep = {}
        for K in (1, 4, 16, 64):
            acts = torch.rand((K, N), device=device, generator=g) * 2 - 1
            outb = (env._obs, torch.empty((K, N), device=device), torch.empty((K, N), dtype=torch.uint8, device=device))
            for _ in range(3):
                env.step_k(acts, out=outb)
            reps = max(3, 64 // K)
            tot = 0.0
            for _ in range(reps):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                env.step_k(acts, out=outb)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            sweep[str(K)] = N * world * K * reps / (max_over_ranks(tot) * 1e-3)
            del acts, outb

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    per_launch_s = dev_ms * 1e-3 / args.steps
    achieved = BYTES_PER_ENV_STEP_F32 * N / per_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("env_step_f32_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_kind": f"of {peak_kind}", "kernel": "env_step_kernel<float,fast,autoreset>",
                "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP_F32, "units_per_launch": N,
                "avg_launch_us": per_launch_s * 1e6, "l2_warm_frac": BYTES_PER_ENV_STEP_F32 * N / (warm_ms * 1e-3 / args.steps) / 1e9 / peaks["hbm_gbs"]}

    cpu = None
    if not args.no_cpu:
        probe_steps, probe_dt, cores = cpu_oracle_run(1 << 13, 16)
        rate = probe_steps / probe_dt
        n_envs = 1 << 17
        n_steps = int(min(1000, max(16, rate * 12 / n_envs)))                 # ~12 s of CPU work
        steps_done, dt, cores = cpu_oracle_run(n_envs, n_steps)
        cpu = {"value": steps_done / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_envs} envs x {n_steps} env-steps ({dt:.1f} s), same distributions, C oracle port (gcc -O2, OpenMP)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: 2^20 envs per GPU, fp32, random init + random combo refs + uniform stick, "
                               "gym step K=1 with autoreset, obs [N,5] + reward + done emitted every step",
                   "envs_per_gpu": N, "substeps_per_launch": 1, "fast_math": fast, "parallelism": f"env-shard x{world}, no collective",
                   "l2": "flushed before every timed step (256 MiB fill, outside the CUDA-event bracket)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "value_l2_warm": total_env_steps / (warm_ms * 1e-3), "value_cuda_graph_replay": total_env_steps / (graph_ms * 1e-3),
        "wall_s_timed_region_incl_flush": wall_s,
        "substeps_sweep_env_steps_per_s": sweep,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

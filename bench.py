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
WORKLOAD = ("BASELINE configs[2]: 2^20 envs per GPU, fp32, random init + random combo refs + uniform stick, "
            "gym step K=1 with autoreset, obs [N,5] + reward + done emitted every step")
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one streaming
    `nvidia-smi -lms` process, the recipe's query in B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        super().__init__(daemon=True)
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def run(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            parts = [c.strip() for c in line.split(",")]
            if len(parts) >= 7:
                self.samples.append((time.perf_counter(), parts))

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.join(timeout=5)
        rows = [p for t, p in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or [p for _, p in self.samples]

        def num(v):
            try:
                return int(float(v))
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        mx = [v for v in (num(r[1]) for r in rows) if v is not None]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------ CPU arm
def host_threads():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_CPU_INPUTS = {}


def cpu_oracle_run(n_envs, n_steps, threads=None, seed=0):
    """Time the CPU oracle on a bounded sample of the same workload: n_envs aircraft from the
    same initial-state distribution, uniform random stick, n_steps env-steps each."""
    import numpy as np

    from oracle import f16_oracle as orc
    import f16_model_b200 as f16

    key = (n_envs, n_steps)
    if key not in _CPU_INPUTS:       # drawn once per sample shape; every call rotates them by `seed` (untimed either way)
        rng = np.random.default_rng(12345)
        d5 = np.radians(5)
        al = rng.uniform(-d5, d5, n_envs)
        x0 = np.zeros((n_envs, 9))
        x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n_envs), rng.uniform(-d5, d5, n_envs), al
        x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n_envs), al
        acts = rng.uniform(-1, 1, (n_steps, n_envs))
        bank = f16.build_reference_bank(0.01, 10, "combo", False, bank_size=64, seed=0)
        idx = rng.integers(0, bank.shape[0], n_envs).astype(np.int32)
        _CPU_INPUTS.clear()
        _CPU_INPUTS[key] = (x0, acts, bank, idx)
    x0, acts, bank, idx = _CPU_INPUTS[key]
    if seed:
        shift = (int(seed) * 7919) % n_envs
        x0, acts, idx = np.roll(x0, shift, axis=0), np.roll(acts, shift, axis=1), np.roll(idx, shift)
    threads = threads or host_threads()
    t0 = time.perf_counter()
    out = orc.rollout(x0, acts, bank, idx, want_states=False, want_obs=True, nthreads=threads)
    dt = time.perf_counter() - t0
    return out["env_steps"], dt, threads


def reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path -- here the C oracle port,
    all host threads.  Rank 0 only."""
    if rank != 0:
        return
    # one bench step = a bounded sample of the 2^20-env gym step: 16 consecutive env-steps of n_envs aircraft, n_envs sized from
    # a probe so that the K timed steps take about 45 s of CPU work whatever K is (2^16 envs = one full 2^20-env-step batch)
    n_steps = 16
    probe_steps, probe_dt, _ = cpu_oracle_run(1 << 14, n_steps)
    per_env_step = probe_dt / max(probe_steps, 1)
    n_envs = 1 << 16
    while n_envs > 1024 and args.steps * n_envs * n_steps * per_env_step > 45.0:
        n_envs //= 2
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
        "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs, "substeps_per_launch": 1,
                   "reference_arm": "bounded CPU sample of that workload in fp64 (the reference's arithmetic): " + sample},
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
    ap.add_argument("--steps", type=int, default=12000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--no-sweep", action="store_true", help="skip the K-substeps sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fast-math", type=int, default=1)
    ap.add_argument("--pool", type=int, default=64, help="rows of the cycled stick tensor per batch (period of each env's stick, in steps)")
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
    # R_SETS independent batches of N envs resident in HBM, stepped round-robin: the cycle touches
    # R_SETS x 77 MB (state + actions + outputs) = 308 MB >> 126 MB of L2, so every step's inputs come
    # from HBM and its dirty lines are written back while the next batches run -- the steady state of a
    # streaming workload, with no flush kernel inside the timed region.
    R_SETS, POOL = 4, max(1, args.pool)
    envs = [make_sharded_env(CFG, N * world, rank, world, device, dtype=torch.float32, autoreset=True, seed=1 + j,
                             fast_math=fast, bank_size=256) for j in range(R_SETS)]
    env = envs[0]
    g = torch.Generator(device=device).manual_seed(1000 + rank)
    pools = [torch.rand((POOL, N), device=device, generator=g) * 2 - 1 for _ in range(R_SETS)]   # stick tensors, cycled
    pool = pools[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)          # > 126 MB L2

    def enqueue_steps(n_steps, first=0):
        for s in range(first, first + n_steps):
            envs[s % R_SETS].step(pools[s % R_SETS][(s // R_SETS) % POOL])

    for _ in range(max(args.warmup, 3)):
        enqueue_steps(R_SETS)
    torch.cuda.synchronize()

    # The K = 1 gym step is launch-bound from Python (~30 us of interpreter + ctypes per call vs a
    # ~25 us kernel), so the timed loop replays a CUDA graph of G consecutive steps.
    G = min(max(48, R_SETS * POOL), args.steps)   # one graph covers a whole cycle of the stick pool
    graphs = {}

    def graph_of(n_steps):
        if n_steps not in graphs:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                enqueue_steps(R_SETS)
            torch.cuda.current_stream(device).wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                enqueue_steps(n_steps)
            graphs[n_steps] = gr
        return graphs[n_steps]

    def timed_concurrent(n_steps):
        """Extra (not the headline): the R_SETS batches are independent, so their steps may overlap -- one capture
        stream per batch inside the graph hides each kernel's ramp-up / ramp-down behind its neighbours."""
        n_steps = (n_steps // G) * G or G
        gr = torch.cuda.CUDAGraph()
        side = [torch.cuda.Stream(device) for _ in range(R_SETS)]
        with torch.cuda.graph(gr):
            cur = torch.cuda.current_stream(device)
            for st_ in side:
                st_.wait_stream(cur)
            for s in range(G):
                with torch.cuda.stream(side[s % R_SETS]):
                    envs[s % R_SETS].step(pools[s % R_SETS][(s // R_SETS) % POOL])
            for st_ in side:
                cur.wait_stream(st_)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps // G):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), n_steps

    def timed_resident(n_steps):
        """Extra (not the headline, breaks the L2 rule on purpose): ONE batch stepped repeatedly, as a PPO loop
        does -- its 77 MB working set stays in the 126 MB L2."""
        n_steps = (n_steps // G) * G or G
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for s in range(G):
                envs[0].step(pools[0][s % POOL])
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps // G):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), n_steps

    def timed_graph_steps(n_steps):
        plan = [G] * (n_steps // G) + ([n_steps % G] if n_steps % G else [])
        for k in set(plan):
            graph_of(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for k in plan:
            graphs[k].replay()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)                                             # ms on the device for exactly n_steps

    def timed_steps(n_steps, do_flush=True):
        """Secondary measurement: eager API calls, one CUDA-event pair per step, optional write-flush of L2."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for s in range(n_steps):
            if do_flush:
                flush.fill_(s & 0xFF)
            ev[s][0].record()
            env.step(pool[s % POOL])
            ev[s][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    timed_graph_steps(min(args.steps, 4 * G))                                  # warm the graphs themselves
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.1)
    t_wall0 = time.perf_counter()
    dev_ms = timed_graph_steps(args.steps)
    t_wall1 = time.perf_counter()
    wall_s = t_wall1 - t_wall0
    n_sec = min(args.steps, 500)
    flushed_ms = timed_steps(n_sec, do_flush=True)
    eager_t0 = time.perf_counter()
    enqueue_steps(n_sec)
    torch.cuda.synchronize()
    eager_s = time.perf_counter() - eager_t0
    conc_ms, conc_steps = timed_concurrent(min(args.steps, 960))
    res_ms, res_steps = timed_resident(min(args.steps, 960))
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    dev_ms = max_over_ranks(dev_ms)
    flushed_ms = max_over_ranks(flushed_ms)
    eager_s = max_over_ranks(eager_s)
    conc_ms = max_over_ranks(conc_ms)
    res_ms = max_over_ranks(res_ms)
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
        n_e2e = max(10, min(args.steps, 100))
        t0 = time.perf_counter()
        for s in range(n_e2e):
            env.step_host(a_host[s % 4], out=out)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": N * world * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": 25 * N,
               "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e,
               "path": ("f16_env_step_host, zero-copy route: ONE kernel launch reads the pinned (mapped) host actions and bulk-stores "
                        "obs/reward/done to pinned host memory over PCIe, then a stream sync (PCIe-bound)"
                        if getattr(env, "last_host_route", "") == "zero_copy" else
                        "f16_env_step_host, staged route: pinned host actions -> H2D -> kernel -> D2H obs/reward/done, "
                        "3 chunks on 3 streams (PCIe-bound)")}
        if os.environ.get("F16B200_BENCH_STAGED_TOO"):   # comparison leg: same call with pageable actions -> staged route
            a_np = a_host.numpy().copy()
            t0 = time.perf_counter()
            for s in range(n_e2e):
                env.step_host(a_np[s % 4], out=out)
            e2e["value_staged_route_pageable_actions"] = N * world * n_e2e / max_over_ranks(time.perf_counter() - t0)

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
                "traffic": traffic, "peak_kind": f"of {peak_kind}", "kernel": "env_step2_kernel<float,fast,single> (two envs per thread)",
                "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP_F32, "units_per_launch": N,
                "avg_launch_us": per_launch_s * 1e6,
                "timing": "CUDA events around the whole timed region (graph replays of consecutive steps) / steps: includes inter-kernel gaps",
                "frac_single_launch_write_flushed_l2": BYTES_PER_ENV_STEP_F32 * N / (flushed_ms * 1e-3 / n_sec) / 1e9 / peaks["hbm_gbs"]}

    cpu = None
    if not args.no_cpu and world == 1:   # the CPU baseline leg belongs to the N = 1 run (the other ranks' processes would share its cores)
        probe_steps, probe_dt, cores = cpu_oracle_run(1 << 13, 16)
        rate = probe_steps / probe_dt
        n_envs, n_steps = 1 << 17, 500
        reps = int(min(64, max(1, round(rate * 12 / (n_envs * n_steps)))))    # ~12 s of CPU work
        steps_done = dt = 0
        for r in range(reps):
            s_, d_, cores = cpu_oracle_run(n_envs, n_steps, seed=r)
            steps_done += s_
            dt += d_
        cpu = {"value": steps_done / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} x ({n_envs} envs x {n_steps} env-steps) = {steps_done:.3g} env-steps in {dt:.1f} s, same "
                         "distributions as the GPU workload, C oracle port (gcc -O2, OpenMP)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": N, "substeps_per_launch": 1, "fast_math": fast, "parallelism": f"env-shard x{world}, no collective",
                   "l2": "inputs larger than L2: 4 resident batches of 2^20 envs stepped round-robin (308 MB per cycle vs 126 MB L2), "
                         "no flush kernel in the timed region",
                   "stick": f"U(-1,1) rows cycled from a pool of {POOL} per batch (each env's stick repeats every {POOL} steps)",
                   "launch": f"CUDA graph of {G} consecutive steps, replayed"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu,
        "value_eager_python_loop": N * world * n_sec / eager_s,
        "value_4_batches_on_4_streams": N * world * conc_steps / (conc_ms * 1e-3),
        "value_one_batch_l2_resident": N * world * res_steps / (res_ms * 1e-3),
        "value_single_launch_events_write_flushed_l2": N * world * n_sec / (flushed_ms * 1e-3),
        "wall_s_timed_region": wall_s,
        "substeps_sweep_env_steps_per_s": sweep,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

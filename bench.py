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

Timing: the timed region is a CUDA graph of consecutive steps replayed on one stream, one
CUDA-event pair around the whole region, max over ranks; four resident batches are stepped
round-robin so that every step's inputs come from HBM (308 MB per cycle vs 126 MB of L2).
The region is measured in the STATIONARY state of the workload, whatever --steps / --warmup
are: before the first timed step every batch has flown >= 128 env-steps (untimed K-step
launches) and the graph has been replayed for >= 0.3 s so that the board sits at the clocks
it sustains; `roofline.frac_early` is the same measurement on freshly reset envs.
`kernels` carries a roofline object for every other kernel of the library (policy, PPO
gradient / optimiser, GAE, trim search, reset), each against the peak that bounds it, with
pipe peaks measured in the same run (f16_peak_probe).
`e2e` repeats the measurement through the host-buffer entry point (f16_env_step_host).
`cpu_baseline` / `--impl reference` time the reference's CPU path on the box's host cores:
the C restatement (oracle/f16_oracle.c, OpenMP, all threads) on the GPU arm's own step shape
(2^20 envs x 1 env-step per bench step), and -- when baseline/_ref is present -- the
reference's unmodified Python `F16.step` in one process per core.
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
PREROLL_ENV_STEPS = 128           # env-steps every batch has flown before the timed region starts
SUSTAIN_S = 0.3                   # untimed graph replays before the timed region (sustained clocks)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"      # B200_PROFILING.md fallback


def profile_constants():
    """Per-launch constants that only ncu can give (DRAM bytes, FP64 instructions per trim search), captured once per
    change and committed under profiles/ (tools/ncu_summary.py writes the file)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


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
        rows = [p for t, p in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        window = "timed region"
        if not rows:      # a region shorter than the sampling period: the samples of the sustain phase right before it
            rows = [p for t, p in self.samples if t0 is None or (t0 - 0.25 <= t <= (t1 or t0) + 0.05)] or [p for _, p in self.samples]
            window = "last 0.25 s before and during the timed region (region shorter than the sampling period)"

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
                "samples": len(rows), "power_w_max": max(pw) if pw else None, "window": window}


# ------------------------------------------------------------------------------------ CPU arm
def host_threads():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuPort:
    """The C restatement of the reference's path (oracle/f16_oracle.c, OpenMP) stepping ONE resident batch the way the GPU
    arm does: every bench step advances all n_envs aircraft by one env-step with a fresh stick row and emits
    obs / reward / done; finished episodes are re-initialised from the workload's initial-state distribution between
    steps (the port freezes an env at `done`; its caller resets it, as the reference's SyncVectorEnv would)."""

    def __init__(self, n_envs, threads=None, seed=12345):
        import numpy as np

        from oracle import f16_oracle as orc
        import f16_model_b200 as f16

        self.np, self.orc = np, orc
        self.n, self.threads = int(n_envs), threads or host_threads()
        self.rng = np.random.default_rng(seed)
        self.bank = f16.build_reference_bank(0.01, 10, "combo", False, bank_size=64, seed=0)
        self.x = self._draw(self.n)
        self.idx = self.rng.integers(0, self.bank.shape[0], self.n).astype(np.int32)
        self.sticks = self.rng.uniform(-1, 1, (16, self.n))
        self.k = 0

    def _draw(self, n):
        np, rng = self.np, self.rng
        d5 = np.radians(5)
        al = rng.uniform(-d5, d5, n)
        x0 = np.zeros((n, 9))
        x0[:, 1], x0[:, 2], x0[:, 3] = rng.uniform(2000, 4000, n), rng.uniform(-d5, d5, n), al
        x0[:, 4], x0[:, 5] = rng.uniform(90, 250, n), al
        return x0

    def step(self, n_steps=1):
        """-> (env-steps done, seconds inside the oracle call)."""
        acts = self.sticks[[(self.k + j) % 16 for j in range(n_steps)]]
        self.k += n_steps
        t0 = time.perf_counter()
        out = self.orc.rollout(self.x, acts, self.bank, self.idx, want_states=False, want_obs=True, nthreads=self.threads)
        dt = time.perf_counter() - t0
        self.x = out["final_state"]
        done = out["done"][-1].astype(bool)
        if done.any():                                    # untimed: the caller-side reset of the episodes that ended
            self.x[done] = self._draw(int(done.sum()))
        return out["env_steps"], dt


_PYREF_WORKER = r'''
import os, sys, time, random, warnings
warnings.simplefilter("ignore")
sys.path[:0] = [os.environ["F16_PYREF_SHIMS"], os.environ["F16_PYREF_ROOT"]]
import numpy as np
from F16model.env import F16
budget, seed = float(sys.argv[1]), int(sys.argv[2])
random.seed(seed); np.random.seed(seed)
cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": True, "scenario": "combo"}
env = F16(cfg); env.reset()
rng = np.random.default_rng(seed)
for _ in range(20):                                   # warm-up: imports, SciPy caches
    env.step(np.array([0.0]))
env.reset()
steps, a = 0, 0.0
t0 = time.perf_counter()
while True:
    if steps % 25 == 0:
        a = float(rng.uniform(-0.3, 0.1))             # piecewise-constant stick, as in BASELINE.md section 3
        if time.perf_counter() - t0 >= budget:
            break
    _, _, done, _, _ = env.step(np.array([a]))
    steps += 1
    if done:
        env.reset()
print(steps, time.perf_counter() - t0, flush=True)
'''


def python_reference_leg(budget_s=6.0, workers=None):
    """The reference's OWN `F16model.env.F16.step` (unmodified, from baseline/_ref, with the stand-ins of oracle/shims for
    its four absent third-party imports), one process per host core, config-1-style episodes (random initial state,
    deterministic combo reference, piecewise-constant stick) for `budget_s` seconds each."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    shims = os.path.join(ROOT, "oracle", "shims")
    if not os.path.isdir(os.path.join(ref_root, "F16model")):
        return {"unavailable": "baseline/_ref/F16model is not installed (python -m pip install --no-index --no-deps --target baseline/_ref <reference>)"}
    workers = workers or host_threads()
    env = dict(os.environ, F16_PYREF_ROOT=ref_root, F16_PYREF_SHIMS=shims, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1",
               MKL_NUM_THREADS="1")
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", _PYREF_WORKER, str(budget_s), str(1000 + w)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for w in range(workers)]
    steps, secs, errs = [], [], []
    for p in procs:
        try:
            out, err = p.communicate(timeout=budget_s * 6 + 120)
        except subprocess.TimeoutExpired:
            p.kill()
            out, err = "", "timeout"
        try:
            s, d = out.split()[-2:]
            steps.append(int(s))
            secs.append(float(d))
        except Exception:
            errs.append((err or out).strip().splitlines()[-1:] or ["no output"])
    wall = time.perf_counter() - t0
    if not steps:
        return {"unavailable": f"the reference did not run: {errs[0][0][:200] if errs else 'unknown'}"}
    import numpy
    import scipy

    rate = sum(s / d for s, d in zip(steps, secs))
    return {"value": rate, "unit": UNIT, "cores": len(steps), "per_core": rate / len(steps), "kind": "reference",
            "sample": f"{len(steps)} processes x {budget_s:.0f} s of the unmodified F16model.env.F16.step (baseline/_ref + oracle/shims), "
                      f"{sum(steps)} env-steps, wall {wall:.1f} s incl. imports; numpy {numpy.__version__}, scipy {scipy.__version__}",
            "failed_workers": len(errs)}


def reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path, all host threads.  Rank 0 only."""
    if rank != 0:
        return
    n_envs = args.envs
    port = CpuPort(n_envs)
    # the reference arm runs the GPU arm's own step shape (n_envs aircraft x 1 env-step per bench step) unless K timed
    # steps of it would take more than ~60 s of CPU; then the batch is halved until they fit (and the line says so)
    steps0, dt0 = port.step()
    while n_envs > 4096 and args.steps * dt0 * (n_envs / port.n) > 60.0:
        n_envs //= 2
    if n_envs != port.n:
        port = CpuPort(n_envs)
    for _ in range(max(args.warmup, 1)):
        port.step()
    total, tsum = 0, 0.0
    for _ in range(args.steps):
        s, d = port.step()
        total += s
        tsum += d
    value = total / tsum
    same = n_envs == args.envs
    sample = (f"{n_envs} envs x 1 env-step per bench step (one resident batch stepped {args.steps} times), config-3 distributions, "
              "C restatement of the reference's path in fp64 (oracle/f16_oracle.c, gcc -O2, OpenMP)")
    pyref = python_reference_leg() if not args.no_pyref else None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * tsum / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs, "substeps_per_launch": 1, "same_config": same,
                   "reference_arm": "the same gym step on the host cores, in the reference's fp64 arithmetic: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.threads, "kind": "port", "sample": sample, "python_reference": pyref},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "value = the C port (the conservative comparator: ~1e4 x faster per core than the reference's Python); "
                "cpu_baseline.python_reference = the reference's own unmodified F16.step timed on the same cores",
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
    ap.add_argument("--no-pyref", action="store_true", help="skip the Python-reference sub-leg of the CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline list")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (eager loop, 4 streams, L2-resident, flushed)")
    ap.add_argument("--fast-math", type=int, default=1)
    ap.add_argument("--pool", type=int, default=64, help="rows of the cycled stick tensor per batch (period of each env's stick, in steps)")
    ap.add_argument("--bank", default="default", help="'default' = the library default (all 6480 random combo signals, segment bank), "
                                                      "a row count, or 'full' (all 6480 as whole rows)")
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
    from f16_model_b200 import _capi
    from f16_model_b200.sharding import make_sharded_env

    L = f16.load_library()                   # fails loudly if the CUDA library is missing
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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = args.envs
    fast = bool(args.fast_math)
    # R_SETS independent batches of N envs resident in HBM, stepped round-robin: the cycle touches
    # R_SETS x 77 MB (state + actions + outputs) = 308 MB >> 126 MB of L2, so every step's inputs come
    # from HBM and its dirty lines are written back while the next batches run -- the steady state of a
    # streaming workload, with no flush kernel inside the timed region.
    R_SETS, POOL = 4, max(1, args.pool)
    bank_kw = {} if args.bank == "default" else {"bank_size": args.bank if args.bank == "full" else int(args.bank)}
    envs = [make_sharded_env(CFG, N * world, rank, world, device, dtype=torch.float32, autoreset=True, seed=1 + j,
                             fast_math=fast, **bank_kw) for j in range(R_SETS)]
    env = envs[0]
    g = torch.Generator(device=device).manual_seed(1000 + rank)
    pools = [torch.rand((POOL, N), device=device, generator=g) * 2 - 1 for _ in range(R_SETS)]   # stick tensors, cycled
    pool = pools[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)          # > 126 MB L2

    def enqueue_steps(n_steps, first=0):
        for s in range(first, first + n_steps):
            envs[s % R_SETS].step(pools[s % R_SETS][(s // R_SETS) % POOL])

    def preroll(env_steps):
        """Advance every batch by `env_steps` env-steps, untimed, in K-step launches over the batch's own stick pool
        (K launches of one step and one launch of K steps are bit-identical: tests/test_gpu_parity.py)."""
        left = env_steps
        while left > 0:
            k = min(left, POOL)
            for j in range(R_SETS):
                envs[j].step_k(pools[j][:k])
            left -= k

    for _ in range(max(args.warmup, 3)):
        enqueue_steps(R_SETS)
    torch.cuda.synchronize()

    # The K = 1 gym step is launch-bound from Python (~30 us of interpreter + ctypes per call vs a
    # ~22 us kernel), so the timed loop replays a CUDA graph of G consecutive steps.
    # One graph covers a whole cycle of the stick pool, whatever --steps is: a shorter graph would replay only the first
    # rows of every pool, i.e. give every env a stick of a few steps' period -- a different (and slower: the periodic stick
    # spreads alpha over more table cells per quarter-warp) workload than the one the long run measures.
    G = R_SETS * POOL
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

    def time_replays(gr, n_replays, warm=3):
        for _ in range(warm):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_replays):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

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
        return time_replays(gr, n_steps // G), n_steps

    def timed_resident(n_steps):
        """Extra (not the headline, breaks the L2 rule on purpose): ONE batch stepped repeatedly, as a PPO loop
        does -- its 77 MB working set stays in the 126 MB L2."""
        n_steps = (n_steps // G) * G or G
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for s in range(G):
                envs[0].step(pools[0][s % POOL])
        return time_replays(gr, n_steps // G), n_steps

    def timed_graph_steps(n_steps, runway=2):
        """Device time [ms] of exactly n_steps steps.  Barrier + synchronize on both sides; between the opening barrier and
        the first event, `runway` untimed replays of the full-cycle graph are queued so that the timed steps start on a busy
        GPU (an event recorded on an idle stream would charge the region with the graph-launch latency and with a first
        kernel that has no predecessor to overlap its prologue with: +5 % on a 20-step region, nothing on a long one)."""
        plan = [G] * (n_steps // G) + ([n_steps % G] if n_steps % G else [])
        for k in set(plan):
            graph_of(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if runway:
            graph_of(G)
        barrier()
        for _ in range(runway):
            graphs[G].replay()
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

    # ---- bring the workload to its stationary state, untimed -------------------------------------------------
    # (1) every env flies PREROLL_ENV_STEPS env-steps: alpha / stab have spread over the table cells (the step kernel's
    #     shared-memory lookups depend on that: a launch on fresh envs is ~10 % faster than a stationary one);
    # (2) the timed graph itself is replayed for >= SUSTAIN_S so that the board settles at the clocks it sustains;
    # (3) the replays stop when the envs are ~mid-episode (age 300..700 of 1000), away from the synchronised time-out
    #     resets that put every env back into the early state for ~35 steps.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()                                                        # nvidia-smi needs ~0.5 s before its first sample
    preroll(PREROLL_ENV_STEPS)
    timed_graph_steps(args.steps if args.steps < G else G)                     # build and warm the graphs themselves
    torch.cuda.synchronize()
    flown = lambda: int(env.episode_length.float().median().item())          # noqa: E731  (typical env age of batch 0)
    gr_full = graph_of(G)
    t_s = time.perf_counter()
    sustain_replays = 0
    while time.perf_counter() - t_s < SUSTAIN_S or not (300 <= flown() <= 700):
        gr_full.replay()
        sustain_replays += 1
        if sustain_replays % 4 == 0:
            torch.cuda.synchronize()
        if sustain_replays > 4000:
            break
    torch.cuda.synchronize()
    age_at_start = flown()

    t_wall0 = time.perf_counter()
    dev_ms = timed_graph_steps(args.steps)
    t_wall1 = time.perf_counter()
    wall_s = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    barrier()

    # ---- the same measurement on freshly reset envs (the "early" state a short cold benchmark would see) -----
    n_early = 20 * R_SETS
    for j in range(R_SETS):
        envs[j].reset(seed=101 + j)
    enqueue_steps(3 * R_SETS)                                                  # W >= 3 warm-up steps per batch
    early_ms = max_over_ranks(timed_graph_steps(n_early, runway=0))            # (no runway: it would age the envs)
    preroll(PREROLL_ENV_STEPS)                                                 # back to the stationary state for everything below

    n_sec = min(args.steps, 500)
    extras = {}
    if not args.no_extras:
        flushed_ms = max_over_ranks(timed_steps(n_sec, do_flush=True))
        eager_t0 = time.perf_counter()
        enqueue_steps(n_sec)
        torch.cuda.synchronize()
        eager_s = max_over_ranks(time.perf_counter() - eager_t0)
        conc_ms, conc_steps = timed_concurrent(min(args.steps, 960))
        res_ms, res_steps = timed_resident(min(args.steps, 960))
        conc_ms, res_ms = max_over_ranks(conc_ms), max_over_ranks(res_ms)
        extras = {"value_eager_python_loop": N * world * n_sec / eager_s,
                  "value_4_batches_on_4_streams": N * world * conc_steps / (conc_ms * 1e-3),
                  "value_one_batch_l2_resident": N * world * res_steps / (res_ms * 1e-3),
                  "value_single_launch_events_write_flushed_l2": N * world * n_sec / (flushed_ms * 1e-3)}
        barrier()

    dev_ms = max_over_ranks(dev_ms)
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
               "d2h_gbs_per_gpu": 25 * N * n_e2e / e2e_s / 1e9,
               "path": ("f16_env_step_host, zero-copy route: ONE kernel launch reads the pinned (mapped) host actions and bulk-stores "
                        "obs/reward/done to pinned host memory over PCIe, then a stream sync (PCIe-bound)"
                        if getattr(env, "last_host_route", "") == "zero_copy" else
                        "f16_env_step_host, staged route: pinned host actions -> H2D -> kernel -> D2H obs/reward/done, "
                        "3 chunks on 3 streams (PCIe-bound)")}
        # the link ceiling the e2e number is to be read against: plain pinned D2H copies of the same 25 B x N, all ranks at once
        dsrc = torch.empty(25 * N, dtype=torch.uint8, device=device)
        hdst = torch.empty(25 * N, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            hdst.copy_(dsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        copy_s = max_over_ranks(time.perf_counter() - t0)
        e2e["d2h_copy_ceiling_gbs_per_gpu"] = 25 * N * 20 / copy_s / 1e9
        e2e["frac_of_concurrent_d2h_ceiling"] = e2e["d2h_gbs_per_gpu"] / e2e["d2h_copy_ceiling_gbs_per_gpu"]
        e2e["d2h_ceiling_note"] = f"cudaMemcpyAsync D2H of the same 25 B x N from pinned memory, {world} rank(s) concurrently, 20 copies back to back"
        del dsrc, hdst
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

    peaks, peak_kind = measured_peaks()
    consts = profile_constants()

    # ---- every other kernel of the library against the peak that bounds it ---------------------
    kernels = None
    if not args.no_kernels:
        del flush
        kernels = kernel_rooflines(torch, f16, _capi, L, device, rank, world, peaks, peak_kind, consts, sweep, N, clocks)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    per_launch_s = dev_ms * 1e-3 / args.steps
    achieved = BYTES_PER_ENV_STEP_F32 * N / per_launch_s / 1e9
    early_s = early_ms * 1e-3 / n_early
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": consts.get("env_step_f32_dram_bytes_per_launch"), "peak_kind": f"of {peak_kind}",
                "kernel": "env_step2_kernel<float,fast,single> (two envs per thread)",
                "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP_F32, "units_per_launch": N,
                "avg_launch_us": per_launch_s * 1e6,
                "state": f"stationary: every env had flown >= {PREROLL_ENV_STEPS} env-steps and the graph had been replayed for >= {SUSTAIN_S} s "
                         f"before the first timed step (median env age at the start of the region: {age_at_start} steps)",
                "frac_early": BYTES_PER_ENV_STEP_F32 * N / early_s / 1e9 / peaks["hbm_gbs"],
                "early_launch_us": early_s * 1e6,
                "early_state": f"{n_early} steps on freshly reset envs (3 warm-up steps per batch): alpha within +-5 deg, stab = 0",
                "timing": "CUDA events around the whole timed region (graph replays of consecutive steps) / steps: includes inter-kernel gaps"}
    if "value_single_launch_events_write_flushed_l2" in extras:
        roofline["frac_single_launch_write_flushed_l2"] = (BYTES_PER_ENV_STEP_F32 * extras["value_single_launch_events_write_flushed_l2"]
                                                           / world / 1e9 / peaks["hbm_gbs"])

    cpu = None
    if not args.no_cpu and world == 1:   # the CPU baseline leg belongs to the N = 1 run (the other ranks' processes would share its cores)
        port = CpuPort(N)
        port.step()                                                            # warm-up (page faults, thread pool)
        steps_done = dt = 0
        while dt < 10.0 and steps_done < 64 * N:                               # ~10 s of CPU work on the GPU arm's own step shape
            s_, d_ = port.step()
            steps_done += s_
            dt += d_
        cpu = {"value": steps_done / dt, "unit": UNIT, "cores": port.threads, "kind": "port",
               "sample": f"{steps_done // N} bench steps of {N} envs x 1 env-step = {steps_done:.3g} env-steps in {dt:.1f} s, same "
                         "distributions and step shape as the GPU workload, C restatement of the reference (gcc -O2, OpenMP)",
               "python_reference": None if args.no_pyref else python_reference_leg()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": N, "substeps_per_launch": 1, "fast_math": fast, "parallelism": f"env-shard x{world}, no collective",
                   "l2": "inputs larger than L2: 4 resident batches of 2^20 envs stepped round-robin (308 MB per cycle vs 126 MB L2), "
                         "no flush kernel in the timed region",
                   "stick": f"U(-1,1) rows cycled from a pool of {POOL} per batch (each env's stick repeats every {POOL} steps)",
                   "reference_signals": f"{env._cfg.n_ref} ({env.ref_mode})",
                   "launch": f"CUDA graphs of consecutive steps ({G} = one cycle of the stick pool, plus one for the remainder), replayed; "
                             "2 untimed replays of the cycle graph are queued ahead of the first event (busy GPU at the start)",
                   "untimed_preroll": f"{max(args.warmup, 3)} warm-up cycles, {PREROLL_ENV_STEPS} env-steps per env in K-step launches, "
                                      f"{sustain_replays} graph replays ({SUSTAIN_S} s) to sustained clocks"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels,
        "wall_s_timed_region": wall_s,
        "substeps_sweep_env_steps_per_s": sweep,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(torch, f16, _capi, L, device, rank, world, peaks, peak_kind, consts, sweep, N, clocks):
    """One roofline object per kernel of the library that is not the headline step kernel.  Every duration is the
    average of a CUDA graph of back-to-back launches timed with one CUDA-event pair (cold-launch effects excluded, the
    working set of one launch may stay in L2 -- said per entry); pipe peaks come from f16_peak_probe in this run."""
    import ctypes as C

    import numpy as np

    out = []
    st = lambda: _capi.stream_ptr(device)     # noqa: E731

    def graph_us(fn, reps=20, replays=5):
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * replays)

    # ---- measured pipe peaks (this board, this run's clocks) ----
    scratch = torch.zeros(16, device=device)
    pipe = {}
    for name, kind, iters in (("fp32_fma", 0, 4000), ("fp64_fma", 1, 600), ("mufu_tanh", 2, 1500)):
        ops = C.c_double(0.0)
        us = graph_us(lambda: _capi.check(L.f16_peak_probe(kind, iters, C.byref(ops), scratch.data_ptr(), st())), reps=2, replays=3)
        pipe[name] = ops.value / (us * 1e-6)
    hbm = peaks["hbm_gbs"]
    how = "graph of back-to-back launches, CUDA events"

    def entry(kernel, bound, achieved, peak, unit, us, units, **kw):
        e = {"kernel": kernel, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak if peak else None,
             "avg_launch_us": us, "units_per_launch": units, "timing": how}
        e.update(kw)
        out.append(e)

    cfg = dict(CFG)
    # ---- K-step env kernel (issue-bound): warp-instructions per second against 4 schedulers x n_sm x clock ----
    if sweep and "64" in sweep and clocks and clocks.get("sm_mhz"):
        instr = consts.get("env_step2_k_sass_instr_per_env_step", 248)
        ach = sweep["64"] / world * instr / 32.0
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        peak = 4.0 * n_sm * clocks["sm_mhz"] * 1e6
        entry("env_step2_kernel<float,fast,K-step> (K = 64, state in registers)", "issue", ach / 1e9, peak / 1e9, "G warp-instr/s",
              64 * N / (sweep["64"] / world) * 1e6, 64 * N, peak_kind=f"4 warp schedulers x {n_sm} SMs x {clocks['sm_mhz']} MHz (sampled in this run)",
              algorithmic_per_unit=f"{instr} SASS instructions per env-step on the K-loop's common path (cuobjdump; two envs per thread)",
              env_steps_per_s=sweep["64"] / world, fp32_fma_peak_tflops=2 * pipe["fp32_fma"] / 1e12)

    # ---- env_reset_kernel (HBM: 85 B per env: 9 state rows + counters + obs written, ref_idx / episode_no / bank read) ----
    renv = f16.F16VecEnv(cfg, N, device=device, dtype=torch.float32, seed=5, fast_math=True, bank_size=256)
    us = graph_us(lambda: renv.reset(), reps=10)
    entry("env_reset_kernel<float>", "hbm", 85 * N / (us * 1e-6) / 1e9, hbm, "GB/s", us, N, peak_kind=f"of {peak_kind}",
          algorithmic_per_unit="85 B per env (36 state + 13 counters/flags + 4 ref row + 20 obs written; 12 read)",
          l2="one batch (89 MB) re-written per launch: write-back overlaps the next launch")

    # ---- policy_forward_kernel (MUFU.TANH: 256 tanh per env) ----
    pol = f16.TensorCorePolicy(device=device, seed=1, logstd_init=-1.0)
    obs = renv.reset()[0]
    act_out = (torch.empty(N, device=device), torch.empty(N, device=device), torch.empty(N, device=device))
    for n_pol in sorted({N, min(N, 65536)}, reverse=True):
        o = obs[:n_pol]
        oo = tuple(t[:n_pol] for t in act_out)
        for prec, name, kpad, tdiv in (("fp16", "policy_forward16_kernel (tcgen05 kind::f16, four 128-row groups per CTA, biases in the MMAs; "
                                                "TensorCorePolicy.act from 49,152 rows up)", 16, 1.0),
                                       ("tf32", "policy_forward_kernel (tcgen05 kind::tf32, one 128-row tile per 256-thread CTA; "
                                                "TensorCorePolicy.act below 49,152 rows)", 8, 2.0)):
            us = graph_us(lambda: pol.act(o, step=0, out=oo, precision=prec))
            tf = 2 * (kpad * 128 + 2 * 64 * 64) * n_pol / (us * 1e-6) / 1e12
            entry(name, "mufu", 256 * n_pol / (us * 1e-6) / 1e9, pipe["mufu_tanh"] / 1e9,
                  "G tanh/s", us, n_pol, peak_kind="MUFU.TANH results/s of f16_peak_probe in this run (16 per clock per SM)",
                  algorithmic_per_unit="256 tanh per env (two 5-64-64-1 MLPs)",
                  tensor_tflops=tf, tensor_frac=tf / peaks["bf16_tflops"] * tdiv,
                  tensor_note="MACs issued (K padded 5 -> %d) against %s the measured bf16 rate; irrelevant for a 5-64-64-1 MLP"
                              % (kpad, "HALF" if tdiv == 2.0 else ""))
    del renv

    # ---- PPO: config 4 (65,536 envs x 128 steps): rollout, GAE, gradient, optimiser ----
    n4, T4 = 65536, 128
    hist = None
    for fused in ((False, True) if world == 1 else ()):     # multi-rank PPO is measured by tools/config_bench.py under torchrun
        envs4 = f16.F16VecEnv(cfg, n4, device=device, dtype=torch.float32, seed=1, fast_math=True)
        pol4 = f16.TensorCorePolicy(device=device, seed=1, logstd_init=-1.0)
        pcfg = f16.PPOConfig(num_steps=T4, num_minibatches=32, update_epochs=4, fused_rollout=fused)
        tr = f16.PPOTrainer(envs4, pol4, pcfg)
        hist = tr.train(num_updates=3)
        h = hist[-1]
        ach = 268 * n4 * T4 / h["rollout_s"] / 1e9
        entry(("rollout_kernel: 128 closed-loop steps of policy (tcgen05 kind::f16) -> Gaussian head -> F16.step in ONE persistent launch"
               if fused else "PPO rollout, two-kernel path: 128 x (policy_forward_kernel -> env_step2_kernel), one CUDA graph"),
              "mufu", ach, pipe["mufu_tanh"] / 1e9, "G MUFU/s", h["rollout_s"] * 1e6, n4 * T4,
              peak_kind="MUFU.TANH results/s of f16_peak_probe in this run (every MUFU op has that rate)",
              algorithmic_per_unit="268 MUFU per env-step: 256 tanh of the two MLPs + 12 in the env step (sin/cos, lg2/ex2, rcp)",
              timing="wall clock around reset + rollout + GAE + sync (PPOTrainer.rollout)",
              env_steps_per_s=h["rollout_sps"], update_s=h["update_s"], minibatches_per_update=pcfg.num_minibatches * pcfg.update_epochs)
    if hist:
        # the same fused kernel at the bench's batch size: 2^20 envs, K = 16 closed-loop steps per launch
        envs_big = f16.F16VecEnv(cfg, N, device=device, dtype=torch.float32, seed=2, fast_math=True)
        Kb = 16
        bufs = dict(obs_next=torch.zeros((Kb, N, 5), device=device), actions=torch.zeros((Kb, N), device=device),
                    logprobs=torch.zeros((Kb, N), device=device), values=torch.zeros((Kb, N), device=device),
                    rewards=torch.zeros((Kb, N), device=device), dones=torch.zeros((Kb, N), dtype=torch.uint8, device=device),
                    next_value=torch.zeros(N, device=device))
        obs0 = envs_big.reset()[0].clone()
        blob16 = f16.pack_train_blob16(pol4.actor_mean, pol4.critic)
        us = graph_us(lambda: f16.fused_rollout(envs_big, blob16, pol4.actor_logstd.detach(), obs0, **bufs, policy_seed=1), reps=2, replays=3)
        entry("rollout_kernel at the bench's batch size: 2^20 envs x 16 closed-loop steps per launch", "mufu",
              268 * N * Kb / (us * 1e-6) / 1e9, pipe["mufu_tanh"] / 1e9, "G MUFU/s", us, N * Kb,
              peak_kind="MUFU.TANH results/s of f16_peak_probe in this run",
              algorithmic_per_unit="268 MUFU per env-step", us_per_closed_loop_step_of_all_envs=us / Kb,
              env_steps_per_s=N * Kb / (us * 1e-6), hbm_gbs=37.0 * N * Kb / (us * 1e-6) / 1e9,
              note="the two-kernel path needs policy_forward_kernel + env_step2_kernel per step (see their entries)")
        del envs_big, bufs
        b = tr._flat
        mb = tr._perm_buf[: tr.minibatch_size]
        adv, ret = f16.gae(tr.rewards, tr.values, tr.done_after, tr.values[-1].clone(), 0.99, 0.95)
        us = graph_us(lambda: f16.gae(tr.rewards, tr.values, tr.done_after, tr.values[-1], 0.99, 0.95), reps=5)
        entry("gae_kernel (reverse scan, one thread per env)", "hbm", 17 * n4 * T4 / (us * 1e-6) / 1e9, hbm, "GB/s", us, n4 * T4,
              peak_kind=f"of {peak_kind}", algorithmic_per_unit="17 B per (step, env): reward, value, done read; advantage, return written",
              l2="151 MB per launch, larger than L2")
        # packed minibatch rows (f16_ppo_pack_rows, once per update) and the keyed-bijection permutation (once per epoch)
        us = graph_us(lambda: _capi.check(L.f16_ppo_pack_rows(tr.batch_size, b["obs"].data_ptr(), b["actions"].data_ptr(), b["logprobs"].data_ptr(),
                                                               b["adv"].data_ptr(), b["ret"].data_ptr(), b["val"].data_ptr(), tr._rows.data_ptr(),
                                                               tr._rv.data_ptr(), st())), reps=5)
        entry("ppo_pack_rows_kernel (six rollout arrays -> one 32-byte + one 8-byte record per row, once per update)", "hbm",
              80 * tr.batch_size / (us * 1e-6) / 1e9, hbm, "GB/s", us, tr.batch_size, peak_kind=f"of {peak_kind}",
              algorithmic_per_unit="80 B per rollout row: 40 read (obs[5], action, logprob, adv, ret, val), 40 written", l2="671 MB per launch, larger than L2")
        us = graph_us(lambda: _capi.check(L.f16_ppo_randperm(tr.batch_size, 12345, tr._perm_buf.data_ptr(), st())), reps=5)
        entry("ppo_randperm_kernel (minibatch order of an epoch as a keyed bijection of [0, n): no sort)", "hbm",
              8 * tr.batch_size / (us * 1e-6) / 1e9, hbm, "GB/s", us, tr.batch_size, peak_kind=f"of {peak_kind}",
              algorithmic_per_unit="8 B written per element (int64 index); torch.randperm: 0.9 ms for the same 2^23 elements",
              l2="67 MB per launch: fits L2, the write-back overlaps the next launch")
        args16 = (mb.numel(), tr._train_blob.data_ptr(), tr._rows.data_ptr(), None, None, None, tr._rv.data_ptr(), None, mb.data_ptr(),
                  pol4.actor_logstd.data_ptr(), tr._adv_stats.data_ptr(), tr._g.data_ptr(), 0.5, 0.5, 0.0, 1, 1)
        rows = mb.numel()
        us = graph_us(lambda: _capi.check(L.f16_ppo_grad16(*args16, tr._grad_scratch.data_ptr(), st())), reps=10)
        flops = 2.0 * rows * (16 * 128 + 2 * 64 * 64 + 2 * 64 * 64 + 128 * 128 + 16 * 128 + 16 * 128)   # MACs issued per row: fwd L1, L2; dZ2 W2; dW2 (both nets in one M=N=128 product); db2|dW1
        entry("ppo_grad16w_kernel + ppo_grad_reduce (forward + clipped-PPO loss + backward of both MLPs, tcgen05 kind::f16; 512 threads, one warp group per tile slot, "
              "token-ordered weight-gradient batches, packed rows)", "mufu",
              256 * rows / (us * 1e-6) / 1e9, pipe["mufu_tanh"] / 1e9, "G tanh/s", us, rows,
              peak_kind="MUFU.TANH results/s of f16_peak_probe in this run",
              algorithmic_per_unit="256 tanh per minibatch row (the forward); the backward reuses H1, H2",
              tensor_tflops=flops / (us * 1e-6) / 1e12, tensor_frac=flops / (us * 1e-6) / 1e12 / peaks["bf16_tflops"],
              tensor_note="fp16 MACs issued against the measured bf16 burst rate: the tensor cores are not the bound of a 64-wide MLP",
              gather_gbs=rows * (32 + 8 + 8) / (us * 1e-6) / 1e9,
              gather_note="per row: one 32-byte record, one 8-byte (ret, val) pair, one 8-byte index")
        m, v, vmax, step = tr._adam
        us = graph_us(lambda: _capi.check(L.f16_ppo_adam(
            tr._n_params, tr._pflat.data_ptr(), tr._g.data_ptr(), m.data_ptr(), v.data_ptr(), vmax.data_ptr(), step.data_ptr(), tr._lr.data_ptr(),
            0.9, 0.999, 1e-8, 0.5, tr._fwd_perm.data_ptr(), pol4._blob.data_ptr(), tr._fwd_perm.numel(), tr._train_perm.data_ptr(),
            tr._blob_tail.data_ptr(), tr._train_perm.numel(), tr._half_perm.data_ptr(), tr._train_blob.data_ptr(), tr._half_perm.numel(),
            tr._diag.data_ptr(), pol4.actor_logstd.data_ptr(), 1.0 / rows, 1, st())), reps=10)
        nbytes = tr._n_params * 4 * 9 + (tr._fwd_perm.numel() + tr._train_perm.numel()) * 12 + tr._half_perm.numel() * 10
        entry("ppo_adam_kernel (clip_grad_norm + Adam(amsgrad) + re-packing of three weight blobs, ONE CTA)", "latency",
              nbytes / (us * 1e-6) / 1e9, hbm, "GB/s", us, tr._n_params, peak_kind=f"of {peak_kind}",
              algorithmic_per_unit="9.3 k parameters: ~0.9 MB touched per launch, all L2-resident",
              note="a 9.3 k-parameter model has no throughput roofline: the launch is a chain of three CTA-wide reductions / barriers; "
                   "its cost is latency (us), which the PDL chain of the update graph overlaps with the next gradient kernel's prologue")
        # the trainer's default minibatch step: gradient kernel + ONE multi-CTA kernel (reduction, clip, Adam, blob scatter)
        if getattr(tr, "_step16", None) is not None:
            us_step = graph_us(lambda: tr._fused_minibatch_step(mb, adv_stats=tr._adv_stats), reps=10)
            us_grad = [e for e in out if e.get("kernel", "").startswith("ppo_grad16w_kernel")][-1]["avg_launch_us"]
            entry("f16_ppo_step16 = ppo_grad16w_kernel + ppo_reduce_adam_kernel (gradient reduction + clip_grad_norm + Adam(amsgrad) + blob scatter in one "
                  "289-CTA launch with a grid-wide arrival counter): one whole minibatch step, two launches", "mufu",
                  256 * rows / (us_step * 1e-6) / 1e9, pipe["mufu_tanh"] / 1e9, "G tanh/s", us_step, rows,
                  peak_kind="MUFU.TANH results/s of f16_peak_probe in this run",
                  algorithmic_per_unit="256 tanh per minibatch row; the optimiser half has no throughput roofline (9.3 k parameters)",
                  optimizer_half_us=us_step - us_grad,
                  note="optimizer_half_us = this entry minus the gradient + plain-reduction entry above: what reduction + clip + Adam + re-packing cost "
                       "inside the chain (the separate single-CTA ppo_adam_kernel below costs its whole launch)")
            # the trainer's default on one rank: ALL minibatch steps of an epoch in one persistent cooperative launch (optimiser step inside the kernel)
            n_mb = tr.batch_size // tr.minibatch_size
            tr._adv_stats_epoch[:, 0], tr._adv_stats_epoch[:, 1] = tr._adv_stats[0], tr._adv_stats[1]
            us_epoch = graph_us(lambda: tr._fused_minibatch_step(tr._perm_buf[: tr.minibatch_size], adv_stats=tr._adv_stats_epoch[0], epoch_minibatches=n_mb),
                                reps=2, replays=3)
            us_mb = us_epoch / n_mb
            entry("f16_ppo_epoch16 = ppo_grad16w_kernel<EPOCH> (persistent over the %d minibatch steps of an epoch: gradient, flush, three grid-wide meetings around "
                  "reduction / norm / clip + Adam(amsgrad) + blob scatter, weight reload -- ONE cooperative launch per epoch); per MINIBATCH step" % n_mb, "mufu",
                  256 * rows / (us_mb * 1e-6) / 1e9, pipe["mufu_tanh"] / 1e9, "G tanh/s", us_mb, rows,
                  peak_kind="MUFU.TANH results/s of f16_peak_probe in this run",
                  algorithmic_per_unit="256 tanh per minibatch row; the optimiser step has no throughput roofline (9.3 k parameters)",
                  launch_us=us_epoch, minibatches_per_launch=n_mb, optimizer_half_us=us_mb - us_grad,
                  note="optimizer_half_us = this entry minus the gradient + plain-reduction entry above (which includes a reduction launch of its own)")
        del tr, envs4, pol4

    # ---- trim search (FP64 pipe) ----
    Vg, Hg = np.meshgrid(np.linspace(100, 300, 21), np.linspace(1000, 10000, 19))
    starts = f16.reference_starts()[::4]
    f16.trim_grid(Vg[:2], Hg[:2], starts=starts, device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f16.trim_grid(Vg, Hg, starts=starts, device=device, return_all=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    searches = Vg.size * starts.shape[0]
    iters = float(np.sum(r["all_iterations"]))
    per_iter = consts.get("trim_fp64_thread_instr_per_iteration")
    ach = iters * per_iter / dt if per_iter else None
    e = {"kernel": "trim_kernel<double> (Nelder-Mead searches, scipy.optimize.fmin rules, as a warp-synchronous state machine: one cost evaluation per loop trip)", "bound": "fp64",
         "achieved": ach / 1e9 if ach else None, "peak": pipe["fp64_fma"] / 1e9, "unit": "G fp64-instr/s",
         "frac": ach / pipe["fp64_fma"] if ach else None, "avg_launch_us": dt * 1e6, "units_per_launch": searches,
         "timing": "wall clock around one trim_grid call (upload, kernel, arg-min, download)",
         "peak_kind": "DFMA/s of f16_peak_probe in this run", "searches_per_s": searches / dt, "nelder_mead_iterations": iters,
         "algorithmic_per_unit": ("%s FP64-pipe thread-instructions per Nelder-Mead iteration (ncu, profiles/traffic.json)" % per_iter) if per_iter
         else "FP64 instructions per iteration not captured yet (profiles/traffic.json)"}
    out.append(e)
    out.append({"pipe_peaks_measured_in_this_run": {"fp32_tflops": 2 * pipe["fp32_fma"] / 1e12, "fp64_tflops": 2 * pipe["fp64_fma"] / 1e12,
                                                    "mufu_tanh_per_s": pipe["mufu_tanh"]}})
    return out


if __name__ == "__main__":
    main()

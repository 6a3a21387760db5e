# r02 final profiling pass (one B200): plain bench + reference arm first, then the ncu passes on commands that exited 0 without ncu.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err || { echo bench failed; tail -5 gpurun_out/bench_r02d.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02d_ref.json 2> gpurun_out/bench_r02d_ref.err; echo ref rc=$?
SHORT="--steps 400 --warmup 10 --no-sweep --no-cpu --no-e2e --no-kernels --no-extras"
python bench.py $SHORT > gpurun_out/bench_r02d_short.json 2>&1 || exit 1
# (a) launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02d_bench_launches.csv python bench.py $SHORT > gpurun_out/ncu_r02d_l.log 2>&1; echo launches rc=$?
# (b) the step kernel in the stationary state (a launch late in the run)
ncu --set full --clock-control none --import-source on -k regex:env_step2_kernel -s 900 -c 2 -f -o gpurun_out/prof_r02d_env_step2 python bench.py $SHORT > gpurun_out/ncu_r02d_e.log 2>&1; echo env_step2 rc=$?
# (c) rollout / gradient / optimiser kernels of one PPO iteration at config 4, policy forward at 2^20
cat > /tmp/upd.py <<'P'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
envs = f16.F16VecEnv(cfg, 65536, device="cuda", dtype=torch.float32, seed=1, fast_math=True)
pol = f16.TensorCorePolicy(device="cuda", seed=1, logstd_init=-1.0)
tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=1))
h = tr.train(num_updates=2)
big = f16.F16VecEnv(cfg, 1 << 20, device="cuda", dtype=torch.float32, seed=2, fast_math=True)
o = big.reset()[0]
for k in range(3):
    pol.act(o, step=k)
torch.cuda.synchronize()
print(h[-1])
P
python /tmp/upd.py > gpurun_out/upd_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02d_rollout python /tmp/upd.py > gpurun_out/ncu_r02d_r.log 2>&1; echo rollout rc=$?
ncu --set full --clock-control none --import-source on -k regex:ppo_grad16_kernel -s 40 -c 1 -f -o gpurun_out/prof_r02d_grad16 python /tmp/upd.py > gpurun_out/ncu_r02d_g.log 2>&1; echo grad16 rc=$?
ncu --set full --clock-control none --import-source on -k regex:policy_forward_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02d_policy python /tmp/upd.py > gpurun_out/ncu_r02d_p.log 2>&1; echo policy rc=$?
# (d) the trim kernel: FP64 thread-instructions per Nelder-Mead iteration
cat > /tmp/trim.py <<'P'
import numpy as np, torch, sys, os, json
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
Vg, Hg = np.meshgrid(np.linspace(100, 300, 5), np.linspace(1000, 10000, 5))
starts = f16.reference_starts()[::4]
r = f16.trim_grid(Vg, Hg, starts=starts, device="cuda", return_all=True)
torch.cuda.synchronize()
print(json.dumps({"searches": int(Vg.size * starts.shape[0]), "iterations": float(np.sum(r["all_iterations"]))}))
P
python /tmp/trim.py > gpurun_out/r02d_trim_plain.json 2>&1 || exit 1
ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum --clock-control none -k regex:trim_kernel -c 1 --csv --log-file gpurun_out/r02d_trim_metrics.csv python /tmp/trim.py > gpurun_out/ncu_r02d_t.log 2>&1; echo trim rc=$?
ls -la gpurun_out/prof_r02d_*.ncu-rep

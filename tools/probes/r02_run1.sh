# r02 first pass: GPU tests, smoke, full bench, rollout-kernel ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo pytest rc=$? 
tail -3 gpurun_out/pytest_gpu_c.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo bench rc=$?
tail -c 600 gpurun_out/bench_r02c.err
cat > /tmp/roll.py <<'P'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
envs = f16.F16VecEnv(cfg, 65536, device="cuda", dtype=torch.float32, seed=1, fast_math=True)
pol = f16.TensorCorePolicy(device="cuda", seed=1, logstd_init=-1.0)
tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=4, fused_rollout=True))
h = tr.train(num_updates=3)
print(h[-1])
P
python /tmp/roll.py > gpurun_out/roll_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2a_rollout python /tmp/roll.py > gpurun_out/ncu_r2a.log 2>&1
echo ncu rc=$?

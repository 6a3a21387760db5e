# launch list (ncu, serialised, cold) of one PPO iteration at config 4: which kernels make up the 17 ms update
mkdir -p gpurun_out
cat > /tmp/upd.py <<'P'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
envs = f16.F16VecEnv(cfg, 65536, device="cuda", dtype=torch.float32, seed=1, fast_math=True)
pol = f16.TensorCorePolicy(device="cuda", seed=1, logstd_init=-1.0)
tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=int(os.environ.get("EPOCHS", "4"))))
h = tr.train(num_updates=int(os.environ.get("UPDATES", "3")))
print(h[-1])
P
python /tmp/upd.py > gpurun_out/upd_plain.log 2>&1 || exit 1
tail -1 gpurun_out/upd_plain.log | cut -c1-400
EPOCHS=1 UPDATES=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_update_launches.csv python /tmp/upd.py > gpurun_out/ncu_upd.log 2>&1
echo ncu rc=$?

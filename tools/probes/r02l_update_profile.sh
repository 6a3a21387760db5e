# r02l: one PPO iteration at config 4 with the final kernels: plain run, ncu launch list (serialised, cold), one --set full capture of
# ppo_reduce_adam_kernel (after the plain run exited 0)
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
timeout 120 python /tmp/upd.py > gpurun_out/r02l_upd_plain.log 2>&1 || exit 1
tail -1 gpurun_out/r02l_upd_plain.log | cut -c1-300
EPOCHS=1 UPDATES=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_update_launches.csv python /tmp/upd.py > gpurun_out/ncu_r02l_l.log 2>&1
echo launches rc=$?
EPOCHS=1 UPDATES=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppo_reduce_adam_kernel -s 10 -c 1 -f -o gpurun_out/prof_r02l_reduce_adam python /tmp/upd.py > gpurun_out/ncu_r02l_r.log 2>&1
echo reduce_adam rc=$?

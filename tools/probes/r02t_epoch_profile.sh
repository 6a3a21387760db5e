# r02t: the final round-2 kernels of one PPO iteration at config 4 (65,536 envs x 128 steps, 32 minibatches x 4 epochs): plain run, the ncu
# launch list of ONE iteration (serialised, cold caches), --set full captures of the persistent epoch kernel and of the paced rollout kernel,
# each after the plain run of the same script exited 0.  Also the probes of the same kernels (plain).
mkdir -p gpurun_out
cat > /tmp/upd.py <<'P'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
cfg = {"dt": 0.01, "tn": 10, "debug_state": False, "determenistic_ref": False, "scenario": "combo"}
envs = f16.F16VecEnv(cfg, 65536, device="cuda", dtype=torch.float32, seed=1, fast_math=True)
pol = f16.TensorCorePolicy(device="cuda", seed=1, logstd_init=-1.0)
tr = f16.PPOTrainer(envs, pol, f16.PPOConfig(num_steps=128, num_minibatches=32, update_epochs=int(os.environ.get("EPOCHS", "4"))))
h = tr.train(num_updates=int(os.environ.get("UPDATES", "6")))
print({k: h[-1][k] for k in ("update_s", "rollout_s", "value_loss", "approx_kl")}, [round(x["update_s"], 5) for x in h])
P
timeout 120 python /tmp/upd.py > gpurun_out/r02t_upd_plain.log 2>&1 || exit 1
tail -1 gpurun_out/r02t_upd_plain.log | cut -c1-300
timeout 100 python tools/probes/grad16_probe.py > gpurun_out/r02t_grad16_probe.txt 2>&1; tail -1 gpurun_out/r02t_grad16_probe.txt | cut -c1-300
EPOCHS=1 UPDATES=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:ppo_grad16w_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02t_epoch python /tmp/upd.py > gpurun_out/ncu_r02t_e.log 2>&1
echo epoch rc=$?

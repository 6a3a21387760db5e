# r02g: ncu captures of the kernels added / changed late in round 2 (after their commands exited 0 without ncu)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_policy.py -m gpu -x -q 2>&1 | tail -3
python tools/probes/grad16_probe.py > gpurun_out/r02g_grad16_probe.txt 2>&1; tail -2 gpurun_out/r02g_grad16_probe.txt
cat > /tmp/p16.py <<'P'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import f16_model_b200 as f16
pol = f16.TensorCorePolicy(device="cuda", seed=1, logstd_init=-1.0)
obs = torch.rand((1 << 20, 5), device="cuda") * 2 - 1
for k in range(3):
    pol.act(obs, step=k, precision="fp16")
torch.cuda.synchronize()
print("ok")
P
python /tmp/p16.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:policy_forward16_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02g_policy16 python /tmp/p16.py > gpurun_out/ncu_r02g_p.log 2>&1; echo policy16 rc=$?

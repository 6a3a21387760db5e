mkdir -p gpurun_out
python bench.py --steps 400 --warmup 10 --no-sweep --no-cpu --no-e2e > gpurun_out/plain_l2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:env_step2_kernel -s 60 -c 2 -f -o gpurun_out/prof_r1l python bench.py --steps 400 --warmup 10 --no-sweep --no-cpu --no-e2e > gpurun_out/ncu_l2.log 2>&1
echo ncu rc=$?
ls -la gpurun_out/prof_r1l.ncu-rep

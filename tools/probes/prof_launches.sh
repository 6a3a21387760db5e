# launch list of the bench command (run plain first, then the same command under ncu)
mkdir -p gpurun_out
python bench.py --steps 400 --warmup 10 --no-sweep --no-cpu --no-e2e > gpurun_out/plain_l.log 2>&1 || exit 1
tail -1 gpurun_out/plain_l.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_l.csv python bench.py --steps 400 --warmup 10 --no-sweep --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
echo ncu rc=$?
grep -c env_step2 gpurun_out/launches_l.csv

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/probes/zero_copy_probe.py 2>&1 | tail -8 | tee gpurun_out/zero_copy.txt
python tools/e2e_probe.py | tee -a gpurun_out/zero_copy.txt
F16B200_HOST_ZERO_COPY=0 python tools/e2e_probe.py | tee -a gpurun_out/zero_copy.txt
for rep in 1 2; do timeout 300 python bench.py --no-cpu --no-sweep 2>/dev/null | tee -a gpurun_out/bench_exp2.jsonl | cut -c1-400; done

mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do timeout 300 python bench.py --no-cpu 2>/dev/null | tee -a gpurun_out/bench_exp3.jsonl | cut -c1-200; done

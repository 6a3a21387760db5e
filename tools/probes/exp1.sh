set -x
mkdir -p gpurun_out
./build/ffma2_probe > gpurun_out/ffma2.txt 2>&1
cat gpurun_out/ffma2.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in "" _pf1 _pf2 _s3 _s3pf2; do
  for rep in 1 2; do
  F16B200_LIB=$PWD/f16-model_b200/lib$v/libf16b200.so timeout 300 python bench.py --no-cpu --no-e2e --no-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VAR$v', d['ms_per_step'], d['value'], {k: v for k, v in d.items() if k.startswith('value_')})
" | tee -a gpurun_out/variants.txt
  done
done
F16B200_RAW_COPY=1 python tools/e2e_probe.py | tee -a gpurun_out/e2e.txt
for cr in "1 1" "3 1" "3 2" "4 2" "5 2" "6 2" "4 3" "5 1.5" "6 1.5" "8 1.5" "4 4"; do set -- $cr; F16B200_HOST_CHUNKS=$1 F16B200_HOST_RATIO=$2 python tools/e2e_probe.py | tee -a gpurun_out/e2e.txt; done

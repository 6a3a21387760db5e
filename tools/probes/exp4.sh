mkdir -p gpurun_out
for v in _c128 _c512; do for rep in 1 2; do
F16B200_LIB=$PWD/f16-model_b200/lib$v/libf16b200.so timeout 300 python bench.py --no-cpu --no-e2e --no-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VAR$v', d['ms_per_step'], d['value'], {k: round(v/1e9,2) for k, v in d.items() if k.startswith('value_')})
" | tee -a gpurun_out/variants4.txt
done; done

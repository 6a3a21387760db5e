mkdir -p gpurun_out
for v in "" _c128 _c64; do
echo "== lib$v" | tee -a gpurun_out/variants5.txt
F16B200_LIB=$PWD/f16-model_b200/lib$v/libf16b200.so timeout 300 python tools/probes/size_sweep.py 2>&1 | grep -E "2\^(20|21|22|23)" | tee -a gpurun_out/variants5.txt
for rep in 1 2; do
F16B200_LIB=$PWD/f16-model_b200/lib$v/libf16b200.so timeout 300 python bench.py --no-cpu --no-e2e --no-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VAR$v', d['ms_per_step'], d['value'], {k: round(v/1e9,2) for k, v in d.items() if k.startswith('value_')})
" | tee -a gpurun_out/variants5.txt
done; done

mkdir -p gpurun_out
for v in f16-model_b200/lib f16-model_b200/lib_k1; do
for rep in 1 2; do
F16B200_LIB=$PWD/$v/libf16b200.so timeout 300 python bench.py --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VAR $v', round(d['ms_per_step']*1e3,2), {k: round(v/1e9,1) for k, v in d['substeps_sweep_env_steps_per_s'].items()})
" | tee -a gpurun_out/variants6.txt
done; done

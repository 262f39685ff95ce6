set -x
python -m pytest tests/test_ops_gpu.py -q -x -k "residual" 2>&1 | tail -3
for r in 3; do
 for m in 0 1; do
  TFX_ARES_DEFERRED=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_ab_ares_${m}_${r}.json 2> gpurun_out/r02_ab_ares_${m}_${r}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_ab_ares_${m}_${r}.json').read().strip().splitlines()[-1])
print('ARES', $m, $r, d['ms_per_step'], d['value'])
for k in d['roofline'].get('kernels', []):
    if 'residual' in k.get('kernel',''): print('   ', k['kernel'], k['launches'], k['us_per_launch'], k['achieved'], k['frac'])
PY
 done
done

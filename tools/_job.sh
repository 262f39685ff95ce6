set -x
TFX_GEMM_CLUSTER=2 timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" 2>&1 | tail -3
for m in 0 1; do
  TFX_GEMM_CLUSTER=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_cl_${m}.json 2> gpurun_out/r02_cl_${m}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_cl_${m}.json').read().strip().splitlines()[-1])
print('CLUSTER', $m, d['ms_per_step'], d['value'], d['roofline']['families']['gemm(tcgen05)'])
for k in d['roofline'].get('kernels', []):
    if 'gemm' in k.get('kernel',''): print('   ', k['kernel'], k['launches'], k['us_per_launch'], k['achieved'], k['frac'])
PY
done
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -3

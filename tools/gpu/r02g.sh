# r02g: rolled dense kernel, three-way full-size comparison, CFL mixed failure, partitions with batched launches,
# whole time steps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zlate.py -m gpu -q -x -k "wavespeed" 2>&1 | grep -E "Error|assert|passed|failed|mismatch|Max|x:|y:" | head -20
PYFR_B200_PARITY_TAG=r02g timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_partitions.py -m gpu -q -k "full_size_rhs or partition" 2>&1 | tail -8
python - <<PY
import json
for r in json.load(open('gpurun_out/parity_errors_r02g.json')):
    if 'full-size' in r['test']: print(json.dumps(r))
PY
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --kernel-times gpurun_out/r02g_kt_mixed.json > gpurun_out/r02g_bench_mixed.json 2> gpurun_out/r02g_bench_mixed.err; tail -2 gpurun_out/r02g_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02g_kt_mixed.json')); b = json.loads(open('gpurun_out/r02g_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step']); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:14])
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --timestep --kernel-times gpurun_out/r02g_kt_ts.json > gpurun_out/r02g_b_ts.json 2> gpurun_out/r02g_b_ts.err; tail -2 gpurun_out/r02g_b_ts.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02g_b_ts.json').read())
print('timestep', round(b['value'],2), json.dumps(b['time_step']))
PY

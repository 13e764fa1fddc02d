# r02k (1 GPU): small tuning sweep on the headline configuration (thread counts of gradflux, row groups of opmul,
# occupancy of intcflux)
mkdir -p gpurun_out
for o in "gradflux-threads=384" "gradflux-threads=640" "mul-rowgroups=6" "mul-rowgroups=8" "cflux-minblocks=4" "cflux-minblocks=6"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt $o --kernel-times gpurun_out/r02k_kt_$o.json > gpurun_out/r02k_b_$o.json 2> gpurun_out/r02k_b_$o.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02k_kt_$o.json')); b = json.loads(open('gpurun_out/r02k_b_$o.json').read())
print('$o', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
done
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --kernel-times gpurun_out/r02k_kt_mixed.json > gpurun_out/r02k_bench_mixed.json 2> gpurun_out/r02k_bench_mixed.err; tail -2 gpurun_out/r02k_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02k_kt_mixed.json')); b = json.loads(open('gpurun_out/r02k_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step']); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:14])
PY
timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --timestep > gpurun_out/r02k_b_ts.json 2> gpurun_out/r02k_b_ts.err; tail -2 gpurun_out/r02k_b_ts.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02k_b_ts.json').read())
print('timestep', round(b['value'],2), json.dumps(b['time_step']))
PY
PYFR_B200_PARITY_TAG=r02k timeout 900 python -m pytest tests/test_gpu_zlate.py -m gpu -q 2>&1 | tail -5

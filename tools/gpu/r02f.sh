# r02f: parity (full-size local criterion, reference host on the device, mixed + late features with the dense
# kernel), dense-operator kernel timing (configs[3]), configs[4] proxy at full size, whole time steps, P3 variant
mkdir -p gpurun_out
PYFR_B200_PARITY_TAG=r02f timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_reference_dropin.py tests/test_gpu_zlate.py -m gpu -q --durations=3 -k "full_size_rhs or reference_host or mixed or fused or rkvdh2 or reduction" 2>&1 | tail -14
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --kernel-times gpurun_out/r02f_kt_mixed.json > gpurun_out/r02f_bench_mixed.json 2> gpurun_out/r02f_bench_mixed.err; tail -2 gpurun_out/r02f_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02f_kt_mixed.json')); b = json.loads(open('gpurun_out/r02f_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step']); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:14])
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt gradflux-metric-late=1 --kernel-times gpurun_out/r02f_kt_late.json > gpurun_out/r02f_b_late.json 2> gpurun_out/r02f_b_late.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02f_kt_late.json')); b = json.loads(open('gpurun_out/r02f_b_late.json').read())
print('metric-late', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --timestep --kernel-times gpurun_out/r02f_kt_ts.json > gpurun_out/r02f_b_ts.json 2> gpurun_out/r02f_b_ts.err; tail -2 gpurun_out/r02f_b_ts.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02f_b_ts.json').read())
print('timestep', round(b['value'],2), json.dumps(b['time_step']))
PY
# configs[4] proxy at its stated size: 96^3 hexes, p = 6, fp32 (automatic SoA width)
timeout 1200 python bench.py --n 96 --order 6 --precision single --steps 10 --warmup 3 --no-cpu --no-e2e --no-clocks --kernel-times gpurun_out/r02f_kt_p6_96.json > gpurun_out/r02f_b_p6_96.json 2> gpurun_out/r02f_b_p6_96.err; tail -2 gpurun_out/r02f_b_p6_96.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02f_kt_p6_96.json')); b = json.loads(open('gpurun_out/r02f_b_p6_96.json').read())
print('p6 96^3', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b['setup_s'], b['parity'], {k: (round(x['ms'],3), round(x['bytes']/x['ms']/1e6)) for k, x in d['kernels'].items()})
PY

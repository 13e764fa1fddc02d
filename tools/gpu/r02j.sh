# r02j (1 GPU): T_1 of the strong-scaling series (128^3 on one rank, 64-bit indices), mixed case with the
# final dense kernel, per-kernel times of a fused RK stage
mkdir -p gpurun_out
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --kernel-times gpurun_out/r02j_kt_mixed.json > gpurun_out/r02j_bench_mixed.json 2> gpurun_out/r02j_bench_mixed.err; tail -2 gpurun_out/r02j_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02j_kt_mixed.json')); b = json.loads(open('gpurun_out/r02j_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step']); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:14])
PY
timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --timestep > gpurun_out/r02j_b_ts.json 2> gpurun_out/r02j_b_ts.err; tail -2 gpurun_out/r02j_b_ts.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02j_b_ts.json').read())
print('timestep', round(b['value'],2), json.dumps(b['time_step']))
PY
timeout 1500 python bench.py --scaling strong --mesh-n 128 --steps 10 --warmup 3 --no-cpu --no-e2e --no-clocks --no-parity --kernel-times gpurun_out/r02j_kt_strong1.json > gpurun_out/r02j_b_strong1.json 2> gpurun_out/r02j_b_strong1.err; tail -3 gpurun_out/r02j_b_strong1.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02j_kt_strong1.json')); b = json.loads(open('gpurun_out/r02j_b_strong1.json').read())
print('strong1', round(b['value'],2), round(b['ms_per_step'],3), b['index_bits'], round(b['setup_s']), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY

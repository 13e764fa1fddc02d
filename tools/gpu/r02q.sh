# r02q (1 GPU): end-of-round state: smoke, the whole device suite, the default bench line, launch list and full ncu
# capture of one RHS (four kernels), mixed-element mesh after the copy elision, whole time steps
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PYFR_B200_PARITY_TAG=r02q timeout 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40
timeout 900 python bench.py --kernel-times gpurun_out/r02q_kt64.json > gpurun_out/r02q_bench64.json 2> gpurun_out/r02q_bench64.err; tail -c 2500 gpurun_out/r02q_bench64.json; tail -3 gpurun_out/r02q_bench64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-clocks --no-parity > gpurun_out/r02q_launches.log 2>&1; tail -2 gpurun_out/r02q_launches.log | cut -c1-200
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"opmul|gradflux|intcflux" -s 12 -c 4 -o gpurun_out/rhs64_r02q python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks --no-parity > gpurun_out/ncu_r02q.log 2>&1; tail -1 gpurun_out/ncu_r02q.log | cut -c1-200
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --no-cpu --no-e2e --kernel-times gpurun_out/r02q_kt_mixed.json > gpurun_out/r02q_bench_mixed.json 2> gpurun_out/r02q_bench_mixed.err; tail -2 gpurun_out/r02q_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02q_kt_mixed.json')); b = json.loads(open('gpurun_out/r02q_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b.get('parity')); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:10])
PY
timeout 900 python bench.py --steps 12 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --timestep > gpurun_out/r02q_b_ts.json 2> gpurun_out/r02q_b_ts.err; tail -2 gpurun_out/r02q_b_ts.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02q_b_ts.json').read())
print('timestep', round(b['value'],2), json.dumps(b['time_step']))
PY
timeout 600 python bench.py --n 96 --order 6 --precision single --steps 10 --warmup 3 --no-cpu --no-e2e --no-clocks --kernel-times gpurun_out/r02q_kt_p6.json > gpurun_out/r02q_b_p6.json 2> gpurun_out/r02q_b_p6.err; tail -2 gpurun_out/r02q_b_p6.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02q_kt_p6.json')); b = json.loads(open('gpurun_out/r02q_b_p6.json').read())
print('p6 fp32 96^3', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b.get('parity'), b.get('rhs_model'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
